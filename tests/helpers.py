"""Shared checks: run an implementation (oracle or GPU product) against a golden pack."""
import hashlib

import numpy as np

SELECTORS = ("CIFE", "JMI", "MIM", "CIFEUnsup", "UniEntropy")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def parse_cols(key):
    body = key.split("c_", 1)[1]
    if body == "empty":
        return []
    return [int(t.replace("m", "-")) for t in body.split("_")]


def rel_close(a, b, tol=1e-9):
    """|a-b| <= tol*max(|a|,|b|) elementwise, with identical infinities."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False
    inf = np.isinf(a) | np.isinf(b)
    if not np.array_equal(a[inf], b[inf]):
        return False
    nan = np.isnan(a) | np.isnan(b)
    if not np.array_equal(np.isnan(a), np.isnan(b)):
        return False
    ok = ~(inf | nan)
    scale = np.maximum(np.abs(a[ok]), np.abs(b[ok]))
    return bool(np.all(np.abs(a[ok] - b[ok]) <= tol * np.maximum(scale, 1e-300) + 1e-300))


def check_entropies(infoset, pack, prefix="", exact=True):
    """Every wrt_*/ent_* entry of a golden pack against infoset.entropy_wrt / entropy."""
    n = 0
    for key in pack.files:
        if not key.startswith(prefix):
            continue
        k = key[len(prefix):]
        if k.startswith("wrt_c_"):
            got = infoset.entropy_wrt(np.array(parse_cols(k), dtype=np.int64))
        elif k.startswith("ent_c_"):
            got = np.float64(infoset.entropy(np.array(parse_cols(k), dtype=np.int64)))
        else:
            continue
        want = pack[key]
        assert np.asarray(got).dtype == np.float64
        if exact:
            assert np.array_equal(got, want), "%s differs (max abs %g)" % (
                key, np.max(np.abs(np.asarray(got) - want)))
        else:
            assert rel_close(got, want), key
        n += 1
    assert n > 0
    return n


def check_selectors(make, infoset, pack, n, kinds=SELECTORS, exact=True):
    for kind in kinds:
        fs = make(kind, infoset)
        cmp = np.array_equal if exact else rel_close
        assert cmp(np.asarray(fs.base_score), pack[kind + "_base"]), kind + " base"
        fs.autoselect(n)
        assert [int(s) for s in fs.S] == list(pack[kind + "_S"]), kind + " sequence"
        assert cmp(np.asarray(fs.score), pack[kind + "_score"]), kind + " score"
        assert cmp(np.asarray(fs.penalty), pack[kind + "_penalty"]), kind + " penalty"
