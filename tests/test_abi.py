"""The C-ABI library builds, loads and exports every symbol include/prb200.h declares
(no compute calls: there is no GPU in the CPU test tier)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "prb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(prb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("prb_stream_hits", "prb_finalize", "prb_finalize_direct", "prb_state_direct",
                 "prb_state_table", "prb_selector_update", "prb_argmax_final", "prb_qd_edges",
                 "prb_qd_digitize", "prb_term_table", "prb_sparse_entropy_wrt_host"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from picturedrocks_b200 import _native, build

    path = build.build()
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), "libprb200.so does not export %s" % name
    # the ctypes table binds exactly the declared entry points
    assert sorted(_native.EXPORTED) == declared_symbols()
    assert lib.prb_abi_version() == 3


def test_term_table_is_host_side_and_matches_libm():
    """prb_term_table runs on the host (no GPU needed) and must equal glibc's log2 terms."""
    import math

    import numpy as np
    from picturedrocks_b200 import _native

    for N in (8, 2730, 68579):
        out = np.empty(N + 1)
        _native.call("prb_term_table", N, out.ctypes.data)
        ref = np.array([0.0] + [-(c / N) * math.log2(c / N) for c in range(1, N + 1)])
        assert np.array_equal(out, ref)


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to compute (it never routes to the oracle)."""
    import numpy as np
    import pytest
    import torch

    import picturedrocks_b200 as pr

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        pr.markers.SparseInformationSet(np.eye(4, dtype=np.int64))
    src = ""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "picturedrocks_b200")):
        for f in files:
            if f.endswith(".py"):
                src += open(os.path.join(dirpath, f)).read()
    assert "oracle" not in src.replace("the oracle check", "").replace("lets the oracle", ""), \
        "product code must not reference oracle/"
