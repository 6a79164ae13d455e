"""Generate golden vectors from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py            # numba JIT mode (canonical)

Runs the reference's own hot-path functions (loaded by oracle/ref_shim.py from
/root/reference; JIT mode hoists one assignment so numba 0.65 can compile
_sparse_entropy_wrt, see the shim) on the fixtures of tests/fixtures.py and stores
the outputs as small .npz files next to this script.  /root/reference does not
exist on the GPU box, so tests only ever read the .npz files.
Versions used: numpy 2.3.5, scipy 1.18.1, numba 0.65.0, Python 3.12.3.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fixtures  # noqa: E402
from oracle import ref_shim  # noqa: E402
from picturedrocks_b200.synth import DiscretisedSpec, poisson_gamma  # noqa: E402

INFOSET, ITER = ref_shim.load()
SELECTORS = ("CIFE", "JMI", "MIM", "CIFEUnsup", "UniEntropy")
I64 = np.int64


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def cols_key(cols):
    return "c_" + "_".join(str(c).replace("-", "m") for c in cols) if cols else "c_empty"


def entropy_pack(infoset, col_sets, scalar=True):
    out = {}
    for cols in col_sets:
        c = np.array(cols, dtype=I64)
        out["wrt_" + cols_key(cols)] = infoset.entropy_wrt(c)
        if scalar:
            out["ent_" + cols_key(cols)] = np.float64(infoset.entropy(c))
    return out


def selector_pack(infoset, n, kinds=SELECTORS):
    out = {}
    for kind in kinds:
        fs = getattr(ITER, kind)(infoset)
        out[kind + "_base"] = fs.base_score.copy()
        fs.autoselect(n)
        out[kind + "_S"] = np.array([int(s) for s in fs.S], dtype=I64)
        out[kind + "_score"] = fs.score.copy()
        out[kind + "_penalty"] = fs.penalty.copy()
    return out


def main():
    # ---- sampleX: both classes, single / multi column sets -------------------------
    X, ent, mi = fixtures.sampleX()
    sets = [[], [0], [1], [2], [3], [4], [5], [0, 1], [0, 1, 2], [4, 5], [5, 4], [2, 3, 4]]
    save("sampleX", X=X,
         **{"sparse_" + k: v for k, v in entropy_pack(INFOSET.SparseInformationSet(X), sets).items()},
         **{"dense_" + k: v for k, v in entropy_pack(INFOSET.InformationSet(X, False), sets).items()})

    # ---- xor / diminish-red: the reference's selector known-answer tests ------------
    for name, (Xt, yt) in (("xor", fixtures.xor_table()), ("diminish_red", fixtures.diminish_red_table())):
        inf = INFOSET.SparseInformationSet(Xt)
        inf.set_y(yt)
        arrays = dict(X=Xt, y=yt)
        arrays.update(entropy_pack(inf, [[], [-1], [0], [-1, 0], [-1, 1], [0, 1], [-1, 0, 1]]))
        arrays.update(selector_pack(inf, 2))
        # reference tests/test_fs.py:76-91 trajectory
        for kind in ("CIFE", "JMI"):
            fs = getattr(ITER, kind)(inf)
            fs.autoselect(1)
            first = int(fs.S[0])
            fs.remove(first)
            fs.add(0)
            arrays[kind + "_traj_first"] = np.int64(first)
            arrays[kind + "_traj_score"] = fs.score.copy()
            fs.autoselect(1)
            arrays[kind + "_traj_S"] = np.array([int(s) for s in fs.S], dtype=I64)
        save(name, **arrays)

    # ---- linearxy through quantile_discretize ------------------------------------------
    X, y, _ = fixtures.linearxy()
    Xq = INFOSET.quantile_discretize(X)
    inf = INFOSET.SparseInformationSet(Xq, y)
    dense = INFOSET.InformationSet(np.concatenate((Xq, y[:, None]), axis=1), True)
    arrays = dict(X_sha=sha(X), codes=Xq.astype(np.uint8), y=y)
    arrays.update({"sparse_" + k: v for k, v in
                   entropy_pack(inf, [[], [-1], [5], [-1, 5]], scalar=False).items()})
    arrays.update({"dense_" + k: v for k, v in
                   entropy_pack(dense, [[], [-1], [5], [-1, 5]], scalar=False).items()})
    arrays["ent_single"] = np.array([inf.entropy(np.array([i])) for i in range(20)])
    arrays["ent_y_single"] = np.array([inf.entropy(np.array([-1, i])) for i in range(20)])
    save("linearxy", **arrays)

    # ---- planted count data: codes, entropies, all five selectors ----------------------
    X, y = fixtures.planted_counts(600, 250, 7, 3)
    Xq = INFOSET.quantile_discretize(X)
    inf = INFOSET.SparseInformationSet(Xq, y)
    arrays = dict(X=X, y=y, codes=Xq.astype(np.uint8))
    arrays.update(entropy_pack(inf, [[], [-1], [3], [-1, 3], [3, 7], [-1, 3, 7], [0, 1, 2], [249]]))
    arrays.update(selector_pack(inf, 12))
    # add/remove trajectory
    for kind in ("CIFE", "JMI", "CIFEUnsup"):
        fs = getattr(ITER, kind)(inf)
        fs.autoselect(4)
        fs.remove(int(fs.S[1]))
        fs.add(17)
        fs.autoselect(3)
        arrays[kind + "_mix_S"] = np.array([int(s) for s in fs.S], dtype=I64)
        arrays[kind + "_mix_score"] = fs.score.copy()
        arrays[kind + "_mix_penalty"] = fs.penalty.copy()
    save("planted600", **arrays)

    # ---- discretiser: dtypes, k, special columns ----------------------------------------
    Xc = fixtures.continuous_matrix(5)
    arrays = dict(X_sha=sha(Xc))
    for dt in (np.float32, np.float64):
        for k in (2, 3, 5, 10, 20):
            arrays["codes_%s_k%d" % (np.dtype(dt).name, k)] = \
                INFOSET.quantile_discretize(Xc.astype(dt), k).astype(np.uint8)
    Xi, _ = poisson_gamma(1500, 60, 5, 11, np.int64)
    arrays["int_X"] = Xi
    for k in (3, 5, 8):
        arrays["codes_int64_k%d" % k] = INFOSET.quantile_discretize(Xi, k).astype(np.uint8)
    save("discretize", **arrays)

    # ---- config-1 shaped run (BASELINE.json configs[0]): 2730 x 3451, 19 clusters -------
    X, y = poisson_gamma(2730, 3451, 19, 1)
    Xq = INFOSET.quantile_discretize(X)
    inf = INFOSET.SparseInformationSet(Xq, y)
    arrays = dict(X_sha=sha(X), codes_sha=sha(Xq.astype(np.uint8)), y=y,
                  code_hist=np.stack([(Xq == v).sum(0) for v in range(5)]))
    arrays.update(entropy_pack(inf, [[], [-1]], scalar=False))
    arrays.update(selector_pack(inf, 20, kinds=("CIFE", "JMI", "MIM")))
    save("config1", **arrays)

    # ---- counter-based synthetic sparse spec (configs 3-5 recipe) at small size ---------
    spec = DiscretisedSpec(N=3000, P=400, C=10, K=5, mean_density=0.05, seed=3)
    yk = spec.cpu_labels()
    Xs = spec.cpu_csc(range(spec.P), yk)
    inf = INFOSET.SparseInformationSet(Xs, yk)
    arrays = dict(indptr=Xs.indptr.astype(I64), indices_sha=sha(Xs.indices.astype(np.int32)),
                  data_sha=sha(Xs.data.astype(np.uint8)), y_sha=sha(yk))
    arrays.update(entropy_pack(inf, [[], [-1], [18], [-1, 18]], scalar=False))
    arrays.update(selector_pack(inf, 15, kinds=("CIFE", "JMI", "CIFEUnsup")))
    save("synth_sparse", **arrays)

    spec = DiscretisedSpec(N=2000, P=300, C=12, K=20, mean_density=0.08, seed=5)
    yk = spec.cpu_labels()
    Xs = spec.cpu_csc(range(spec.P), yk)
    inf = INFOSET.SparseInformationSet(Xs, yk)
    arrays = dict(indptr=Xs.indptr.astype(I64), data_sha=sha(Xs.data.astype(np.uint8)))
    arrays.update(entropy_pack(inf, [[], [-1], [2], [-1, 2]], scalar=False))
    arrays.update(selector_pack(inf, 8, kinds=("CIFE", "JMI")))
    save("synth_sparse_k20", **arrays)


if __name__ == "__main__":
    main()
