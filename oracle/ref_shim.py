"""Import shim for the REAL reference hot-path modules.

TEST / BASELINE INFRASTRUCTURE — not part of the product.  The files are loaded from
`/root/reference` (the authoring container) or, where that does not exist (the GPU box), from
the travelling copy `oracle/_ref/` that `oracle/make_ref.py` makes of the same three files,
unmodified.  Used by `tests/golden/make_golden.py` to generate golden vectors from the
unmodified reference, by `tests/test_oracle_vs_reference.py` to pin the oracle restatement,
and by `bench.py`'s CPU legs to time the reference's own numba path (`kind: "reference"`).

`import picturedrocks` itself fails here (anndata/scanpy/plotly/umap absent,
picturedrocks/__init__.py:7-10), so the three hot-path files are loaded with
stub parent packages registered in `sys.modules`:

* picturedrocks/markers/mutualinformation/infoset.py
* picturedrocks/markers/mutualinformation/iterative.py
* picturedrocks/markers/_iterative.py

Mode "jit" (default, the reference's intended execution mode): numba JIT on.
`_sparse_entropy_wrt` does not compile under numba 0.65 / Python 3.12 because
`yshift` is only assigned under `if ybits:` (infoset.py:354-355) and read under
a later `if ybits:` (infoset.py:387); the source text is loaded with that one
assignment hoisted above the `if` (no arithmetic change).
Mode "strict": `NUMBA_DISABLE_JIT=1` must be set in the environment before
numba is imported; the source is loaded unmodified.
"""
import importlib.util
import os
import sys
import types

_PROBE = "picturedrocks/markers/mutualinformation/infoset.py"


def _find_root():
    for root in (os.environ.get("PRB_REFERENCE_ROOT"), "/root/reference",
                 os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")):
        if root and os.path.isfile(os.path.join(root, _PROBE)):
            return root
    return "/root/reference"


REF_ROOT = _find_root()


def available():
    return os.path.isfile(os.path.join(REF_ROOT, _PROBE))


def _stub(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


def _load(name, path, patch=None):
    with open(path) as f:
        src = f.read()
    if patch is not None:
        src = patch(src)
    spec = importlib.util.spec_from_loader(name, loader=None, origin=path)
    mod = importlib.util.module_from_spec(spec)
    mod.__file__ = path
    sys.modules[name] = mod
    # numba needs real source lines for closures: register in linecache
    import linecache

    linecache.cache[path + ":shim"] = (len(src), None, src.splitlines(True), path + ":shim")
    exec(compile(src, path + ":shim", "exec"), mod.__dict__)
    return mod


def _hoist_yshift(src):
    old = (
        "        if ybits:\n"
        "            yshift = shift * (n_mcols + 1)\n"
    )
    new = (
        "        yshift = shift * (n_mcols + 1)\n"
        "        if ybits:\n"
    )
    assert src.count(old) == 1, "reference source changed; hoist patch does not apply"
    return src.replace(old, new)


_cache = {}


def load(mode=None):
    """Return (infoset_module, iterative_module) of the real reference."""
    if mode is None:
        mode = "strict" if os.environ.get("NUMBA_DISABLE_JIT") == "1" else "jit"
    if mode in _cache:
        return _cache[mode]
    assert available(), "reference tree not present (expected in the build container only)"
    if mode == "strict":
        assert os.environ.get("NUMBA_DISABLE_JIT") == "1", "strict mode needs NUMBA_DISABLE_JIT=1"
    base = os.path.join(REF_ROOT, "picturedrocks")
    _stub("picturedrocks", base)
    _stub("picturedrocks.markers", os.path.join(base, "markers"))
    _stub("picturedrocks.markers.mutualinformation", os.path.join(base, "markers/mutualinformation"))
    _load("picturedrocks.markers._iterative", os.path.join(base, "markers/_iterative.py"))
    infoset = _load(
        "picturedrocks.markers.mutualinformation.infoset",
        os.path.join(base, "markers/mutualinformation/infoset.py"),
        patch=None if mode == "strict" else _hoist_yshift,
    )
    iterative = _load(
        "picturedrocks.markers.mutualinformation.iterative",
        os.path.join(base, "markers/mutualinformation/iterative.py"),
    )
    _cache[mode] = (infoset, iterative)
    return _cache[mode]
