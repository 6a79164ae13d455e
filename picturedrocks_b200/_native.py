"""ctypes binding of libprb200.so (include/prb200.h).

There is deliberately NO fallback: if the shared library is missing and cannot
be built, importing the compute path raises.  PyTorch only provides device
memory and streams; every kernel runs from libprb200.so.
"""
import ctypes
import os

from . import build as _build

_LIB = None

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_i32 = ctypes.c_int32
_int = ctypes.c_int
_u32 = ctypes.c_uint32
_u64 = ctypes.c_uint64
_f64 = ctypes.c_double

# name -> argtypes (all return int unless listed in _RESTYPES)
_SIGS = {
    "prb_abi_version": [],
    "prb_last_error": [],
    "prb_device_info": [ctypes.POINTER(_int), ctypes.POINTER(_int)],
    "prb_term_table": [_i64, _vp],
    "prb_pack_csc": [_vp, _vp, _i64, _vp, _vp],
    "prb_unpack_csc": [_vp, _i64, _vp, _vp, _vp],
    "prb_pack_chunk": [_vp, _int, _vp, _int, _int, _i64, _i64, _vp, _vp, _vp],
    "prb_check_columns": [_vp, _vp, _i64, _vp, _vp],
    "prb_csr_count": [_vp, _vp, _int, _i64, _i64, _int, _vp, _vp, _vp],
    "prb_csr_fill": [_vp, _vp, _int, _vp, _int, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp],
    "prb_dense_count_nnz": [_vp, _i64, _i64, _i64, _vp, _vp],
    "prb_dense_fill_csc": [_vp, _i64, _i64, _i64, _vp, _vp, _vp],
    "prb_stream_plan": [_i64, _i64, _int, ctypes.POINTER(_i64), ctypes.POINTER(_int),
                        ctypes.POINTER(_int), ctypes.POINTER(_int), ctypes.POINTER(_i64)],
    "prb_tile_ptrs": [_vp, _vp, _i64, _i64, _int, _vp, _vp],
    "prb_stream_hits": [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _int, _int, _vp, _vp, _vp],
    "prb_v_tile_capacity": [ctypes.POINTER(_i64), ctypes.POINTER(_int)],
    "prb_v_count": [_vp, _vp, _i64, _vp, _vp, _int, _int, _vp, _vp, _vp],
    "prb_v_scatter": [_vp, _vp, _i64, _vp, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp],
    "prb_v_scatter16": [_vp, _vp, _i64, _vp, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp],
    "prb_v_tile_span": [],
    "prb_stream_v": [_vp, _vp, _vp, _int, _i64, _int, _i64, _vp, _vp, _vp, _int, _i64, _vp, _int,
                     _vp, _vp, _vp, _vp, _vp, _i64, _int, _vp],
    "prb_stream_v16": [_vp, _vp, _vp, _int, _i64, _int, _i64, _vp, _vp, _vp, _int, _i64, _vp, _int,
                       _vp, _vp, _vp, _vp, _vp, _i64, _int, _vp],
    "prb_v_warps": [],
    "prb_scatter_gene_u8": [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp],
    "prb_scatter_gene_or32": [_vp, _vp, _vp, _i64, _int, _vp, _vp],
    "prb_pick_scatter": [_vp, _int, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                         _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _int, _vp, _vp, _vp],
    "prb_peer_share": [_vp, _vp, _int, _int, _vp, _vp],
    "prb_v_scatter_flat": [_vp, _vp, _i64, _vp, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp],
    "prb_ell_build": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _int, _vp],
    "prb_ell2_geometry": [ctypes.POINTER(_int), ctypes.POINTER(_int), ctypes.POINTER(_int),
                          ctypes.POINTER(_int)],
    "prb_stream_ell2": [_vp, _vp, _vp, _vp, _int, _int, _int, _vp, _vp, _int, _int, _int, _vp, _int,
                        _i64, _vp, _vp],
    "prb_stream_ell": [_vp, _vp, _vp, _i64, _int, _vp, _vp, _vp, _int, _i64, _vp, _int, _vp, _vp, _vp,
                       _vp, _i64, _int, _i64, _vp, _vp],
    "prb_ell_warps": [],
    "prb_v_check_smem_base": [],
    "prb_finalize_update_blocks": [_i64, _int],
    "prb_finalize_update": [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _vp, _i64, _i64, _int, _int,
                            _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp,
                            _vp, _vp, _vp, _vp, _vp, _int, _vp, _int, _vp, _int, _int, _vp, _vp],
    "prb_wide_terms_stride": [_int, _int],
    "prb_wide_supported": [_int, _int],
    "prb_wide_chain_blocks": [_i64],
    "prb_finalize_update_wide": [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _vp, _i64, _i64, _vp, _vp,
                                 _vp, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _vp, _vp],
    "prb_v_unscatter": [_vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "prb_best_record": [_vp, _vp, _i64, _vp, _vp],
    "prb_state_direct": [_vp, _vp, _i64, _int, _int, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp],
    "prb_state_table": [_vp, _vp, _i64, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "prb_state_assign": [_vp, _vp, _i64, _int, _vp, _vp, _int, _vp],
    "prb_finalize": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _vp, _i64, _i64,
                     _vp, _vp, _vp],
    "prb_finalize_direct": [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _vp, _i64, _i64, _vp, _vp,
                            _int, _vp, _int, _vp],
    "prb_entropy_keys": [_vp, _vp, _i64, _int, _i64, _vp, _vp, _vp, _vp],
    "prb_mi_base": [_vp, _f64, _vp, _i64, _vp, _vp],
    "prb_selector_update": [_int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64,
                            _i32, _vp, _vp, _vp],
    "prb_argmax_blocks": [_vp, _i64, _i32, _vp, _vp, _vp],
    "prb_argmax_final": [_vp, _vp, _i64, _i64, _vp, _vp, _vp],
    "prb_commit_add": [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp],
    "prb_argmax_commit": [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp],
    "prb_transpose": [_vp, _int, _i64, _i64, _vp, _vp],
    "prb_sort_columns_workspace": [_int, _i64, _i64, ctypes.POINTER(_i64)],
    "prb_sort_columns": [_vp, _vp, _int, _i64, _i64, _vp, _i64, _vp],
    "prb_qd_edges": [_vp, _int, _i64, _i64, _int, _vp, _vp, _vp],
    "prb_qd_digitize": [_vp, _int, _i64, _i64, _int, _vp, _vp, _vp, _i64, _vp],
    "prb_sort_segments_workspace": [_int, _i64, _i64, ctypes.POINTER(_i64)],
    "prb_sort_segments": [_vp, _vp, _int, _i64, _i64, _vp, _vp, _i64, _vp],
    "prb_qd_edges_sparse": [_vp, _vp, _int, _i64, _i64, _int, _vp, _vp, _vp, _vp],
    "prb_qd_count_sparse": [_vp, _vp, _vp, _int, _i64, _int, _vp, _vp, _vp, _vp],
    "prb_qd_fill_sparse": [_vp, _vp, _vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _vp],
    "prb_synth_labels": [_vp, _i64, _int, _u64, _vp],
    "prb_synth_count": [_vp, _i64, _i64, _i64, _int, _int, _u32, _u64, _vp, _vp, _vp],
    "prb_synth_fill": [_vp, _i64, _i64, _i64, _int, _int, _u32, _u64, _vp, _vp, _vp, _vp],
    "prb_sparse_entropy_wrt_host": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64,
                                    _vp, _vp, _i64, _vp],
    "prb_sparse_entropy_host": [_vp, _i64, _i64, _i64, _vp],
}
_RESTYPES = {"prb_last_error": ctypes.c_char_p, "prb_finalize_update_blocks": _i64,
             "prb_wide_terms_stride": _i64, "prb_wide_chain_blocks": _i64}

EXPORTED = tuple(_SIGS)


def lib_path():
    return _build.LIB


def load():
    """Load libprb200.so, building it in-tree with nvcc if it is absent/stale."""
    global _LIB
    if _LIB is not None:
        return _LIB
    try:
        path = _build.build()
    except Exception as e:  # no nvcc and no prebuilt library: fail loudly
        if os.path.exists(_build.LIB):
            path = _build.LIB
        else:
            raise ImportError(
                "picturedrocks_b200 needs its CUDA library libprb200.so (no CPU fallback exists): "
                + str(e))
    lib = ctypes.CDLL(path)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, _int)
    if lib.prb_abi_version() != 3:
        raise ImportError("libprb200.so ABI version mismatch")
    _LIB = lib
    return lib


class PrbError(RuntimeError):
    pass


# kernels launched per entry point (for bench.py's gpu_launches claim)
KERNELS = {
    "prb_pack_csc": 1, "prb_unpack_csc": 1, "prb_pack_chunk": 1, "prb_check_columns": 1, "prb_csr_count": 1, "prb_csr_fill": 2, "prb_dense_count_nnz": 1, "prb_dense_fill_csc": 1,
    "prb_tile_ptrs": 1, "prb_stream_hits": 1,
    "prb_argmax_commit": 1, "prb_pick_scatter": 1, "prb_peer_share": 1, "prb_finalize_update": 1,
    "prb_finalize_update_wide": 2,
    "prb_v_unscatter": 1, "prb_best_record": 1, "prb_v_scatter_flat": 1, "prb_ell_build": 1,
    "prb_stream_ell": 1, "prb_stream_ell2": 1, "prb_v_count": 1, "prb_v_scatter": 1,
    "prb_stream_v": 1, "prb_scatter_gene_u8": 1,
    "prb_scatter_gene_or32": 1, "prb_state_direct": 2, "prb_state_table": 2, "prb_state_assign": 1, "prb_finalize": 1, "prb_finalize_direct": 1,
    "prb_entropy_keys": 2, "prb_mi_base": 1, "prb_selector_update": 1, "prb_argmax_blocks": 1,
    "prb_argmax_final": 1, "prb_commit_add": 1, "prb_transpose": 1, "prb_sort_columns": 1,
    "prb_qd_edges": 1, "prb_qd_digitize": 1, "prb_sort_segments": 1, "prb_qd_edges_sparse": 1,
    "prb_qd_count_sparse": 1, "prb_qd_fill_sparse": 1, "prb_synth_labels": 1, "prb_synth_count": 1,
    "prb_synth_fill": 1,
    "prb_v_scatter16": 1, "prb_stream_v16": 1,
}
LAUNCHES = 0


def call(name, *args):
    global LAUNCHES
    lib = load()
    LAUNCHES += KERNELS.get(name, 0)
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise PrbError("%s failed: %s" % (name, lib.prb_last_error().decode()))


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array; None -> NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def stream():
    import torch

    return torch.cuda.current_stream().cuda_stream


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError(
            "picturedrocks_b200 runs on a CUDA device (B200, sm_100a) only; there is no CPU path")
