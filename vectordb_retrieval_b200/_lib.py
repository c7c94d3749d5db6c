"""ctypes binding of libvdbcuda.so (the C ABI declared in include/vdb_cuda.h).

There is no CPU fallback: if the shared library is missing this module raises at first use,
and every call that returns non-zero raises RuntimeError with the library's message (device
failures must surface as RuntimeError - the reference harness swallows ValueError/TypeError
from batch_search and silently degrades, src/experiments/experiment_runner.py:442-455)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvdbcuda.so")

METRIC_L2, METRIC_IP = 0, 1
OUT_SQRT, OUT_NEGATE, OUT_ONE_MINUS = 1, 2, 4
IMPL_AUTO, IMPL_TCGEN05, IMPL_TCGEN05_1CTA, IMPL_SIMT = 0, 1, 2, 3
IMPL_NAMES = {"auto": IMPL_AUTO, "tcgen05": IMPL_TCGEN05, "tcgen05_1cta": IMPL_TCGEN05_1CTA, "simt": IMPL_SIMT}

_p, _i64, _i32, _f32, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/vdb_cuda.h one to one
SIGNATURES = {
    "vdb_last_error": (C.c_char_p, []),
    "vdb_abi_version": (_i32, []),
    "vdb_launch_count": (_i64, []),
    "vdb_flat_timing_enable": (_i32, [_i32]),
    "vdb_flat_timing_read": (_i32, [C.POINTER(C.c_float), _i32, C.POINTER(C.c_int)]),
    "vdb_sm_count": (_i32, [C.POINTER(C.c_int)]),
    "vdb_row_norms": (_i32, [_p, _i64, _i32, _i64, _p, _p]),
    "vdb_normalize_rows": (_i32, [_p, _i64, _i32, _i64, _p, _i64, _p]),
    "vdb_flat_kpad": (_i32, [_i32]),
    "vdb_flat_npad": (_i64, [_i64]),
    "vdb_flat_nqpad": (_i64, [_i64]),
    "vdb_flat_prepare": (_i32, [_p, _i64, _i32, _i64, _i32, _p, _p, _p, _p]),
    "vdb_flat_prepare_queries": (_i32, [_p, _i64, _i32, _i64, _p, _p, _p]),
    "vdb_flat_topk_workspace_bytes": (_sz, [_i64, _i32]),
    "vdb_flat_topk": (_i32, [_i32, _p, _p, _p, _i64, _i32, _i64, _p, _p, _i64, _i32, _i32, _f32, _i32,
                             _p, _p, _p, _sz, _p]),
    "vdb_flat_grid_clusters": (_i32, [_i32, _i32, C.POINTER(C.c_int)]),
    "vdb_flat_dense_keys": (_i32, [_p, _p, _p, _i64, _i32, _p, _p, _i64, _i32, _p, _p]),
    "vdb_set_debug_mode": (_i32, [_i32]),
    "vdb_debug_read_prof": (_i32, [C.POINTER(C.c_uint64)]),
    "vdb_flat_set_seeding": (_i32, [_i32, _i32]),
    "vdb_flat_set_seeding_margin": (_i32, [_i32]),
    "vdb_debug_redo_queries": (_i32, [C.POINTER(C.c_uint64)]),
    "vdb_merge_topk": (_i32, [_p, _p, _i32, _i64, _i32, _i32, _f32, _p, _p, _p]),
    "vdb_merge_topk_strided": (_i32, [_p, _p, _i64, _i64, _i32, _i64, _i32, _i32, _f32, _p, _p, _p]),
    "vdb_rerank_topk": (_i32, [_i32, _p, _i64, _i32, _i64, _p, _i64, _i32, _p, _i64, _i32, _i32, _f32, _p, _p, _p]),
    "vdb_lsh_code_words": (_i32, [_i32]),
    "vdb_lsh_encode": (_i32, [_p, _i64, _i32, _i64, _p, _i32, _p, _p]),
    "vdb_hamming_topk_workspace_bytes": (_sz, [_i64, _i32]),
    "vdb_hamming_topk": (_i32, [_p, _i64, _p, _i64, _i32, _i32, _i64, _p, _p, _p, _sz, _p]),
    "vdb_hamming_tc_row_bytes": (_i32, [_i32]),
    "vdb_hamming_tc_expand": (_i32, [_p, _i64, _i32, _i32, _i32, _p, _p, _i64, _p]),
    "vdb_hamming_tc_workspace_bytes": (_sz, [_i64, _i32, _i32, _i64]),
    "vdb_hamming_topk_tc": (_i32, [_p, _p, _p, _i64, _p, _p, _i64, _i32, _i32, _i64, _p, _p, _p, _sz, _p]),
    "vdb_hamming_topk_tc_f16": (_i32, [_p, _p, _p, _i64, _p, _p, _i64, _i32, _i32, _i64, _p, _p, _p, _sz, _p]),
    "vdb_lsh_candidates_workspace_bytes": (_sz, [_i64, _i64]),
    "vdb_lsh_candidates": (_i32, [_p, _i64, _p, _p, _p, _i64, _i32, _i64, _i64, _i32, _p, _p, _p, _sz, _p]),
    "vdb_ivf_d4": (_i32, [_i32]),
    "vdb_ivf_count": (_i32, [_p, _i64, _i32, _p, _p]),
    "vdb_kmeans_accumulate": (_i32, [_p, _i64, _i32, _i64, _p, _i32, _p, _p, _p]),
    "vdb_ivf_fill": (_i32, [_p, _i64, _i32, _i64, _p, _p, _i32, _p, _p, _p, _p]),
    "vdb_ivf_scan_topk": (_i32, [_i32, _p, _p, _p, _i32, _i32, _p, _i32, _p, _i64, _i64, _i32, _i32, _f32, _i64,
                                 _p, _p, _p, _p]),
    "vdb_ivf_scan_topk_ex": (_i32, [_i32, _p, _p, _p, _i32, _i32, _p, _i32, _p, _i64, _i64, _i32, _i32, _f32, _i64,
                                    _p, _p, _p, _i64, _p]),
    "vdb_sq8_d16": (_i32, [_i32]),
    "vdb_sq8_residuals": (_i32, [_p, _i64, _i32, _i64, _p, _p, _p, _p]),
    "vdb_sq8_train": (_i32, [_p, _i64, _i32, _i64, _p, _p, _p, _p]),
    "vdb_sq8_fill": (_i32, [_p, _i64, _i32, _i64, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "vdb_ivf_sq8_scan_topk": (_i32, [_i32, _p, _p, _p, _i32, _i32, _p, _p, _p, _p, _i32, _p, _i64, _i64, _i32, _i32, _f32, _i64,
                                     _p, _p, _i64, _p]),
    "vdb_pq_encode": (_i32, [_p, _i64, _i32, _i64, _p, _i32, _p, _p]),
    "vdb_pq_bias": (_i32, [_p, _i64, _i32, _i32, _p, _p, _p, _p, _p]),
    "vdb_pq_decode": (_i32, [_p, _i64, _i32, _i32, _p, _p, _p, _p, _i64, _p]),
    "vdb_bytes_fill": (_i32, [_p, _i64, _i32, _p, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "vdb_ivf_pq_scan_topk": (_i32, [_i32, _p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _p, _i32, _p, _i64, _i64, _i32, _i32, _f32, _i64,
                                    _p, _p, _p]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the library (once) and bind every symbol of the header; fail loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m vectordb_retrieval_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    if lib.vdb_abi_version() != 1:
        raise RuntimeError("libvdbcuda.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().vdb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()
