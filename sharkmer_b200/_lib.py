"""ctypes loader for libsharkmer_b200.so — the C ABI declared in include/sharkmer_b200.h.

Fails loudly: if the CUDA library is missing or cannot be loaded there is nothing
to fall back to (the engine has no CPU path).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsharkmer_b200.so")

OK = 0
ERR_INVALID_ARG, ERR_INVALID_BASE, ERR_CUDA, ERR_OOM, ERR_STATE, ERR_CONSERVATION, ERR_K_MISMATCH, ERR_NO_READS, ERR_CAPACITY = range(1, 10)
INSERT_AUTO, INSERT_DIRECT, INSERT_PARTITIONED = 0, 1, 2
LOOKUP_CANONICAL, LOOKUP_EXACT, LOOKUP_EITHER = 0, 1, 2
INGEST_ASYNC = 1


class SkmParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("k", C.c_uint32), ("chunks", C.c_uint32), ("insert_mode", C.c_uint32),
        ("histo_max", C.c_uint64), ("capacity_hint", C.c_uint64),
        ("device", C.c_int32), ("n_ranks", C.c_uint32), ("rank", C.c_uint32), ("reserved", C.c_uint32),
        ("stream", C.c_uint64),
    ]


class SkmTotals(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("n_reads", "n_bases", "n_bases_read", "n_kmers", "n_unique", "n_singletons", "n_saturated")]


class SkmStageMs(C.Structure):
    _fields_ = [(n, C.c_float) for n in
                ("h2d", "pack", "count", "partition", "insert", "histogram", "grow", "total_finalize")] + [
        ("launches", C.c_uint32 * 8),
        ("kernel_launches", C.c_uint32), ("n_grows", C.c_uint32),
        ("table_capacity", C.c_uint64), ("table_bytes", C.c_uint64),
        ("insert_kmers", C.c_uint64), ("insert_bases", C.c_uint64),
        ("sort", C.c_float), ("scan", C.c_float), ("sort_launches", C.c_uint32), ("scan_launches", C.c_uint32),
        ("tiled_launches", C.c_uint32), ("tiled_retries", C.c_uint32)]


# every symbol include/sharkmer_b200.h declares: name -> (restype, argtypes)
_vp, _u8p, _u32, _u64, _i32 = C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
SYMBOLS = {
    "skm_abi_version": (_u32, []),
    "skm_create": (_i32, [C.POINTER(SkmParams), C.POINTER(_vp)]),
    "skm_destroy": (None, [_vp]),
    "skm_last_error": (C.c_char_p, [_vp]),
    "skm_pinned_alloc": (_i32, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "skm_pinned_free": (_i32, [_vp, _vp]),
    "skm_ingest_batch": (_i32, [_vp, _u32, _u8p, _u64, _u32]),
    "skm_ingest_reads": (_i32, [_vp, _u32, _u8p, _vp, _u64]),
    "skm_ingest_device": (_i32, [_vp, _u32, _u8p, _u64]),
    "skm_sync": (_i32, [_vp]),
    "skm_finalize": (_i32, [_vp]),
    "skm_reset": (_i32, [_vp]),
    "skm_histogram": (_i32, [_vp, _u32, _vp, _u64]),
    "skm_totals_get": (_i32, [_vp, C.POINTER(SkmTotals)]),
    "skm_chunk_totals": (_i32, [_vp, _u32, C.POINTER(SkmTotals)]),
    "skm_stage_times": (_i32, [_vp, C.POINTER(SkmStageMs)]),
    "skm_table_len": (_i32, [_vp, C.POINTER(_u64)]),
    "skm_export": (_i32, [_vp, _vp, _vp, _u64, _i32, C.POINTER(_u64)]),
    "skm_table_digest": (_i32, [_vp, C.POINTER(_u64)]),
    "skm_lookup_batch": (_i32, [_vp, _vp, _u64, _u32, _i32, _vp, _vp]),
    "skm_scan_oligos": (_i32, [_vp, _vp, _u64, _u32, _u32, _vp, _vp, _u64, C.POINTER(_u64)]),
    "skm_insert_counts": (_i32, [_vp, _vp, _vp, _u64]),
    "skm_mg_arena_create": (_i32, [_vp, _u64]),
    "skm_mg_arena_handle": (_i32, [_vp, _vp]),
    "skm_mg_arena_ptr": (_i32, [_vp, C.POINTER(_vp)]),
    "skm_mg_open_peer": (_i32, [_vp, _u32, _vp]),
    "skm_mg_set_peer": (_i32, [_vp, _u32, _vp, _i32]),
    "skm_mg_finalize": (_i32, [_vp, _vp]),
    "skm_mg_flush": (_i32, [_vp, _vp]),
    "skm_mg_bytes_sent": (_i32, [_vp, C.POINTER(_u64)]),
    "skm_group_create": (_i32, [C.POINTER(SkmParams), _u32, _vp, _u64, C.POINTER(_vp)]),
    "skm_group_ctx": (_vp, [_vp, _u32]),
    "skm_group_finalize": (_i32, [_vp]),
    "skm_group_flush": (_i32, [_vp]),
    "skm_group_reset": (_i32, [_vp]),
    "skm_group_last_error": (C.c_char_p, [_vp]),
    "skm_group_destroy": (None, [_vp]),
    "skm_insert_kmers_device": (_i32, [_vp, _vp, _u64]),
    "skm_snapshot_histogram": (_i32, [_vp, _u32]),
    "skm_extract_kmers": (_i32, [_vp, _u8p, _u64, _vp]),
    "skm_pack": (_i32, [_vp, _u8p, _u64, _vp, _vp]),
    "skm_synth_device": (_i32, [_vp, _u64, _u64, _u32, _u32, _u32, _u32, _u32, _u64, _u64, _vp]),
    "skm_device_alloc": (_i32, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "skm_device_free": (_i32, [_vp, _vp]),
    "skm_memcpy_d2h": (_i32, [_vp, _vp, _vp, C.c_size_t]),
    "skm_memcpy_h2d": (_i32, [_vp, _vp, _vp, C.c_size_t]),
    "skm_bench_gups": (_i32, [_vp, _u32, _u64, _u32, _i32, C.POINTER(C.c_float)]),
}

_lib = None


def load():
    """Load the CUDA library.  Raises if it is missing — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m sharkmer_b200.build` "
            "(nvcc, sm_100a).  sharkmer_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if L.skm_abi_version() != 1:
        raise RuntimeError("libsharkmer_b200.so ABI version mismatch")
    _lib = L
    return L
