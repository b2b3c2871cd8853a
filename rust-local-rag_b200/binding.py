"""ctypes binding of include/rlr_b200.h -- the same symbols the reference's `-sys` crate
would bind.  Nothing here computes; it only marshals pointers."""
import ctypes as C
import os

import numpy as np

from . import _build

RLR_OK = 0
ERR_NAMES = {
    1: "RLR_ERR_INVALID_ARG", 2: "RLR_ERR_NO_DEVICE", 3: "RLR_ERR_CUDA", 4: "RLR_ERR_OOM",
    5: "RLR_ERR_DIM_MISMATCH", 6: "RLR_ERR_UNSUPPORTED", 7: "RLR_ERR_NONFINITE",
}
RLR_ERR_INVALID_ARG, RLR_ERR_NO_DEVICE, RLR_ERR_CUDA, RLR_ERR_OOM = 1, 2, 3, 4
RLR_ERR_DIM_MISMATCH, RLR_ERR_UNSUPPORTED, RLR_ERR_NONFINITE = 5, 6, 7

RLR_MAX_TOP_K = 100
RLR_MAX_M = 1024
RLR_MAX_DIM = 4096
RLR_STORE_KEEP_F16 = 0x1
RLR_STORE_CHECK_FINITE = 0x2
RLR_STORE_F16_ONLY = 0x4
RLR_STORE_NORMALIZE_ON_UPLOAD = 0x8
RLR_QUERY_PRENORMALIZED = 0x1
RLR_WANT_TIMINGS = 0x2
RLR_SEARCH_F16 = 0x4
RLR_BATCH_EXACT_RESCORE = 0x8
RLR_BATCH_F16 = 0x10
RLR_BATCH_BF16 = 0x20
RLR_BATCH_TF32 = 0x40
RLR_STORE_KEEP_BF16 = 0x10
RLR_STORE_NO_LATENCY_PATH = 0x20
RLR_IPC_HANDLE_BYTES = 64
RLR_SYNTH_IID = 0
RLR_SYNTH_CLUSTERED = 1
RLR_MAX_SHARDS = 16
RLR_MAX_MULTI = 3


class RlrError(RuntimeError):
    """A non-zero status from the C ABI (maps to anyhow::Error in the reference glue)."""

    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class QueryWeightsC(C.Structure):
    _fields_ = [("embedding", C.c_float), ("lexical", C.c_float), ("reranker", C.c_float),
                ("initial", C.c_float), ("has", C.c_uint32)]


class ResolvedWeightsC(C.Structure):
    _fields_ = [("embedding", C.c_float), ("lexical", C.c_float), ("reranker", C.c_float),
                ("initial", C.c_float)]


class StoreInfoC(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("row_base", C.c_uint64), ("dim", C.c_uint32), ("pitch", C.c_uint32),
                ("device", C.c_int32), ("flags", C.c_uint32), ("bytes_device", C.c_uint64)]


class DeviceInfoC(C.Structure):
    _fields_ = [("device", C.c_int32), ("sm_count", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
                ("total_mem", C.c_uint64), ("name", C.c_char * 128)]


class ClusterInfoC(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("dim", C.c_uint32), ("pitch", C.c_uint32), ("flags", C.c_uint32),
                ("n_shards", C.c_uint32), ("device", C.c_int32 * 16), ("row_base", C.c_uint64 * 16),
                ("shard_rows", C.c_uint64 * 16)]


class TimingsC(C.Structure):
    _fields_ = [("scan_ms", C.c_float), ("merge_ms", C.c_float), ("mmr_ms", C.c_float), ("total_ms", C.c_float),
                ("launches", C.c_uint32)]


CAND_DTYPE = np.dtype([("key", "<u8"), ("emb", "<f4"), ("lex", "<f4")])  # rlr_cand, 16 bytes

_vp, _u32, _u64, _f32, _int = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float, C.c_int
_pf, _pu32 = C.POINTER(C.c_float), C.POINTER(C.c_uint32)

# name -> (restype, argtypes): every symbol include/rlr_b200.h declares
PROTOTYPES = {
    "rlr_abi_version": (_int, []),
    "rlr_last_error": (C.c_char_p, []),
    "rlr_device_count": (_int, [C.POINTER(C.c_int)]),
    "rlr_device_query": (_int, [_int, C.POINTER(DeviceInfoC)]),
    "rlr_normalize": (_int, [_pf, C.c_size_t]),
    "rlr_resolve_weights": (_int, [C.POINTER(QueryWeightsC), C.POINTER(ResolvedWeightsC)]),
    "rlr_store_create": (_int, [_int, _u32, _u64, _vp, _u64, _u64, _u32, C.POINTER(_vp)]),
    "rlr_store_destroy": (_int, [_vp]),
    "rlr_store_info_get": (_int, [_vp, C.POINTER(StoreInfoC)]),
    "rlr_store_upload": (_int, [_vp, _u64, _u64, _vp, _u64]),
    "rlr_store_reserve": (_int, [_vp, _u64]),
    "rlr_store_append": (_int, [_vp, _u64, _vp, _u64, C.POINTER(_u64)]),
    "rlr_store_remove_rows": (_int, [_vp, _vp, _u64, _vp, _vp, C.POINTER(_u64)]),
    "rlr_store_read_rows": (_int, [_vp, _vp, _u64, _vp]),
    "rlr_store_fill_synthetic": (_int, [_vp, _int, _u64, _u64, _u32, _f32]),
    "rlr_search_topm": (_int, [_vp, _vp, _u32, _u32, C.POINTER(ResolvedWeightsC), _vp, _vp, _u32, _u32,
                               _vp, _vp, _vp, _vp, _pu32]),
    "rlr_mmr": (_int, [_vp, _vp, _vp, _u32, _u32, _f32, _u32, _vp, _pu32]),
    "rlr_search_mmr": (_int, [_vp, _vp, _u32, _u32, _u32, _f32, C.POINTER(ResolvedWeightsC), _vp, _vp, _u32,
                              _vp, _vp, _vp, _vp, _pu32]),
    "rlr_embedding_candidates": (_int, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _pu32]),
    "rlr_search_mmr_multi": (_int, [_vp, _vp, _u32, _u32, _u32, _u32, _f32, C.POINTER(ResolvedWeightsC), _vp, _vp, _vp,
                                    _vp, _vp, _vp, _vp, _vp]),
    "rlr_search_mmr_multi_async": (_int, [_vp, _u32, _vp, _u32, _f32, _f32, _f32, _vp, _vp, _vp]),
    "rlr_search_batch": (_int, [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp, _vp]),
    "rlr_search_batch_device": (_int, [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp, _vp]),
    "rlr_batch_merge_async": (_int, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp]),
    "rlr_bm25_create": (_int, [_vp, C.POINTER(_vp)]),
    "rlr_bm25_destroy": (_int, [_vp]),
    "rlr_bm25_set_doc": (_int, [_vp, _u32, _vp, _vp, _u32]),
    "rlr_bm25_set_docs": (_int, [_vp, _u32, _u32, _vp, _vp, _vp]),
    "rlr_bm25_remove_doc": (_int, [_vp, _u32]),
    "rlr_bm25_move_doc": (_int, [_vp, _u32, _u32]),
    "rlr_bm25_stats": (_int, [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    "rlr_bm25_score": (_int, [_vp, _vp, _u32, _u32, _vp, _vp, _u32, _pu32]),
    "rlr_search_text_topm": (_int, [_vp, _vp, _vp, _u32, _u32, C.POINTER(ResolvedWeightsC), _vp, _u32, _u32, _vp, _vp, _vp, _vp, _pu32]),
    "rlr_search_text_mmr": (_int, [_vp, _vp, _vp, _u32, _u32, _u32, _f32, C.POINTER(ResolvedWeightsC), _vp, _u32, _vp, _vp, _vp, _vp, _pu32]),
    "rlr_cluster_bm25_create": (_int, [_vp, C.POINTER(_vp)]),
    "rlr_cluster_bm25_destroy": (_int, [_vp]),
    "rlr_cluster_bm25_set_doc": (_int, [_vp, _u32, _vp, _vp, _u32]),
    "rlr_cluster_bm25_set_docs": (_int, [_vp, _u32, _u32, _vp, _vp, _vp]),
    "rlr_cluster_bm25_remove_doc": (_int, [_vp, _u32]),
    "rlr_cluster_bm25_stats": (_int, [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    "rlr_cluster_bm25_score": (_int, [_vp, _vp, _u32, _u32, _vp, _vp, _u32, _pu32]),
    "rlr_cluster_search_text_topm": (_int, [_vp, _vp, _vp, _u32, _u32, C.POINTER(ResolvedWeightsC), _vp, _u32, _u32, _vp, _vp, _vp, _vp, _pu32]),
    "rlr_cluster_search_text_mmr": (_int, [_vp, _vp, _vp, _u32, _u32, _u32, _f32, C.POINTER(ResolvedWeightsC), _vp, _u32, _vp, _vp, _vp, _vp, _pu32]),
    "rlr_last_timings": (_int, [C.POINTER(TimingsC)]),
    "rlr_cluster_create": (_int, [_vp, _u32, _u32, _u64, _vp, _u64, _u32, _vp, C.POINTER(_vp)]),
    "rlr_cluster_destroy": (_int, [_vp]),
    "rlr_cluster_info_get": (_int, [_vp, C.POINTER(ClusterInfoC)]),
    "rlr_cluster_upload": (_int, [_vp, _u64, _u64, _vp, _u64]),
    "rlr_cluster_read_rows": (_int, [_vp, _vp, _u64, _vp]),
    "rlr_cluster_fill_synthetic": (_int, [_vp, _int, _u64, _u64, _u32, _f32]),
    "rlr_cluster_search_topm": (_int, [_vp, _vp, _u32, _u32, C.POINTER(ResolvedWeightsC), _vp, _vp, _u32, _u32,
                                       _vp, _vp, _vp, _vp, _pu32]),
    "rlr_cluster_mmr": (_int, [_vp, _vp, _vp, _u32, _u32, _f32, _u32, _vp, _pu32]),
    "rlr_cluster_search_mmr": (_int, [_vp, _vp, _u32, _u32, _u32, _f32, C.POINTER(ResolvedWeightsC), _vp, _vp, _u32,
                                      _vp, _vp, _vp, _vp, _pu32]),
    "rlr_cluster_search_mmr_multi": (_int, [_vp, _vp, _u32, _u32, _u32, _u32, _f32, C.POINTER(ResolvedWeightsC), _vp, _vp, _vp,
                                     _vp, _vp, _vp, _vp, _vp]),
    "rlr_cluster_search_batch": (_int, [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp, _vp]),
    "rlr_cluster_embedding_candidates": (_int, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _pu32]),
    "rlr_cluster_last_scan_ms": (_int, [_vp, _u32, _pu32]),
    "rlr_cluster_launch_count": (_int, [_vp, C.POINTER(_u64)]),
    "rlr_ctx_create": (_int, [_vp, C.POINTER(_vp)]),
    "rlr_ctx_destroy": (_int, [_vp]),
    "rlr_topm_async": (_int, [_vp, _vp, _f32, _f32, _vp, _vp, _u32, _u32, _vp, _vp, _vp]),
    "rlr_merge_async": (_int, [_vp, _vp, _u32, _u32, _vp, _vp, _vp]),
    "rlr_gather_async": (_int, [_vp, _vp, _vp, _u32, _vp, _vp]),
    "rlr_mmr_async": (_int, [_vp, _vp, _u32, _u32, _vp, _vp, _u32, _u32, _f32, _vp, _vp, _vp, _vp]),
    "rlr_mmr_store_async": (_int, [_vp, _vp, _vp, _u32, _u32, _f32, _vp, _vp, _vp, _vp]),
    "rlr_store_ipc_export": (_int, [_vp, _u32, _vp]),
    "rlr_peer_set_open": (_int, [_vp, _u32, _u32, _vp, _vp, _vp, _u32, C.POINTER(_vp)]),
    "rlr_peer_set_close": (_int, [_vp]),
    "rlr_mmr_peers_async": (_int, [_vp, _vp, _vp, _vp, _u32, _u32, _f32, _vp, _vp, _vp, _vp]),
    "rlr_mailbox_create": (_int, [_int, _u32, _u32, _u32, C.POINTER(_vp)]),
    "rlr_mailbox_ipc_export": (_int, [_vp, _vp]),
    "rlr_mailbox_open": (_int, [_int, _vp, _u32, _u32, _u32, C.POINTER(_vp)]),
    "rlr_mailbox_close": (_int, [_vp]),
    "rlr_mailbox_status": (_int, [_vp, _pu32]),
    "rlr_topm_post_async": (_int, [_vp, _vp, _u32, _u64, _vp, _f32, _f32, _vp, _vp, _u32, _u32, _vp]),
    "rlr_mailbox_merge_async": (_int, [_vp, _vp, _u64, _u32, _vp, _vp, _vp]),
    "rlr_search_mmr_async": (_int, [_vp, _vp, _u32, _f32, _f32, _f32, _vp, _vp, _vp]),
    "rlr_ctx_set_flags": (_int, [_vp, _u32]),
    "rlr_ctx_launch_count": (_int, [_vp, C.POINTER(_u64)]),
    "rlr_time_scan": (_int, [_vp, _vp, _u32, _u32, _vp, _pf]),
}

# librlr_hostmirror.so (include/rlr_hostmirror.h): host-mirror support, not the product boundary
HOSTMIRROR_PROTOTYPES = {
    "rlr_hostmirror_last_error": (C.c_char_p, []),
    "rlr_lexical_create": (_int, [C.POINTER(_vp)]),
    "rlr_lexical_destroy": (_int, [_vp]),
    "rlr_lexical_add_chunk": (_int, [_vp, _u64, C.c_char_p, C.c_size_t]),
    "rlr_lexical_remove_chunk": (_int, [_vp, _u64]),
    "rlr_lexical_contains": (_int, [_vp, _u64, C.POINTER(C.c_int)]),
    "rlr_lexical_stats": (_int, [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    "rlr_lexical_score": (_int, [_vp, C.c_char_p, C.c_size_t, _u32, _vp, _vp, _u32, _pu32]),
    "rlr_tokenize": (_int, [C.c_char_p, C.c_size_t, _vp, C.c_size_t, C.POINTER(C.c_size_t), _pu32]),
    "rlr_hostmirror_unicode_dump": (_int, [_vp, _vp, _u32]),
}

_lib = None
_hm = None


def load_hostmirror(build: bool = True) -> C.CDLL:
    """dlopen librlr_hostmirror.so: the BM25 / tokenizer twin for the host mirrors' text queries (plain C++)."""
    global _hm
    if _hm is not None:
        return _hm
    path = _build.build_hostmirror() if build else _build.HM_LIB
    lib = C.CDLL(path)
    for name, (res, args) in HOSTMIRROR_PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _hm = lib
    return lib


def check_hm(rc: int) -> None:
    if rc != 0:
        msg = load_hostmirror().rlr_hostmirror_last_error()
        raise RlrError(rc, msg.decode("utf-8", "replace") if msg else "")


def load(build: bool = True) -> C.CDLL:
    """dlopen librlr_b200.so (building it first if it is missing or stale).  Fails loudly:
    there is no Python/CPU fallback for any compute entry point."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build:
        path = _build.build()
    if not os.path.exists(path):
        raise RlrError(RLR_ERR_NO_DEVICE, f"{path} is missing: run __graft_entry__.build()")
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != RLR_OK:
        msg = load().rlr_last_error()
        raise RlrError(rc, msg.decode("utf-8", "replace") if msg else "")


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def key_row(keys: np.ndarray) -> np.ndarray:
    return (~(keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)).astype(np.uint32)


def key_score(keys: np.ndarray) -> np.ndarray:
    o = (keys >> np.uint64(32)).astype(np.uint32)
    bits = np.where(o & np.uint32(0x80000000), o & np.uint32(0x7FFFFFFF), ~o).astype(np.uint32)
    return bits.view(np.float32)
