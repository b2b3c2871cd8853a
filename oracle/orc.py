"""ctypes loader for the CPU oracle (oracle/rlr_oracle.c).  TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs -- never from the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liborc.so")

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "rlr_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(src) > os.path.getmtime(LIB):
        res = subprocess.run(["make", "-C", HERE, "-B", "liborc.so"], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return LIB


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        f32p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        lib.orc_dot.restype = C.c_float
        lib.orc_dot.argtypes = [f32p, C.c_size_t, f32p, C.c_size_t]
        lib.orc_normalize.restype = None
        lib.orc_normalize.argtypes = [f32p, C.c_size_t]
        lib.orc_cosine.restype = C.c_float
        lib.orc_cosine.argtypes = [f32p, C.c_size_t, f32p, C.c_size_t]
        lib.orc_resolve_weight.restype = C.c_float
        lib.orc_resolve_weight.argtypes = [C.c_int, C.c_float, C.c_float]
        lib.orc_default_weights.restype = None
        lib.orc_default_weights.argtypes = [f32p]
        lib.orc_search.restype = C.c_int64
        lib.orc_search.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_uint64, f32p, C.c_int, C.c_uint64, C.c_float,
                                   C.c_float, u32p, f32p, C.c_uint64, C.c_int, C.c_int, u32p, f32p, f32p, f32p]
        lib.orc_embedding_candidates.restype = C.c_int64
        lib.orc_embedding_candidates.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_uint64, f32p, C.c_int, C.c_uint64,
                                                 C.c_int, u32p, f32p]
        lib.orc_mmr.restype = C.c_int64
        lib.orc_mmr.argtypes = [f32p, C.c_uint64, f32p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_float, C.c_int, u32p]
        lib.orc_search_with_diversity.restype = C.c_int64
        lib.orc_search_with_diversity.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_uint64, f32p, C.c_int, C.c_uint64,
                                                  C.c_float, C.c_float, C.c_float, u32p, f32p, C.c_uint64, C.c_int,
                                                  C.c_int, u32p, f32p, f32p, f32p]
        lib.orc_synth_rows.restype = None
        lib.orc_synth_rows.argtypes = [f32p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_uint64,
                                       C.c_uint64, C.c_uint32, C.c_float, C.c_int]
        lib.orc_max_threads.restype = C.c_int
        lib.orc_max_threads.argtypes = []
        _lib = lib
    return _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def dot(a, b) -> float:
    a, b = f32(a), f32(b)
    return float(load().orc_dot(_f(a), len(a), _f(b), len(b)))


def normalize(v) -> np.ndarray:
    v = np.array(v, dtype=np.float32, order="C")
    load().orc_normalize(_f(v), len(v))
    return v


def normalize_rows(rows: np.ndarray) -> np.ndarray:
    rows = np.array(rows, dtype=np.float32, order="C")
    lib = load()
    for i in range(rows.shape[0]):
        lib.orc_normalize(_f(rows[i]), rows.shape[1])
    return rows


def cosine(a, b) -> float:
    a, b = f32(a), f32(b)
    return float(load().orc_cosine(_f(a), len(a), _f(b), len(b)))


def resolve_weight(override, default: float) -> float:
    has = override is not None
    return float(load().orc_resolve_weight(int(has), float(override) if has else 0.0, float(default)))


def max_threads() -> int:
    return int(load().orc_max_threads())


def _lex(lex_rows, lex_scores):
    if lex_rows is None or len(lex_rows) == 0:
        return None, None, 0
    lr = np.ascontiguousarray(lex_rows, dtype=np.uint32)
    ls = np.ascontiguousarray(lex_scores, dtype=np.float32)
    return lr, ls, len(lr)


def search(rows, q, top_k, w_embed=0.7, w_lex=0.3, lex_rows=None, lex_scores=None, normalize_query=True,
           full_sort=False, threads=1):
    rows = f32(rows); q = f32(q)
    n, pitch = rows.shape
    cap = max(int(top_k), 1)
    o_rows = np.zeros(cap, np.uint32); o_score = np.zeros(cap, np.float32)
    o_emb = np.zeros(cap, np.float32); o_lex = np.zeros(cap, np.float32)
    lr, ls, nl = _lex(lex_rows, lex_scores)
    k = load().orc_search(_f(rows), n, len(q), pitch, _f(q), int(normalize_query), int(top_k), w_embed, w_lex,
                          _u(lr) if nl else None, _f(ls) if nl else None, nl, int(full_sort), threads,
                          _u(o_rows), _f(o_score), _f(o_emb), _f(o_lex))
    return o_rows[:k], o_score[:k], o_emb[:k], o_lex[:k]


def embedding_candidates(rows, q, count, normalize_query=True, threads=1):
    rows = f32(rows); q = f32(q)
    n, pitch = rows.shape
    cap = max(int(count), 1)
    o_rows = np.zeros(cap, np.uint32); o_score = np.zeros(cap, np.float32)
    k = load().orc_embedding_candidates(_f(rows), n, len(q), pitch, _f(q), int(normalize_query), int(count), threads,
                                        _u(o_rows), _f(o_score))
    return o_rows[:k], o_score[:k]


def mmr(emb, relevance, top_k, lam, threads=1):
    emb = f32(emb); rel = f32(relevance)
    p = emb.shape[0]
    if p == 0:
        return np.zeros(0, np.uint32)
    out = np.zeros(p, np.uint32)
    k = load().orc_mmr(_f(emb), emb.shape[1], _f(rel), p, emb.shape[1], int(top_k), float(lam), threads, _u(out))
    return out[:k]


def search_with_diversity(rows, q, top_k, lam, w_embed=0.7, w_lex=0.3, lex_rows=None, lex_scores=None,
                          normalize_query=True, full_sort=False, threads=1):
    rows = f32(rows); q = f32(q)
    n, pitch = rows.shape
    cap = max(int(top_k), 1)
    o_rows = np.zeros(cap, np.uint32); o_score = np.zeros(cap, np.float32)
    o_emb = np.zeros(cap, np.float32); o_lex = np.zeros(cap, np.float32)
    lr, ls, nl = _lex(lex_rows, lex_scores)
    k = load().orc_search_with_diversity(_f(rows), n, len(q), pitch, _f(q), int(normalize_query), int(top_k),
                                         float(lam), w_embed, w_lex, _u(lr) if nl else None, _f(ls) if nl else None,
                                         nl, int(full_sort), threads, _u(o_rows), _f(o_score), _f(o_emb), _f(o_lex))
    return o_rows[:k], o_score[:k], o_emb[:k], o_lex[:k]


def synth_rows(n, dim, kind=0, seed=0x5EED0001, centroid_seed=0x5EED00C0, n_clusters=4096, sigma=0.65, row0=0,
               threads=0):
    out = np.zeros((n, dim), np.float32)
    if threads <= 0:
        threads = max_threads()
    load().orc_synth_rows(_f(out), dim, row0, n, dim, kind, seed, centroid_seed, n_clusters, sigma, threads)
    return out
