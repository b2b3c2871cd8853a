"""Host-side mirror of the reference's `RagEngine` for the retrieval path, above the C ABI.

Same method names, argument meaning and error behaviour as
/root/reference/src/rag_engine.rs: `search` (:470), `search_with_diversity` (:717),
`get_embedding_candidates` (:415), `QueryWeights` (:1846), `SearchResult` (:72-100), and the
API clamps of src/mcp_server.rs:81-110 (`search_documents`).  Everything numeric happens in
librlr_b200.so on the GPU; this file only keeps the `row -> chunk` table, resolves weights
and formats results.  There is no CPU fallback: without the CUDA library / a B200 every
search raises RlrError.

The embedding service (Ollama HTTP, src/embeddings.rs) and the reranker are out of scope;
`query` is therefore either an embedding vector or a string handed to a caller-supplied
`embedder` callable.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import binding as B

MAX_TOP_K = 100  # src/mcp_server.rs:364


@dataclass
class QueryWeights:
    """src/rag_engine.rs:1846-1863 -- all fields Option<f32>."""
    embedding: Optional[float] = None
    lexical: Optional[float] = None
    reranker: Optional[float] = None
    initial: Optional[float] = None

    def to_c(self) -> B.QueryWeightsC:
        c = B.QueryWeightsC()
        has = 0
        for bit, name in enumerate(("embedding", "lexical", "reranker", "initial")):
            v = getattr(self, name)
            if v is not None:
                setattr(c, name, float(v))
                has |= 1 << bit
        c.has = has
        return c


@dataclass
class ResolvedWeights:
    embedding: float
    lexical: float
    reranker: float
    initial: float


def resolve_weights(weights: Optional[QueryWeights]) -> ResolvedWeights:
    """ResolvedWeights::from_query_weights, src/rag_engine.rs:1888-1896."""
    lib = B.load()
    out = B.ResolvedWeightsC()
    B.check(lib.rlr_resolve_weights(C.byref(weights.to_c()) if weights is not None else None, C.byref(out)))
    return ResolvedWeights(out.embedding, out.lexical, out.reranker, out.initial)


@dataclass
class DocumentChunk:
    """src/rag_engine.rs:46-59 minus the embedding (which lives on the device)."""
    id: str
    document_name: str = ""
    text: str = ""
    chunk_index: int = 0
    page_number: int = 0
    section: Optional[str] = None
    metadata: dict = field(default_factory=dict)


@dataclass
class SearchResult:
    """src/rag_engine.rs:72-100."""
    text: str
    score: float
    document: str
    chunk_id: str
    chunk_index: int
    page_number: int
    section: Optional[str]
    embedding_score: Optional[float] = None
    lexical_score: Optional[float] = None
    initial_score: Optional[float] = None
    reranker_score: Optional[float] = None
    yes_logprob: Optional[float] = None
    no_logprob: Optional[float] = None
    row: int = -1  # position in the device store (not in the reference struct)

    def to_json(self) -> dict:
        d = {"text": self.text, "score": self.score, "document": self.document, "chunk_id": self.chunk_id,
             "chunk_index": self.chunk_index, "page_number": self.page_number, "section": self.section}
        for k in ("embedding_score", "lexical_score", "initial_score", "reranker_score", "yes_logprob", "no_logprob"):
            v = getattr(self, k)
            if v is not None:  # skip_serializing_if = "Option::is_none"
                d[k] = v
        return d


def sanitize_model_name(model_name: str) -> str:
    """src/rag_engine.rs:1435-1462."""
    trimmed = model_name.strip()
    if not trimmed:
        return "default"
    s = "".join(c if (c.isascii() and c.isalnum()) or c in "-_." else "_" for c in trimmed)
    if not s or all(c in "_." for c in s):
        return "default"
    return s


def get_index_path(data_dir: str, model_name: str) -> str:
    """src/rag_engine.rs:1465-1468."""
    return os.path.join(data_dir, f"chunks_{sanitize_model_name(model_name)}.json")


def get_sidecar_path(data_dir: str, model_name: str) -> str:
    """Binary twin of chunks_{model}.json (same sanitised model name, :1465-1468): `.rlrbin`."""
    return get_index_path(data_dir, model_name)[:-len(".json")] + ".rlrbin"


def get_legacy_path(data_dir: str) -> str:
    """src/rag_engine.rs:1471-1473."""
    return os.path.join(data_dir, "chunks.json")


def _f32_json(v: np.ndarray) -> List[str]:
    """Shortest decimal strings that round-trip every f32 (what serde_json's ryu writes, :1504; numpy's f32 -> str
    is the same shortest-repr algorithm): parsing them as f64 and narrowing -- what serde and `json.load` +
    np.float32 do -- returns the same bits."""
    return np.asarray(v, dtype=np.float32).astype(str).tolist()


def write_chunks_json(path: str, model: str, chunks: Sequence["DocumentChunk"], rows: np.ndarray, needs_reindex: bool,
                      document_hashes: dict) -> None:
    """save_to_disk, src/rag_engine.rs:1477-1518: PersistedState {version: 2, model, chunks: {id -> DocumentChunk},
    needs_reindex, document_hashes (omitted when empty)} as pretty JSON (serde_json::to_string_pretty: 2-space
    indent, one array element per line), written to `<path minus .json>.json.tmp` and renamed over `path`
    (:1494,:1506-1511).  `rows[i]` is chunk i's embedding as stored (normalised)."""
    if len(chunks) != len(rows):
        raise ValueError("one embedding row per chunk")
    marker = "@@RLR_EMBEDDING_%d@@"
    state = {"version": 2, "model": model,
             "chunks": {c.id: {"id": c.id, "document_name": c.document_name, "text": c.text, "embedding": marker % i,
                               "chunk_index": c.chunk_index, "page_number": c.page_number, "section": c.section,
                               "metadata": _metadata_json(c.metadata)} for i, c in enumerate(chunks)},
             "needs_reindex": bool(needs_reindex)}
    if document_hashes:                          # skip_serializing_if = "HashMap::is_empty"
        state["document_hashes"] = dict(document_hashes)
    text = json.dumps(state, indent=2, ensure_ascii=False)
    tmp = path[:-len(".json")] + ".json.tmp" if path.endswith(".json") else path + ".tmp"     # with_extension("json.tmp")
    with open(tmp, "w", encoding="utf-8") as f:
        pos = 0
        for i in range(len(chunks)):
            tag = '"' + (marker % i) + '"'
            at = text.index(tag, pos)
            f.write(text[pos:at])
            vals = _f32_json(rows[i])
            f.write("[\n        " + ",\n        ".join(vals) + "\n      ]" if vals else "[]")
            pos = at + len(tag)
        f.write(text[pos:])
        f.flush()
        os.fsync(f.fileno())
    os.replace(tmp, path)                        # atomic rename (:1509)


def _metadata_json(md: dict) -> dict:
    """ChunkMetadata, :35-42 (all five fields always serialised; Default::default() when absent)."""
    md = md or {}
    return {"page_range": md.get("page_range"), "sentence_range": md.get("sentence_range"),
            "section_title": md.get("section_title"), "token_count": int(md.get("token_count", 0)),
            "overlap_with_previous": int(md.get("overlap_with_previous", 0))}


@dataclass
class LoadDecision:
    """What load_from_disk (:1520-1652) decided before any embedding is touched."""
    state: Optional[dict]            # parsed PersistedState to apply, or None: start fresh
    source: Optional[str]            # file the state came from
    migrate: bool = False            # legacy chunks.json of the SAME model: save to the model-specific file after loading
    needs_reindex: bool = False      # start fresh but rebuild (corrupt model file / pre-model legacy chunks)


def _valid_state(d) -> bool:
    return isinstance(d, dict) and isinstance(d.get("version"), int) and isinstance(d.get("model"), str) \
        and isinstance(d.get("chunks"), dict)


def decide_load(data_dir: str, model: str) -> LoadDecision:
    """The file-selection half of load_from_disk, src/rag_engine.rs:1520-1652:
      1. the model-specific `chunks_{sanitized}.json` wins; if it does not parse, start fresh with
         needs_reindex (the corrupt file is kept, :1570-1582);
      2. else a legacy `chunks.json` is migrated ONLY when its `model` is the current one (:1592-1617); another
         model's legacy file is left alone (:1618-1626); a pre-model file (a bare id -> chunk map) with chunks in
         it asks for a reindex (:1627-1644);
      3. else start fresh."""
    specific, legacy = get_index_path(data_dir, model), get_legacy_path(data_dir)
    if os.path.exists(specific):
        try:
            with open(specific, "r", encoding="utf-8") as f:
                state = json.load(f)
            if not _valid_state(state):
                raise ValueError("not a PersistedState")
            return LoadDecision(state, specific)
        except (ValueError, OSError, UnicodeDecodeError):
            return LoadDecision(None, None, needs_reindex=True)
    if os.path.exists(legacy):
        try:
            with open(legacy, "r", encoding="utf-8") as f:
                data = json.load(f)
        except (ValueError, OSError, UnicodeDecodeError):
            return LoadDecision(None, None)
        if isinstance(data, dict) and isinstance(data.get("model"), str):            # ModelOnly peek, :1589
            if data["model"] == model and _valid_state(data):
                return LoadDecision(data, legacy, migrate=True)
            return LoadDecision(None, None)                                          # other model / unparsable: fresh
        if isinstance(data, dict) and data and all(isinstance(v, dict) and "embedding" in v for v in data.values()):
            return LoadDecision(None, None, needs_reindex=True)                      # :1632-1643
    return LoadDecision(None, None)


SIDECAR_MAGIC = b"RLRB200\x00"
# header: magic[8] | u32 index version (the reference's, 2) | u32 dim | u64 n_rows | u64 meta_bytes | u64 reserved
SIDECAR_HEADER = 40


class DeviceStore:
    """Owns one rlr_store handle."""

    _HOT_PREFIX = "rlr_"          # ClusterStore: "rlr_cluster_" (same arguments, same results)

    def __init__(self, handle, lib):
        self._h = handle
        self._lib = lib

    def _hot(self, name: str):
        return getattr(self._lib, self._HOT_PREFIX + name)

    @classmethod
    def from_rows(cls, rows: np.ndarray, device: int = 0, row_base: int = 0, flags: int = 0) -> "DeviceStore":
        lib = B.load()
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2:
            raise ValueError("rows must be (n, dim)")
        n, dim = rows.shape
        h = C.c_void_p()
        B.check(lib.rlr_store_create(device, dim, n, B.ptr(rows) if n else None, dim, row_base, flags, C.byref(h)))
        return cls(h, lib)

    @classmethod
    def empty(cls, n_rows: int, dim: int, device: int = 0, row_base: int = 0, flags: int = 0) -> "DeviceStore":
        lib = B.load()
        h = C.c_void_p()
        B.check(lib.rlr_store_create(device, dim, n_rows, None, dim, row_base, flags, C.byref(h)))
        return cls(h, lib)

    @classmethod
    def synthetic(cls, n_rows: int, dim: int, kind: int = B.RLR_SYNTH_IID, seed: int = 0x5EED0001,
                  centroid_seed: int = 0x5EED00C0, n_clusters: int = 4096, sigma: float = 0.65,
                  device: int = 0, row_base: int = 0, flags: int = 0) -> "DeviceStore":
        s = cls.empty(n_rows, dim, device=device, row_base=row_base, flags=flags)
        B.check(s._lib.rlr_store_fill_synthetic(s._h, kind, seed, centroid_seed, n_clusters, sigma))
        return s

    @property
    def handle(self):
        return self._h

    def info(self) -> B.StoreInfoC:
        out = B.StoreInfoC()
        B.check(self._lib.rlr_store_info_get(self._h, C.byref(out)))
        return out

    def read_rows(self, rows: Sequence[int]) -> np.ndarray:
        r = np.ascontiguousarray(rows, dtype=np.uint32)
        out = np.empty((len(r), self.info().dim), dtype=np.float32)
        B.check(self._lib.rlr_store_read_rows(self._h, B.ptr(r), len(r), B.ptr(out)))
        return out

    def upload(self, row0: int, rows: np.ndarray) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        B.check(self._lib.rlr_store_upload(self._h, row0, rows.shape[0], B.ptr(rows), rows.shape[1]))

    def append(self, rows: np.ndarray) -> int:
        """rlr_store_append: returns the global row of the first appended row."""
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        first = C.c_uint64(0)
        B.check(self._lib.rlr_store_append(self._h, rows.shape[0], B.ptr(rows) if rows.shape[0] else None,
                                           rows.shape[1] if rows.ndim == 2 else 0, C.byref(first)))
        return first.value

    def remove_rows(self, rows: Sequence[int]):
        """rlr_store_remove_rows: returns the (from, to) row moves that keep the store dense."""
        r = np.ascontiguousarray(rows, dtype=np.uint32)
        mf = np.zeros(max(len(r), 1), np.uint32); mt = np.zeros(max(len(r), 1), np.uint32)
        n = C.c_uint64(0)
        B.check(self._lib.rlr_store_remove_rows(self._h, B.ptr(r) if len(r) else None, len(r), B.ptr(mf), B.ptr(mt), C.byref(n)))
        return mf[:n.value], mt[:n.value]

    def close(self) -> None:
        if self._h:
            self._lib.rlr_store_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- raw hot-path calls (numpy in / numpy out) ----
    def search_topm(self, query: np.ndarray, m: int, w: ResolvedWeights, lex_rows=None, lex_scores=None,
                    flags: int = 0):
        q = np.ascontiguousarray(query, dtype=np.float32)
        rows = np.empty(m, np.uint32); comb = np.empty(m, np.float32)
        emb = np.empty(m, np.float32); lex = np.empty(m, np.float32)
        n = C.c_uint32(0)
        wc = B.ResolvedWeightsC(w.embedding, w.lexical, w.reranker, w.initial)
        lr = np.ascontiguousarray(lex_rows, dtype=np.uint32) if lex_rows is not None and len(lex_rows) else None
        ls = np.ascontiguousarray(lex_scores, dtype=np.float32) if lr is not None else None
        B.check(self._hot("search_topm")(self._h, B.ptr(q), q.shape[0], flags, C.byref(wc), B.ptr(lr), B.ptr(ls),
                                          0 if lr is None else len(lr), m, B.ptr(rows), B.ptr(comb), B.ptr(emb),
                                          B.ptr(lex), C.byref(n)))
        k = n.value
        return rows[:k], comb[:k], emb[:k], lex[:k]

    def mmr(self, cand_rows, relevance, top_k: int, lam: float, flags: int = 0) -> np.ndarray:
        r = np.ascontiguousarray(cand_rows, dtype=np.uint32)
        rel = np.ascontiguousarray(relevance, dtype=np.float32)
        out = np.empty(max(len(r), 1), np.uint32)
        n = C.c_uint32(0)
        B.check(self._hot("mmr")(self._h, B.ptr(r) if len(r) else None, B.ptr(rel) if len(r) else None, len(r),
                                  top_k, lam, flags, B.ptr(out), C.byref(n)))
        return out[:n.value]

    def search_mmr(self, query: np.ndarray, top_k: int, diversity: float, w: ResolvedWeights, lex_rows=None,
                   lex_scores=None, flags: int = 0):
        q = np.ascontiguousarray(query, dtype=np.float32)
        cap = max(top_k, 1)
        rows = np.empty(cap, np.uint32); score = np.empty(cap, np.float32)
        emb = np.empty(cap, np.float32); lex = np.empty(cap, np.float32)
        n = C.c_uint32(0)
        wc = B.ResolvedWeightsC(w.embedding, w.lexical, w.reranker, w.initial)
        lr = np.ascontiguousarray(lex_rows, dtype=np.uint32) if lex_rows is not None and len(lex_rows) else None
        ls = np.ascontiguousarray(lex_scores, dtype=np.float32) if lr is not None else None
        B.check(self._hot("search_mmr")(self._h, B.ptr(q), q.shape[0], flags, top_k, diversity, C.byref(wc),
                                         B.ptr(lr), B.ptr(ls), 0 if lr is None else len(lr), B.ptr(rows),
                                         B.ptr(score), B.ptr(emb), B.ptr(lex), C.byref(n)))
        k = n.value
        return rows[:k], score[:k], emb[:k], lex[:k]

    def search_text_topm(self, query: np.ndarray, m: int, w: ResolvedWeights, bm25, terms, flags: int = 0):
        """rlr_search_text_topm: RagEngine::search for a TEXT query -- BM25 (`bm25`: an rlr_bm25 handle, `terms`: the
        query's term ids in bytewise term order), blend, top-m, all on the device."""
        q = np.ascontiguousarray(query, dtype=np.float32)
        t = np.ascontiguousarray(terms, dtype=np.uint32)
        rows = np.empty(m, np.uint32); comb = np.empty(m, np.float32)
        emb = np.empty(m, np.float32); lex = np.empty(m, np.float32)
        n = C.c_uint32(0)
        wc = B.ResolvedWeightsC(w.embedding, w.lexical, w.reranker, w.initial)
        B.check(self._hot("search_text_topm")(self._h, bm25, B.ptr(q), q.shape[0], flags, C.byref(wc), B.ptr(t) if len(t) else None,
                                               len(t), m, B.ptr(rows), B.ptr(comb), B.ptr(emb), B.ptr(lex), C.byref(n)))
        k = n.value
        return rows[:k], comb[:k], emb[:k], lex[:k]

    def search_text_mmr(self, query: np.ndarray, top_k: int, diversity: float, w: ResolvedWeights, bm25, terms, flags: int = 0):
        """rlr_search_text_mmr: RagEngine::search_with_diversity for a TEXT query, one device sequence."""
        q = np.ascontiguousarray(query, dtype=np.float32)
        t = np.ascontiguousarray(terms, dtype=np.uint32)
        cap = max(top_k, 1)
        rows = np.empty(cap, np.uint32); score = np.empty(cap, np.float32)
        emb = np.empty(cap, np.float32); lex = np.empty(cap, np.float32)
        n = C.c_uint32(0)
        wc = B.ResolvedWeightsC(w.embedding, w.lexical, w.reranker, w.initial)
        B.check(self._hot("search_text_mmr")(self._h, bm25, B.ptr(q), q.shape[0], flags, top_k, diversity, C.byref(wc),
                                              B.ptr(t) if len(t) else None, len(t), B.ptr(rows), B.ptr(score), B.ptr(emb),
                                              B.ptr(lex), C.byref(n)))
        k = n.value
        return rows[:k], score[:k], emb[:k], lex[:k]

    def search_mmr_multi(self, queries: np.ndarray, top_k: int, diversity: float, w: ResolvedWeights, lex=None,
                         flags: int = 0):
        """rlr_search_mmr_multi / rlr_cluster_search_mmr_multi (throughput mode): up to RLR_MAX_MULTI queries answered by
        ONE pass over the rows (of every shard).
        `lex`: optional list of (lex_rows, lex_scores) per query (None entries allowed).  Returns a list of
        (rows, score, emb, lex) per query, each identical to search_mmr's result for that query."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq, dim = q.shape
        cap = max(top_k, 1)
        rows = np.zeros((nq, cap), np.uint32); score = np.zeros((nq, cap), np.float32)
        emb = np.zeros((nq, cap), np.float32); lx = np.zeros((nq, cap), np.float32)
        n = np.zeros(nq, np.uint32)
        wc = B.ResolvedWeightsC(w.embedding, w.lexical, w.reranker, w.initial)
        lr_ptrs = ls_ptrs = nl = None
        keep = []
        if lex is not None:
            lr_ptrs = (C.c_void_p * nq)(); ls_ptrs = (C.c_void_p * nq)(); nl = np.zeros(nq, np.uint32)
            for i, pair in enumerate(lex):
                if pair is None or pair[0] is None or len(pair[0]) == 0:
                    continue
                a = np.ascontiguousarray(pair[0], dtype=np.uint32); b = np.ascontiguousarray(pair[1], dtype=np.float32)
                keep += [a, b]
                lr_ptrs[i] = a.ctypes.data; ls_ptrs[i] = b.ctypes.data; nl[i] = len(a)
        B.check(self._hot("search_mmr_multi")(self._h, B.ptr(q), nq, dim, flags, top_k, diversity, C.byref(wc),
                                               lr_ptrs, ls_ptrs, B.ptr(nl) if nl is not None else None,
                                               B.ptr(rows), B.ptr(score), B.ptr(emb), B.ptr(lx), B.ptr(n)))
        return [(rows[i, :n[i]].copy(), score[i, :n[i]].copy(), emb[i, :n[i]].copy(), lx[i, :n[i]].copy()) for i in range(nq)]

    def embedding_candidates(self, query: np.ndarray, count: int, flags: int = 0):
        q = np.ascontiguousarray(query, dtype=np.float32)
        rows = np.empty(max(count, 1), np.uint32); score = np.empty(max(count, 1), np.float32)
        n = C.c_uint32(0)
        B.check(self._hot("embedding_candidates")(self._h, B.ptr(q), q.shape[0], flags, count, B.ptr(rows),
                                                   B.ptr(score), C.byref(n)))
        return rows[:n.value], score[:n.value]

    def search_batch(self, queries: np.ndarray, m: int, flags: int = 0):
        """rlr_search_batch: (rows [nq, m], scores [nq, m], n [nq]) for a batch of query embeddings."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq, dim = q.shape
        rows = np.zeros((nq, m), np.uint32); scores = np.zeros((nq, m), np.float32); n = np.zeros(nq, np.uint32)
        B.check(self._lib.rlr_search_batch(self._h, B.ptr(q), nq, dim, flags, m, B.ptr(rows), B.ptr(scores), B.ptr(n)))
        return rows, scores, n

    def search_batch_device(self, queries: np.ndarray, m: int, d_keys, d_cnt=None, stream=None, flags: int = 0) -> None:
        """rlr_search_batch_device: per-query rank keys stay in HBM (d_keys: [nq, m] int64 torch tensor)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq, dim = q.shape
        B.check(self._lib.rlr_search_batch_device(self._h, B.ptr(q), nq, dim, flags, m, C.c_void_p(d_keys.data_ptr()),
                                                  C.c_void_p(d_cnt.data_ptr()) if d_cnt is not None else None,
                                                  C.c_void_p(stream) if stream else None))

    def batch_merge(self, d_lists, n_lists: int, nq: int, m: int, d_out, d_out_cnt=None, stream=None) -> None:
        """rlr_batch_merge_async over all-gathered key lists [n_lists, nq, m]."""
        B.check(self._lib.rlr_batch_merge_async(self._h, C.c_void_p(d_lists.data_ptr()), n_lists, nq, m,
                                                C.c_void_p(d_out.data_ptr()),
                                                C.c_void_p(d_out_cnt.data_ptr()) if d_out_cnt is not None else None,
                                                C.c_void_p(stream) if stream else None))

    def last_timings(self) -> B.TimingsC:
        t = B.TimingsC()
        B.check(self._lib.rlr_last_timings(C.byref(t)))
        return t


class ClusterStore(DeviceStore):
    """Owns one rlr_cluster handle: the same store row-sharded over several GPUs of one box and driven from
    THIS process (the reference is one process, src/main.rs:140-167).  Same raw hot-path calls as DeviceStore
    (`search_topm`, `mmr`, `search_mmr`, `embedding_candidates`), same results bit for bit; rows are global."""

    _HOT_PREFIX = "rlr_cluster_"

    @staticmethod
    def _create(lib, devices, dim, n, rows, flags, shard_rows):
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        sr = np.ascontiguousarray(shard_rows, dtype=np.uint64) if shard_rows is not None else None
        if sr is not None and len(sr) != len(dev):
            raise ValueError("shard_rows needs one entry per device")
        h = C.c_void_p()
        B.check(lib.rlr_cluster_create(B.ptr(dev), len(dev), dim, n, B.ptr(rows) if rows is not None and n else None, dim,
                                       flags, B.ptr(sr), C.byref(h)))
        return h

    @classmethod
    def from_rows(cls, rows: np.ndarray, devices: Sequence[int] = (0,), flags: int = 0, shard_rows=None) -> "ClusterStore":
        lib = B.load()
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2:
            raise ValueError("rows must be (n, dim)")
        return cls(cls._create(lib, devices, rows.shape[1], rows.shape[0], rows, flags, shard_rows), lib)

    @classmethod
    def empty(cls, n_rows: int, dim: int, devices: Sequence[int] = (0,), flags: int = 0, shard_rows=None) -> "ClusterStore":
        lib = B.load()
        return cls(cls._create(lib, devices, dim, n_rows, None, flags, shard_rows), lib)

    @classmethod
    def synthetic(cls, n_rows: int, dim: int, kind: int = B.RLR_SYNTH_IID, seed: int = 0x5EED0001,
                  centroid_seed: int = 0x5EED00C0, n_clusters: int = 4096, sigma: float = 0.65,
                  devices: Sequence[int] = (0,), flags: int = 0, shard_rows=None) -> "ClusterStore":
        s = cls.empty(n_rows, dim, devices=devices, flags=flags, shard_rows=shard_rows)
        B.check(s._lib.rlr_cluster_fill_synthetic(s._h, kind, seed, centroid_seed, n_clusters, sigma))
        return s

    def cluster_info(self) -> B.ClusterInfoC:
        out = B.ClusterInfoC()
        B.check(self._lib.rlr_cluster_info_get(self._h, C.byref(out)))
        return out

    def info(self) -> B.StoreInfoC:
        """The cluster seen as one store (row_base 0, device = the root's)."""
        ci = self.cluster_info()
        out = B.StoreInfoC()
        out.n_rows, out.row_base, out.dim, out.pitch, out.device, out.flags = ci.n_rows, 0, ci.dim, ci.pitch, ci.device[0], ci.flags
        return out

    def read_rows(self, rows: Sequence[int]) -> np.ndarray:
        r = np.ascontiguousarray(rows, dtype=np.uint32)
        out = np.empty((len(r), self.cluster_info().dim), dtype=np.float32)
        B.check(self._lib.rlr_cluster_read_rows(self._h, B.ptr(r), len(r), B.ptr(out)))
        return out

    def upload(self, row0: int, rows: np.ndarray) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        B.check(self._lib.rlr_cluster_upload(self._h, row0, rows.shape[0], B.ptr(rows), rows.shape[1]))

    def append(self, rows):
        raise B.RlrError(B.RLR_ERR_UNSUPPORTED, "a cluster is a bulk-loaded snapshot: mutate a single-GPU store or rebuild")

    remove_rows = append

    def last_scan_ms(self) -> List[float]:
        buf = np.zeros(B.RLR_MAX_SHARDS, np.float32)
        n = C.c_uint32(0)
        B.check(self._lib.rlr_cluster_last_scan_ms(B.ptr(buf), len(buf), C.byref(n)))
        return buf[:n.value].tolist()

    def launches(self) -> int:
        n = C.c_uint64(0)
        B.check(self._lib.rlr_cluster_launch_count(self._h, C.byref(n)))
        return n.value

    def search_batch(self, queries: np.ndarray, m: int, flags: int = 0):
        """rlr_cluster_search_batch: the tcgen05 contraction on every GPU's shard, per-query merge on the root."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq, dim = q.shape
        rows = np.zeros((nq, m), np.uint32); scores = np.zeros((nq, m), np.float32); n = np.zeros(nq, np.uint32)
        B.check(self._lib.rlr_cluster_search_batch(self._h, B.ptr(q), nq, dim, flags, m, B.ptr(rows), B.ptr(scores), B.ptr(n)))
        return rows, scores, n

    def close(self) -> None:
        if self._h:
            self._lib.rlr_cluster_destroy(self._h)
            self._h = None


Query = Union[str, Sequence[float], np.ndarray]
LexicalFn = Callable[[str, int], List[Tuple[str, float]]]


class LexicalIndex:
    """LexicalIndex (src/rag_engine.rs:2083-2231) over the host-mirror support library's BM25 twin
    (librlr_hostmirror.so, rlr_lexical_*; NOT part of the product library): chunk ids are mapped to u64 keys in insertion order, which is also the
    tie order of equal scores."""

    def __init__(self):
        self._lib = B.load_hostmirror()
        self._h = C.c_void_p()
        B.check_hm(self._lib.rlr_lexical_create(C.byref(self._h)))
        self._key_of, self._id_of, self._next = {}, {}, 0

    def add_chunk(self, chunk_id: str, text: str) -> None:
        key = self._key_of.get(chunk_id)
        if key is None:
            key = self._next
            self._next += 1
            self._key_of[chunk_id] = key
            self._id_of[key] = chunk_id
        b = text.encode("utf-8")
        B.check_hm(self._lib.rlr_lexical_add_chunk(self._h, key, b, len(b)))

    def remove_chunk(self, chunk_id: str) -> None:
        key = self._key_of.pop(chunk_id, None)
        if key is not None:
            self._id_of.pop(key, None)
            B.check_hm(self._lib.rlr_lexical_remove_chunk(self._h, key))

    def contains(self, chunk_id: str) -> bool:
        key = self._key_of.get(chunk_id)
        if key is None:
            return False
        out = C.c_int(0)
        B.check_hm(self._lib.rlr_lexical_contains(self._h, key, C.byref(out)))
        return bool(out.value)

    def score(self, query: str, limit: int) -> List[Tuple[str, float]]:
        b = query.encode("utf-8")
        cap = max(int(limit), 1) if limit > 0 else max(len(self._key_of), 1)
        keys, scores, n = np.zeros(cap, np.uint64), np.zeros(cap, np.float32), C.c_uint32(0)
        B.check_hm(self._lib.rlr_lexical_score(self._h, b, len(b), int(limit), B.ptr(keys), B.ptr(scores), cap, C.byref(n)))
        return [(self._id_of[int(k)], float(s)) for k, s in zip(keys[:n.value], scores[:n.value])]

    def __call__(self, query: str, limit: int) -> List[Tuple[str, float]]:
        return self.score(query, limit)

    def close(self) -> None:
        if self._h:
            self._lib.rlr_lexical_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def tokenize(text: str) -> List[str]:
    """fn tokenize (src/rag_engine.rs:2242-2247) through the host-mirror support library (strings stay on the host)."""
    lib = B.load_hostmirror()
    b = text.encode("utf-8")
    out = C.create_string_buffer(3 * len(b) + 16)          # lowercase mappings may grow a token
    n, nt = C.c_size_t(0), C.c_uint32(0)
    B.check_hm(lib.rlr_tokenize(b, len(b), out, len(out), C.byref(n), C.byref(nt)))
    return out.raw[:n.value].decode("utf-8").split("\n") if n.value else []


class DeviceLexicalIndex:
    """LexicalIndex (src/rag_engine.rs:2083-2231) with the postings scored ON THE DEVICE (rlr_bm25_*).
    The host keeps what is string work: the tokenizer and the term -> id dictionary.  Chunks are identified by their
    row in the store."""

    def __init__(self, store: "DeviceStore"):
        self._lib = B.load()
        self._store = store
        self._h = C.c_void_p()
        # a ClusterStore gets rlr_cluster_bm25_* (one device index per shard, global statistics): same calls, same results
        self._prefix = store._HOT_PREFIX + "bm25_"
        B.check(self._fn("create")(store.handle, C.byref(self._h)))
        self.vocab = {}

    def _fn(self, name: str):
        return getattr(self._lib, self._prefix + name)

    @property
    def handle(self):
        return self._h

    def add_chunk(self, row: int, text: str) -> None:
        counts = {}
        for t in tokenize(text):
            counts[t] = counts.get(t, 0) + 1
        ids = np.array([self.vocab.setdefault(t, len(self.vocab)) for t in counts], np.uint32)
        tfs = np.array(list(counts.values()), np.uint32)
        B.check(self._fn("set_doc")(self._h, row, B.ptr(ids) if len(ids) else None, B.ptr(tfs) if len(ids) else None, len(ids)))

    def add_documents_csr(self, row0: int, offsets: np.ndarray, term_ids: np.ndarray, term_freqs: np.ndarray) -> None:
        """Bulk add_chunk for rows [row0, row0 + len(offsets) - 1) from numeric term ids (rlr_bm25_set_docs)."""
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        ids = np.ascontiguousarray(term_ids, dtype=np.uint32); tfs = np.ascontiguousarray(term_freqs, dtype=np.uint32)
        B.check(self._fn("set_docs")(self._h, row0, len(off) - 1, B.ptr(off), B.ptr(ids) if len(ids) else None, B.ptr(tfs) if len(ids) else None))

    def remove_chunk(self, row: int) -> None:
        B.check(self._fn("remove_doc")(self._h, row))

    def move(self, from_row: int, to_row: int) -> None:
        if isinstance(self._store, ClusterStore):
            raise B.RlrError(B.RLR_ERR_UNSUPPORTED, "a cluster has no row removal to follow")
        B.check(self._lib.rlr_bm25_move_doc(self._h, from_row, to_row))

    def query_terms(self, query: str) -> np.ndarray:
        """The query's known term ids in bytewise order of the term strings (the summation order, see rlr_b200.h)."""
        terms = sorted(set(tokenize(query)), key=lambda t: t.encode("utf-8"))
        return np.array([self.vocab[t] for t in terms if t in self.vocab], np.uint32)

    def stats(self):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        B.check(self._fn("stats")(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def score(self, query: str, limit: int) -> List[Tuple[int, float]]:
        """LexicalIndex::score(query, limit) -> [(row, score)], score desc (ties: lower row)."""
        t = self.query_terms(query)
        rows, sc, n = np.zeros(limit, np.uint32), np.zeros(limit, np.float32), C.c_uint32(0)
        B.check(self._fn("score")(self._h, B.ptr(t) if len(t) else None, len(t), limit, B.ptr(rows), B.ptr(sc), limit, C.byref(n)))
        return [(int(r), np.float32(x)) for r, x in zip(rows[:n.value], sc[:n.value])]

    def close(self) -> None:
        if self._h:
            self._fn("destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _make_store(rows: np.ndarray, device: int, devices: Optional[Sequence[int]]):
    """One GPU -> DeviceStore; `devices=[...]` -> the same rows sharded over those GPUs (ClusterStore)."""
    if devices is not None and len(devices) > 1:
        return ClusterStore.from_rows(rows, devices=devices)
    return DeviceStore.from_rows(rows, device=device if devices is None else devices[0])


class RagEngine:
    """The retrieval half of the reference's RagEngine (src/rag_engine.rs:104-113): chunk
    metadata on the host, embeddings on the device."""

    def __init__(self, chunks: List[DocumentChunk], store: DeviceStore, model: str = "nomic-embed-text",
                 embedder: Optional[Callable[[str], Sequence[float]]] = None,
                 lexical: Union[None, str, LexicalFn, "LexicalIndex"] = None):
        self.chunks = chunks                     # row -> chunk (the `row -> chunk_id` table)
        self.row_of = {c.id: i for i, c in enumerate(chunks)}
        self.store = store
        self.model = model
        self.embedder = embedder
        if lexical == "bm25-device":             # the same index with the postings on the GPU (rlr_bm25_*)
            lexical = DeviceLexicalIndex(store)
            for i, c in enumerate(chunks):
                lexical.add_chunk(i, c.text)
        if lexical == "bm25":                    # validate_index_sync, :1375-1389: index every chunk's text
            lexical = LexicalIndex()
            for c in chunks:
                lexical.add_chunk(c.id, c.text)
        self.lexical = lexical                   # LexicalIndex (host BM25) or any callable (query, limit) -> [(id, score)]
        self.needs_reindex = False
        self.document_hashes = {}

    # ---- construction ----
    @classmethod
    def load_from_disk(cls, data_dir: str, model: str = "nomic-embed-text", device: int = 0,
                       devices: Optional[Sequence[int]] = None, **kw) -> "RagEngine":
        """load_from_disk + apply_loaded_state, src/rag_engine.rs:1520-1696: the model-specific file
        `chunks_{sanitized}.json` (:1465-1468) first, then the legacy `chunks.json` migration; see decide_load."""
        d = decide_load(data_dir, model)
        if d.state is None:
            eng = cls([], DeviceStore.from_rows(np.zeros((0, 1), np.float32), device=device), model=model, **kw)
            eng.needs_reindex = d.needs_reindex
            eng.data_dir = data_dir
            return eng
        eng = cls._from_state(d.state, model=model, device=device, devices=devices, **kw)
        eng.data_dir = data_dir
        if int(d.state["version"]) < 2:          # :1664-1673: wiped, marked, and the wipe is persisted
            eng.save_to_disk()
        elif d.migrate:                          # :1699-1706: legacy file preserved, model-specific file written
            eng.save_to_disk()
        return eng

    def save_to_disk(self, data_dir: Optional[str] = None) -> str:
        """save_to_disk, src/rag_engine.rs:1477-1518 (see write_chunks_json).  The embeddings are read back from
        the device: they are the normalised rows the searches scan."""
        data_dir = data_dir if data_dir is not None else getattr(self, "data_dir", None)
        if data_dir is None:
            raise ValueError("no data_dir: pass one or load the engine with load_from_disk")
        path = get_index_path(data_dir, self.model)
        n = len(self.chunks)
        rows = self.store.read_rows(np.arange(n)) if n else np.zeros((0, 0), np.float32)
        write_chunks_json(path, self.model, self.chunks, rows, self.needs_reindex, self.document_hashes)
        return path

    @classmethod
    def from_chunks_json(cls, path: str, model: str = "nomic-embed-text", device: int = 0,
                         devices: Optional[Sequence[int]] = None, **kw) -> "RagEngine":
        with open(path, "r", encoding="utf-8") as f:
            state = json.load(f)
        return cls._from_state(state, model=model, device=device, devices=devices, **kw)

    @classmethod
    def _from_state(cls, state: dict, model: str = "nomic-embed-text", device: int = 0,
                    devices: Optional[Sequence[int]] = None, **kw) -> "RagEngine":
        """apply_loaded_state, src/rag_engine.rs:1655-1709."""
        version = int(state["version"])
        chunks_map = state.get("chunks", {})
        if version < 2:                          # :1664-1673 outdated index: wipe, mark for reindex
            eng = cls([], DeviceStore.from_rows(np.zeros((0, 1), np.float32), device=device), model=model, **kw)
            eng.needs_reindex = True
            return eng
        # HashMap iteration order is arbitrary in the reference; rows here follow file order.
        metas: List[DocumentChunk] = []
        dim = None
        embs = []
        for cid, ch in chunks_map.items():
            e = np.asarray(ch["embedding"], dtype=np.float32)
            if dim is None:
                dim = e.shape[0]
            elif e.shape[0] != dim:
                raise ValueError(f"chunk {cid}: embedding has {e.shape[0]} dims, expected {dim}")
            embs.append(e)
            metas.append(DocumentChunk(id=ch.get("id", cid), document_name=ch.get("document_name", ""),
                                       text=ch.get("text", ""), chunk_index=int(ch.get("chunk_index", 0)),
                                       page_number=int(ch.get("page_number", 0)), section=ch.get("section"),
                                       metadata=ch.get("metadata") or {}))
        rows = np.stack(embs) if embs else np.zeros((0, 1), np.float32)
        lib = B.load()
        for i in range(rows.shape[0]):           # :1678-1680 re-normalise every embedding at load
            B.check(lib.rlr_normalize(rows[i].ctypes.data_as(C.POINTER(C.c_float)), rows.shape[1]))
        eng = cls(metas, _make_store(rows, device, devices), model=model, **kw)
        eng.needs_reindex = bool(state.get("needs_reindex", False))
        eng.document_hashes = dict(state.get("document_hashes", {}))
        if not eng.document_hashes and metas:    # :1686-1691
            eng.needs_reindex = True
        return eng

    # ---- binary sidecar: the same PersistedState without ~10 bytes of JSON per float ----
    def save_sidecar(self, path: str, block_rows: int = 1 << 17) -> None:
        """Atomic tmp + rename like save_to_disk (:1600-1640).  Layout: 40-byte header, n_rows x dim f32
        rows (little endian, as stored = as the reference would serialise them), then a UTF-8 JSON blob
        with everything else of PersistedState (:1479-1486) and the chunks in ROW order."""
        import struct
        info = self.store.info()
        n, dim = int(info.n_rows), int(info.dim)
        meta = json.dumps({"model": self.model, "needs_reindex": bool(self.needs_reindex),
                           "document_hashes": self.document_hashes,
                           "chunks": [{"id": c.id, "document_name": c.document_name, "text": c.text,
                                       "chunk_index": c.chunk_index, "page_number": c.page_number, "section": c.section,
                                       "metadata": c.metadata} for c in self.chunks]}, ensure_ascii=False).encode("utf-8")
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(SIDECAR_MAGIC + struct.pack("<IIQQQ", 2, dim, n, len(meta), 0))
            base = int(info.row_base)
            for r0 in range(0, n, block_rows):
                r1 = min(n, r0 + block_rows)
                f.write(np.ascontiguousarray(self.store.read_rows(np.arange(base + r0, base + r1)), np.float32).tobytes())
            f.write(meta)
            f.flush()
            os.fsync(f.fileno())
        os.replace(tmp, path)

    @classmethod
    def from_sidecar(cls, path: str, model: str = "nomic-embed-text", device: int = 0, store_flags: int = 0,
                     block_rows: int = 1 << 18, **kw) -> "RagEngine":
        """apply_loaded_state (:1655-1696) for the binary sidecar: version gate, re-normalise EVERY row at load
        (:1678-1680) -- on the device, same sequential arithmetic, same bits (RLR_STORE_NORMALIZE_ON_UPLOAD) --
        and the unfingerprinted-index rule (:1686-1691).  Rows stream from a memory map in blocks."""
        import struct
        with open(path, "rb") as f:
            head = f.read(SIDECAR_HEADER)
        if len(head) != SIDECAR_HEADER or head[:8] != SIDECAR_MAGIC:
            raise ValueError(f"{path}: not an rlr_b200 sidecar")
        version, dim, n, meta_bytes, _ = struct.unpack("<IIQQQ", head[8:])
        if version < 2:                          # :1664-1673 outdated index: wipe, mark for reindex
            eng = cls([], DeviceStore.from_rows(np.zeros((0, 1), np.float32), device=device), model=model, **kw)
            eng.needs_reindex = True
            return eng
        rows_bytes = n * dim * 4
        if os.path.getsize(path) != SIDECAR_HEADER + rows_bytes + meta_bytes:
            raise ValueError(f"{path}: truncated or corrupt sidecar")
        with open(path, "rb") as f:
            f.seek(SIDECAR_HEADER + rows_bytes)
            meta = json.loads(f.read(meta_bytes).decode("utf-8"))
        if len(meta["chunks"]) != n:
            raise ValueError(f"{path}: {len(meta['chunks'])} chunk records for {n} rows")
        metas = [DocumentChunk(id=ch["id"], document_name=ch.get("document_name", ""), text=ch.get("text", ""),
                               chunk_index=int(ch.get("chunk_index", 0)), page_number=int(ch.get("page_number", 0)),
                               section=ch.get("section"), metadata=ch.get("metadata") or {}) for ch in meta["chunks"]]
        if n == 0:
            store = DeviceStore.from_rows(np.zeros((0, max(dim, 1)), np.float32), device=device)
        else:
            rows = np.memmap(path, dtype="<f4", mode="r", offset=SIDECAR_HEADER, shape=(n, dim))
            store = DeviceStore.empty(n, dim, device=device, flags=store_flags | B.RLR_STORE_NORMALIZE_ON_UPLOAD)
            for r0 in range(0, n, block_rows):
                store.upload(r0, np.ascontiguousarray(rows[r0:min(n, r0 + block_rows)]))
            del rows
        eng = cls(metas, store, model=model, **kw)
        eng.needs_reindex = bool(meta.get("needs_reindex", False))
        eng.document_hashes = dict(meta.get("document_hashes", {}))
        if not eng.document_hashes and metas:    # :1686-1691
            eng.needs_reindex = True
        return eng

    @classmethod
    def from_rows(cls, rows: np.ndarray, chunk_ids: Optional[Sequence[str]] = None, normalize: bool = True,
                  device: int = 0, devices: Optional[Sequence[int]] = None, **kw) -> "RagEngine":
        rows = np.array(rows, dtype=np.float32, order="C")
        if normalize:                            # :359 normalise at insert
            lib = B.load()
            for i in range(rows.shape[0]):
                B.check(lib.rlr_normalize(rows[i].ctypes.data_as(C.POINTER(C.c_float)), rows.shape[1]))
        ids = list(chunk_ids) if chunk_ids is not None else [f"chunk-{i}" for i in range(rows.shape[0])]
        metas = [DocumentChunk(id=i) for i in ids]
        return cls(metas, _make_store(rows, device, devices), **kw)

    # ---- mutation: add_document, src/rag_engine.rs:219-402 (the store half: :347-386) ----
    def replace_document(self, document_name: str, chunks: List[DocumentChunk], embeddings: np.ndarray) -> None:
        """Drop every chunk of `document_name` (`chunks.retain`, :347-348) and insert the new ones with
        their embeddings normalised (:359).  Keeps the device store dense and `row -> chunk` in sync."""
        old = [i for i, c in enumerate(self.chunks) if c.document_name == document_name]
        if isinstance(self.lexical, LexicalIndex):
            for i in old:                        # validate_index_sync -> drop_stale, :1379
                self.lexical.remove_chunk(self.chunks[i].id)
            for c in chunks:                     # :382
                self.lexical.add_chunk(c.id, c.text)
        if old:
            mf, mt = self.store.remove_rows(old)
            for f, t in zip(mf.tolist(), mt.tolist()):
                self.chunks[t] = self.chunks[f]
            del self.chunks[len(self.chunks) - len(old):]
        emb = np.array(embeddings, dtype=np.float32, order="C")
        lib = B.load()
        if not (self.store.info().flags & B.RLR_STORE_NORMALIZE_ON_UPLOAD):     # else the store normalises on append
            for i in range(emb.shape[0]):
                B.check(lib.rlr_normalize(emb[i].ctypes.data_as(C.POINTER(C.c_float)), emb.shape[1]))
        if len(chunks):
            first = self.store.append(emb)
            assert first == len(self.chunks)
            self.chunks.extend(chunks)
        self.row_of = {c.id: i for i, c in enumerate(self.chunks)}

    # ---- helpers ----
    def _embed(self, query: Query) -> np.ndarray:
        if isinstance(query, str):
            if self.embedder is None:
                raise B.RlrError(B.RLR_ERR_INVALID_ARG, "string query but no embedder configured "
                                 "(EmbeddingService is out of scope; pass an embedding)")
            return np.asarray(self.embedder(query), dtype=np.float32)
        return np.asarray(query, dtype=np.float32)

    def _lex(self, query: Query, top_k: int):
        if self.lexical is None or not isinstance(query, str) or isinstance(self.lexical, DeviceLexicalIndex):
            return None, None
        pairs = self.lexical(query, top_k * 5)   # :505 lexical_index.score(query, top_k*5)
        rows, scores = [], []
        for cid, sc in pairs:
            r = self.row_of.get(cid)
            if r is not None:
                rows.append(r); scores.append(sc)
        return (np.asarray(rows, np.uint32), np.asarray(scores, np.float32)) if rows else (None, None)

    def _result(self, row: int, score: float, emb: float, lex: float) -> SearchResult:
        ch = self.chunks[row]
        # fallback-fill fields, :680-695 (no reranker => reranker fields None)
        return SearchResult(text=ch.text, score=float(score), document=ch.document_name, chunk_id=ch.id,
                            chunk_index=ch.chunk_index, page_number=ch.page_number, section=ch.section,
                            embedding_score=float(emb), lexical_score=float(lex), initial_score=float(score),
                            row=int(row))

    # ---- the reference's public methods ----
    def search(self, query: Query, top_k: int, weights: Optional[QueryWeights] = None) -> List[SearchResult]:
        """RagEngine::search, :470-701 (reranker absent)."""
        if not self.chunks:                      # :476-478
            return []
        w = resolve_weights(weights)             # :481
        top_k = max(int(top_k), 1)               # :490
        q = self._embed(query)
        if isinstance(self.lexical, DeviceLexicalIndex) and isinstance(query, str):
            rows, comb, emb, lex = self.store.search_text_topm(q, top_k, w, self.lexical.handle, self.lexical.query_terms(query))
            return [self._result(r, c, e, l) for r, c, e, l in zip(rows, comb, emb, lex)]
        lr, ls = self._lex(query, top_k)
        rows, comb, emb, lex = self.store.search_topm(q, top_k, w, lr, ls)
        return [self._result(r, c, e, l) for r, c, e, l in zip(rows, comb, emb, lex)]

    def search_with_diversity(self, query: Query, top_k: int, diversity_factor: float,
                              weights: Optional[QueryWeights] = None) -> List[SearchResult]:
        """RagEngine::search_with_diversity, :717-759 (one fused device call)."""
        if not self.chunks:
            return []
        w = resolve_weights(weights)
        q = self._embed(query)
        if isinstance(self.lexical, DeviceLexicalIndex) and isinstance(query, str):
            rows, score, emb, lex = self.store.search_text_mmr(q, int(top_k), float(diversity_factor), w, self.lexical.handle,
                                                                self.lexical.query_terms(query))
            return [self._result(r, c, e, l) for r, c, e, l in zip(rows, score, emb, lex)]
        lam = min(max(float(diversity_factor), 0.0), 1.0) if diversity_factor == diversity_factor else diversity_factor
        pool = max(int(top_k), 1) if lam == 0.0 else max(3 * int(top_k), int(top_k) + 10)
        lr, ls = self._lex(query, pool)
        rows, score, emb, lex = self.store.search_mmr(q, int(top_k), float(diversity_factor), w, lr, ls)
        return [self._result(r, c, e, l) for r, c, e, l in zip(rows, score, emb, lex)]

    def get_embedding_candidates(self, query: Query, count: int):
        """RagEngine::get_embedding_candidates, :415-461 -> (chunk_id, document, text, page, section, initial_score)."""
        if not self.chunks:
            return []
        q = self._embed(query)
        rows, score = self.store.embedding_candidates(q, int(count))
        out = []
        for r, s in zip(rows, score):
            ch = self.chunks[int(r)]
            out.append({"chunk_id": ch.id, "document": ch.document_name, "text": ch.text,
                        "page_number": ch.page_number, "section": ch.section, "initial_score": float(s)})
        return out


def search_documents(engine: RagEngine, query: Query, top_k: Optional[int] = None,
                     diversity_factor: Optional[float] = None, weights: Optional[QueryWeights] = None):
    """MCP tool `search_documents`, src/mcp_server.rs:81-110: parameter defaults and clamps."""
    k = min(5 if top_k is None else int(top_k), MAX_TOP_K)          # :85
    lam = 0.3 if diversity_factor is None else float(diversity_factor)
    lam = min(max(lam, 0.0), 1.0)                                    # :86
    return engine.search_with_diversity(query, k, lam, weights)
