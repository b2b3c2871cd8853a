"""Row-sharded search across the GPUs of one NVSwitch box: one process per GPU,
`torch.distributed` (NCCL over NVLink) for the plumbing.  SURVEY.md 8(e).

Per query, on every rank g of G -- FUSED EXCHANGE (the product path on an NVSwitch box):
  1. local fused scan + top-P; the scan kernel's last CTA stores the rank's list straight into
     rank 0's mailbox in rank 0's HBM (NVLink peer stores from inside the compute kernel) and
     publishes the query's sequence number                           (rlr_topm_post_async)
     -- ranks != 0 are done here and move on to the next query;
  2. rank 0: one kernel waits for the G flags, merges the lists, frees the slot
                                                                     (rlr_mailbox_merge_async)
  3. rank 0: MMR whose pairwise kernel loads the pool rows from the owning GPUs' HBM
                                                                     (rlr_mmr_peers_async)
  Rank 0 carries the merge + MMR tail, so `ShardPlan(head_rows=...)` gives it a smaller row
  block (tail-balanced sharding): all ranks then finish a query at the same time.
COLLECTIVE path (NCCL; what the gloo tests drive with a stand-in backend):
  1. local fused scan + top-P over the rank's contiguous row block   (rlr_topm_async)
  2. ONE all-gather of the fixed-size per-rank lists (P x 16 B)      (NCCL)
  3. merge of the G lists to the global pool of P                    (rlr_merge_async)
  4. MMR on rank 0, whose pairwise-similarity kernel loads the pool rows straight from the
     owning GPUs' HBM through CUDA-IPC peer mappings (NVLink loads)  (rlr_mmr_peers_async)
  Fallback when peer mappings are not set up (and what the gloo tests drive):
  4'. gather of the pool rows this rank owns into a P x pitch matrix (rlr_gather_async)
  5'. reduce-to-rank-0 of that matrix as int32 bit patterns: every row is non-zero on
      exactly one rank, and integer x + 0 == x, so the transport is bit-exact (NCCL)
  6'. MMR on rank 0 over the gathered matrix                         (rlr_mmr_async)
Nothing here computes on the host; the steps are enqueued on the current CUDA stream.

The choreography takes a `backend` object so that tests/test_dist_gloo.py can drive the
same collective sequence on CPU tensors over gloo with a test-only stand-in backend; the
product backend is `CudaBackend` and needs the CUDA library and a B200.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

CAND_WORDS = 2  # one rlr_cand == two int64 words: key, (emb f32 | lex f32 << 32)


@dataclass(frozen=True)
class ShardPlan:
    """Contiguous row blocks (global row = row0 + local).  Even split [g*N/G, (g+1)*N/G) by
    default; with `head_rows` rank 0 owns exactly that many rows and ranks 1..G-1 split the
    rest evenly (tail-balanced sharding: rank 0 also runs the merge + MMR of every query)."""
    n_total: int
    world: int
    rank: int
    head_rows: Optional[int] = None

    @staticmethod
    def bounds(n_total: int, world: int, rank: int, head_rows: Optional[int] = None):
        if head_rows is None or world == 1:
            lo = (n_total * rank) // world
            hi = (n_total * (rank + 1)) // world
            return lo, hi
        head = min(max(int(head_rows), 0), n_total)
        if rank == 0:
            return 0, head
        rest, g = n_total - head, world - 1
        return head + (rest * (rank - 1)) // g, head + (rest * rank) // g

    @staticmethod
    def balanced_head_rows(n_total: int, world: int, tail_rows: float) -> int:
        """Rows for rank 0 such that scan(rank 0) + tail == scan(other ranks), the tail being
        expressed in rows-scanned-per-second units: n0 = N/G - tail_rows*(G-1)/G."""
        if world == 1:
            return n_total
        n0 = n_total / world - tail_rows * (world - 1) / world
        return int(min(max(n0, min(n_total, 1024)), n_total))

    def _b(self, rank: int):
        return self.bounds(self.n_total, self.world, rank, self.head_rows)

    @property
    def row0(self) -> int:
        return self._b(self.rank)[0]

    @property
    def n_local(self) -> int:
        lo, hi = self._b(self.rank)
        return hi - lo

    def owner(self, row: int) -> int:
        for g in range(self.world):
            lo, hi = self._b(g)
            if lo <= row < hi:
                return g
        raise ValueError(row)


def pool_size(top_k: int, lam: float) -> int:
    """m handed to `search` by search_with_diversity: src/rag_engine.rs:728-734 (+ :490)."""
    if lam == 0.0:
        return max(int(top_k), 1)
    return max(3 * int(top_k), int(top_k) + 10)


def clamp_lambda(lam: float) -> float:
    """f32::clamp(0.0, 1.0), src/rag_engine.rs:725 (NaN stays NaN)."""
    if lam != lam:
        return lam
    return min(max(lam, 0.0), 1.0)


class Buffers:
    """Per-searcher device buffers (fixed shapes so that the collectives are fixed-size)."""

    def __init__(self, world: int, p_cap: int, pitch: int, device):
        i64, i32, f32 = torch.int64, torch.int32, torch.float32
        self.local = torch.zeros((p_cap, CAND_WORDS), dtype=i64, device=device)
        self.local_n = torch.zeros(1, dtype=i32, device=device)
        self.gathered = torch.zeros((world, p_cap, CAND_WORDS), dtype=i64, device=device)
        self.pool = torch.zeros((p_cap, CAND_WORDS), dtype=i64, device=device)
        self.pool_n = torch.zeros(1, dtype=i32, device=device)
        self.emb = torch.zeros((p_cap, pitch), dtype=f32, device=device)
        self.sel_pos = torch.zeros(p_cap, dtype=i32, device=device)
        self.sel_n = torch.zeros(1, dtype=i32, device=device)
        self.result = torch.zeros((p_cap, CAND_WORDS), dtype=i64, device=device)
        self.p_cap = p_cap


def stage_lex(lo: int, hi: int, lex_rows, lex_scores, device):
    """src/rag_engine.rs:505-530 for ONE shard [lo, hi): the lexical map that LexicalIndex::score
    returned (global rows, raw BM25 scores; every rank passes the SAME lists) becomes this rank's
    (sorted local rows, score / max_lexical) device tensors.  max_lexical is the GLOBAL maximum
    (fold(0.0, max).max(EPSILON), :511-515), so a row's lexical_score does not depend on the sharding;
    later duplicates of a row win (HashMap collect).  Returns (d_rows int32, d_norm f32, n) or None."""
    import numpy as np
    if lex_rows is None or len(lex_rows) == 0:
        return None
    rows = np.asarray(lex_rows, dtype=np.int64)
    sc = np.asarray(lex_scores, dtype=np.float32)
    max_lex = np.float32(max(np.float32(0.0), sc.max()))
    max_lex = max(max_lex, np.finfo(np.float32).eps)
    last = {}
    for i, r in enumerate(rows.tolist()):
        if lo <= r < hi:
            last[r - lo] = i
    if not last:
        return None
    loc = np.array(sorted(last), dtype=np.uint32)
    norm = (sc[[last[int(r)] for r in loc]] / np.float32(max_lex)).astype(np.float32)     # :527-530, one f32 division
    return (torch.from_numpy(loc.view(np.int32)).to(device), torch.from_numpy(norm).to(device), int(loc.shape[0]))


def sharded_search(backend, group, bufs: Buffers, query, top_k: int, diversity_factor: float,
                   w_embed: float, w_lex: float, lex=None):
    """search_with_diversity (src/rag_engine.rs:717-759, no reranker) over a row-sharded
    corpus.  Returns (result, result_n) device tensors; they are meaningful on rank 0
    (for diversity_factor == 0 on every rank).  `lex`: this rank's `stage_lex(...)` triple (the
    BM25 term of the blend, :505-532) or None for an embedding-only query."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lam = clamp_lambda(float(diversity_factor))
    m = pool_size(top_k, lam)
    if m > bufs.p_cap:
        raise ValueError(f"pool {m} exceeds buffer capacity {bufs.p_cap}")
    if world > 1 and getattr(backend, "mailbox_ready", False):
        # fused exchange: no collective call.  The scan kernel posts this rank's list into rank
        # 0's HBM; rank 0 merges inside a kernel that waits for the flags, then runs MMR with
        # peer loads.  Other ranks are done after their scan.
        if m > backend.mailbox_m_cap:        # checked BEFORE taking a sequence number: ranks must stay in step
            raise ValueError(f"pool {m} exceeds the mailbox capacity {backend.mailbox_m_cap}")
        seq = backend.next_seq()
        backend.topm_post(query, w_embed, w_lex, m, seq, lex)
        if rank != 0:
            return bufs.result, bufs.sel_n
        backend.mailbox_merge(seq, m, bufs.pool[:m], bufs.pool_n)
        if lam == 0.0:
            return bufs.pool[:m], bufs.pool_n
        backend.mmr_peers(bufs.pool[:m], bufs.pool_n, m, top_k, lam, bufs.sel_pos, bufs.sel_n, bufs.result)
        return bufs.result, bufs.sel_n
    local = bufs.local[:m]
    gathered = bufs.gathered[:, :m]
    backend.topm(query, w_embed, w_lex, m, local, bufs.local_n, lex)
    if world > 1:
        g_flat = gathered if gathered.is_contiguous() else None
        if g_flat is not None and dist.get_backend(group) == "nccl":
            dist.all_gather_into_tensor(g_flat.view(world * m, CAND_WORDS), local, group=group)
            lists = g_flat
        else:
            parts = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(parts, local.contiguous(), group=group)
            lists = torch.stack(parts)
        backend.merge(lists, world, m, bufs.pool[:m], bufs.pool_n)
        pool, pool_n = bufs.pool[:m], bufs.pool_n
    else:
        pool, pool_n = local, bufs.local_n
    if lam == 0.0:
        return pool, pool_n
    emb = bufs.emb[:m]
    if world > 1 and getattr(backend, "peers_ready", False):
        # peer-memory path: rank 0's MMR kernels load the pool rows straight from the owning
        # GPUs' HBM over NVLink (CUDA IPC mappings) -- no gather kernel, no second collective
        if rank != 0:
            return bufs.result, bufs.sel_n
        backend.mmr_peers(pool, pool_n, m, top_k, lam, bufs.sel_pos, bufs.sel_n, bufs.result)
        return bufs.result, bufs.sel_n
    if world > 1:
        backend.gather(pool, pool_n, m, emb)
        dist.reduce(emb.view(torch.int32), dst=dist.get_global_rank(group, 0) if group is not None else 0,
                    op=dist.ReduceOp.SUM, group=group)
        if rank != 0:
            return bufs.result, bufs.sel_n
        backend.mmr_matrix(emb, pool, pool_n, m, top_k, lam, bufs.sel_pos, bufs.sel_n, bufs.result)
    else:
        backend.mmr_store(pool, pool_n, m, top_k, lam, bufs.sel_pos, bufs.sel_n, bufs.result)
    return bufs.result, bufs.sel_n


def sharded_search_batch(store, group, queries, m: int, flags: int = 0, device=None):
    """Batched-query path over a row-sharded corpus (BASELINE config 4).  Every rank runs the
    tcgen05 contraction over its shard for the same query batch (rlr_search_batch_device), the
    per-rank [nq, m] key lists are exchanged with ONE all-gather (nq*m*8 B per rank: 0.8 MB for
    1024 x 100) and merged per query on the device (rlr_batch_merge_async).  Returns
    (rows [nq, m] uint32 global, scores [nq, m] f32, n [nq]) as numpy arrays, valid on every
    rank.  `store` is this rank's engine.DeviceStore (needs a binary16 copy)."""
    import numpy as np
    from . import binding as B
    world = dist.get_world_size(group) if group is not None else 1
    q = np.ascontiguousarray(queries, dtype=np.float32)
    nq = q.shape[0]
    dev = device if device is not None else torch.device("cuda", store.info().device)
    stream = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else None   # cpu: the gloo tests' stand-in store
    local = torch.empty((nq, m), dtype=torch.int64, device=dev)
    store.search_batch_device(q, m, local, None, stream, flags)
    if world > 1:
        gathered = torch.empty((world, nq, m), dtype=torch.int64, device=dev)
        if dist.get_backend(group) == "nccl":
            dist.all_gather_into_tensor(gathered.view(world * nq, m), local, group=group)
        else:
            parts = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(parts, local, group=group)
            gathered = torch.stack(parts)
        merged = torch.empty((nq, m), dtype=torch.int64, device=dev)
        cnt = torch.empty(nq, dtype=torch.int32, device=dev)
        store.batch_merge(gathered, world, nq, m, merged, cnt, stream)
    else:
        merged = local
        cnt = (local != 0).sum(dim=1).to(torch.int32)
    keys = merged.cpu().numpy().view(np.uint64)
    n = cnt.cpu().numpy().astype(np.uint32)
    return B.key_row(keys), B.key_score(keys), n


class CudaBackend:
    """The product backend: every step is a kernel launch from librlr_b200.so on the
    current torch CUDA stream.  No fallback."""

    def __init__(self, store, device: Optional[torch.device] = None, search_flags: int = 0):
        from . import binding as B
        self.B = B
        self.lib = B.load()
        self.store = store
        info = store.info()
        self.pitch, self.dim = info.pitch, info.dim
        self.device = device if device is not None else torch.device("cuda", info.device)
        self.ctx = C.c_void_p()
        B.check(self.lib.rlr_ctx_create(store.handle, C.byref(self.ctx)))
        if search_flags:
            B.check(self.lib.rlr_ctx_set_flags(self.ctx, search_flags))
        self.search_flags = search_flags
        self.peer_set = None
        self.peers_ready = False
        self.mailbox = None
        self.mailbox_ready = False
        self.mailbox_m_cap = 0
        self.mailbox_ring = 0
        self._seq = 0
        self._rank = 0

    def open_mailbox(self, group, m_cap: int = 1024, ring: int = 4):
        """Rank 0 allocates the mailbox, every other rank maps it (CUDA IPC).  Collective; raises on every
        rank if any rank failed.  Needs open_peers() as well (rank 0's MMR reads the pool rows from peer HBM)."""
        import numpy as np
        B = self.B
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self._rank = rank
        mb = C.c_void_p()
        handle = np.zeros(B.RLR_IPC_HANDLE_BYTES, np.uint8)
        err = None
        if rank == 0:
            try:
                B.check(self.lib.rlr_mailbox_create(self.device.index, world, m_cap, ring, C.byref(mb)))
                B.check(self.lib.rlr_mailbox_ipc_export(mb, B.ptr(handle)))
            except Exception as e:      # noqa: BLE001 -- reported collectively below
                err = str(e)
        box = [handle.tobytes() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if rank != 0:
            try:
                h = np.frombuffer(box[0], np.uint8).copy()
                B.check(self.lib.rlr_mailbox_open(self.device.index, B.ptr(h), world, m_cap, ring, C.byref(mb)))
            except Exception as e:      # noqa: BLE001
                err = str(e)
        self.mailbox = mb if mb else None
        self.mailbox_m_cap = m_cap
        self.mailbox_ring = ring
        self._agree(group, err)
        self.mailbox_ready = True

    # ---- lanes: several queries in flight on one rank (one ctx = one workspace per lane; the caller
    # runs each lane on its own CUDA stream, e.g. `with torch.cuda.stream(s): sharded_search(...)`).
    # While the last CTA of one scan merges its per-CTA lists and rank 0 runs merge + MMR, the next
    # query's scan already streams rows on the other SMs.  The mailbox ring must be a multiple of the
    # number of lanes (each slot is then always used by the same lane, in order).
    def add_lane(self) -> int:
        if not hasattr(self, "ctxs"):
            self.ctxs = [self.ctx]
        if self.mailbox_ready and self.mailbox_ring % (len(self.ctxs) + 1) != 0:
            raise ValueError(f"mailbox ring {self.mailbox_ring} is not a multiple of {len(self.ctxs) + 1} lanes: "
                             "a slot must always be used by the same lane")
        c = C.c_void_p()
        self.B.check(self.lib.rlr_ctx_create(self.store.handle, C.byref(c)))
        if self.search_flags:
            self.B.check(self.lib.rlr_ctx_set_flags(c, self.search_flags))
        self.ctxs.append(c)
        return len(self.ctxs) - 1

    def use_lane(self, i: int) -> None:
        self.ctx = self.ctxs[i] if hasattr(self, "ctxs") else self.ctx

    def next_seq(self) -> int:
        self._seq += 1
        return self._seq

    def topm_post(self, query, w_embed, w_lex, m, seq, lex=None):
        lr, ln, nl = (self._p(lex[0]), self._p(lex[1]), lex[2]) if lex is not None else (None, None, 0)
        self.B.check(self.lib.rlr_topm_post_async(self.ctx, self.mailbox, self._rank, seq, self._p(query), w_embed, w_lex,
                                                  lr, ln, nl, m, self._stream()))

    def mailbox_merge(self, seq, m, out, out_n):
        self.B.check(self.lib.rlr_mailbox_merge_async(self.ctx, self.mailbox, seq, m, self._p(out), self._p(out_n),
                                                      self._stream()))

    def mailbox_status(self) -> int:
        v = C.c_uint32(0)
        self.B.check(self.lib.rlr_mailbox_status(self.mailbox, C.byref(v)))
        return v.value

    def check_mailbox(self) -> None:
        """Raise if any in-kernel mailbox wait of this rank ever timed out (the affected queries returned
        EMPTY results; see scan_topm.cu / merge.cu).  Synchronises the device: call it when reading results."""
        if self.mailbox_ready:
            st = self.mailbox_status()
            if st:
                raise self.B.RlrError(self.B.RLR_ERR_CUDA, f"a mailbox wait timed out on this rank (status {st}): "
                                      "a peer rank is dead or out of step; results since then are empty")

    def _agree(self, group, err: Optional[str]) -> None:
        """Collective: every rank learns whether ANY rank failed its local step, and all raise together
        (a rank that raised alone would leave the others waiting in the next collective)."""
        objs = [None] * dist.get_world_size(group)
        dist.all_gather_object(objs, err, group=group)
        bad = [(r, e) for r, e in enumerate(objs) if e]
        if bad:
            raise self.B.RlrError(self.B.RLR_ERR_CUDA, "; ".join(f"rank {r}: {e}" for r, e in bad))

    def open_peers(self, group, plan: "ShardPlan"):
        """Exchange CUDA-IPC handles of the shards; rank 0 maps every peer shard.  Collective; raises
        on every rank if the mapping failed on any."""
        import numpy as np
        B = self.B
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        handle = np.zeros(64, np.uint8)
        err = None
        try:
            B.check(self.lib.rlr_store_ipc_export(self.store.handle, self.search_flags, B.ptr(handle)))
        except Exception as e:          # noqa: BLE001 -- reported collectively below
            err = str(e)
        info = self.store.info()
        mine = (handle.tobytes(), int(info.row_base), int(info.n_rows))
        allv = [None] * world
        dist.all_gather_object(allv, mine, group=group)
        if rank == 0 and err is None:
            try:
                handles = np.frombuffer(b"".join(v[0] for v in allv), np.uint8).copy()
                row_base = np.array([v[1] for v in allv], np.uint64)
                n_rows = np.array([v[2] for v in allv], np.uint64)
                ps = C.c_void_p()
                B.check(self.lib.rlr_peer_set_open(self.store.handle, 0, world, B.ptr(handles), B.ptr(row_base),
                                                   B.ptr(n_rows), self.search_flags, C.byref(ps)))
                self.peer_set = ps
            except Exception as e:      # noqa: BLE001
                err = str(e)
        self._agree(group, err)
        self.peers_ready = True

    def mmr_peers(self, pool, pool_n, p_cap, top_k, lam, sel_pos, sel_n, result):
        self.B.check(self.lib.rlr_mmr_peers_async(self.ctx, self.peer_set, self._p(pool), self._p(pool_n), p_cap, top_k,
                                                  lam, self._p(sel_pos), self._p(sel_n), self._p(result), self._stream()))

    def close(self, group=None):
        """With `group` (collective): mappings of other ranks' memory are closed before the
        owners free it."""
        if group is not None and dist.get_world_size(group) > 1:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=group)
            if self.mailbox and self._rank != 0:
                self.lib.rlr_mailbox_close(self.mailbox)
                self.mailbox = None
            if self.peer_set:
                self.lib.rlr_peer_set_close(self.peer_set)
                self.peer_set = None
            dist.barrier(group=group)
        if self.mailbox:
            self.lib.rlr_mailbox_close(self.mailbox)
            self.mailbox = None
        self.mailbox_ready = False
        self.peers_ready = False
        if self.peer_set:
            self.lib.rlr_peer_set_close(self.peer_set)
            self.peer_set = None
        for c in list(getattr(self, "ctxs", [self.ctx])) + list(getattr(self, "_extra_ctxs", [])):
            if c:
                self.lib.rlr_ctx_destroy(c)
        self.ctxs = []
        self._extra_ctxs = []
        self._multi_ctxs = {}
        self.ctx = None

    @staticmethod
    def _p(t: torch.Tensor):
        return C.c_void_p(t.data_ptr())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def launches(self) -> int:
        total = 0
        for c in getattr(self, "ctxs", [self.ctx]):
            n = C.c_uint64(0)
            self.B.check(self.lib.rlr_ctx_launch_count(c, C.byref(n)))
            total += n.value
        return total

    def topm(self, query, w_embed, w_lex, m, out, out_n, lex=None):
        lr, ln, nl = (self._p(lex[0]), self._p(lex[1]), lex[2]) if lex is not None else (None, None, 0)
        self.B.check(self.lib.rlr_topm_async(self.ctx, self._p(query), w_embed, w_lex, lr, ln, nl, m,
                                             self._p(out), self._p(out_n), self._stream()))

    def merge(self, lists, n_lists, m, out, out_n):
        self.B.check(self.lib.rlr_merge_async(self.ctx, self._p(lists), n_lists, m, self._p(out), self._p(out_n),
                                              self._stream()))

    def gather(self, pool, pool_n, m, emb):
        self.B.check(self.lib.rlr_gather_async(self.ctx, self._p(pool), self._p(pool_n), m, self._p(emb),
                                               self._stream()))

    def mmr_matrix(self, emb, pool, pool_n, p_cap, top_k, lam, sel_pos, sel_n, result):
        self.B.check(self.lib.rlr_mmr_async(self.ctx, self._p(emb), emb.stride(0), self.dim, self._p(pool),
                                            self._p(pool_n), p_cap, top_k, lam, self._p(sel_pos), self._p(sel_n),
                                            self._p(result), self._stream()))

    def mmr_store(self, pool, pool_n, p_cap, top_k, lam, sel_pos, sel_n, result):
        # single GPU: candidates are read straight from the store (no gather)
        self.B.check(self.lib.rlr_mmr_store_async(self.ctx, self._p(pool), self._p(pool_n), p_cap, top_k, lam,
                                                  self._p(sel_pos), self._p(sel_n), self._p(result), self._stream()))

    def search_mmr_multi(self, queries, top_k, diversity_factor, w_embed, w_lex, results, result_ns):
        """throughput mode (rlr_search_mmr_multi_async): len(queries) <= RLR_MAX_MULTI queries, ONE pass over the rows.
        Each lane keeps its own set of per-query ctxs (pool / MMR buffers); query 0 uses the lane's ctx."""
        nq = len(queries)
        key = id(self.ctx) if not isinstance(self.ctx, C.c_void_p) else self.ctx.value
        sets = self.__dict__.setdefault("_multi_ctxs", {})
        cs = sets.setdefault(key, [self.ctx])
        while len(cs) < nq:
            c = C.c_void_p()
            self.B.check(self.lib.rlr_ctx_create(self.store.handle, C.byref(c)))
            if self.search_flags:
                self.B.check(self.lib.rlr_ctx_set_flags(c, self.search_flags))
            cs.append(c)
            self.__dict__.setdefault("_extra_ctxs", []).append(c)
        arr = lambda ptrs: (C.c_void_p * nq)(*ptrs)       # noqa: E731
        self.B.check(self.lib.rlr_search_mmr_multi_async(arr([c.value for c in cs[:nq]]), nq,
                                                         arr([q.data_ptr() for q in queries]), top_k, diversity_factor,
                                                         w_embed, w_lex, arr([r.data_ptr() for r in results]),
                                                         arr([r.data_ptr() for r in result_ns]), self._stream()))

    def search_mmr(self, query, top_k, diversity_factor, w_embed, w_lex, result, result_n):
        """fused single-GPU path (rlr_search_mmr_async)."""
        self.B.check(self.lib.rlr_search_mmr_async(self.ctx, self._p(query), top_k, diversity_factor, w_embed, w_lex,
                                                   self._p(result), self._p(result_n), self._stream()))


def decode_result(result: torch.Tensor, n: int):
    """(rows, score, emb, lex) numpy arrays from rlr_cand records held as int64 pairs."""
    import numpy as np
    from . import binding as B
    a = result[:n].cpu().numpy().reshape(-1).view(B.CAND_DTYPE)
    return B.key_row(a["key"]), B.key_score(a["key"]), a["emb"].copy(), a["lex"].copy()
