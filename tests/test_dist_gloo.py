"""world_size-2 (and 3) gloo tests of the row-sharded search choreography in
rust_local_rag_b200/dist.py: shard plan, fixed-size all-gather of candidate records, merge,
bit-exact int32 reduce of the pool embeddings, MMR on rank 0.

There is no GPU here, so the per-rank kernels are replaced by a TEST-ONLY backend built on
the CPU oracle (the product backend is CudaBackend and has no CPU fallback).  What is under
test is the host-side logic that is identical on the GPU box: the result on rank 0 must be
bit-identical to the unsharded oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F32 = np.float32


def _ord(score):
    b = (np.asarray(score, F32) + F32(0.0)).view(np.uint32)
    return np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def _records(rows, score, emb, lex, m):
    from rust_local_rag_b200 import binding as B
    rec = np.zeros(m, B.CAND_DTYPE)
    n = len(rows)
    rec["key"][:n] = (_ord(score).astype(np.uint64) << np.uint64(32)) | (~rows.astype(np.uint32)).astype(np.uint64)
    rec["emb"][:n] = emb
    rec["lex"][:n] = lex
    return rec


class OracleBackend:
    """TEST-ONLY stand-in for CudaBackend (same method contract, CPU tensors)."""

    def __init__(self, orc, shard_rows, row0, dim):
        self.orc, self.rows, self.row0, self.dim = orc, shard_rows, row0, dim

    @staticmethod
    def _view(t):
        from rust_local_rag_b200 import binding as B
        return t.numpy().reshape(-1).view(B.CAND_DTYPE)

    def topm(self, query, w_embed, w_lex, m, out, out_n, lex=None):
        q = query.numpy()[:self.dim]
        lr = ls = None
        if lex is not None:
            # `lex` is dist.stage_lex's triple: local rows + scores ALREADY divided by the global max.  The
            # oracle divides by the max of what it is given, so a sentinel (row outside the shard, 1.0) makes
            # that division exact and a no-op.
            lr = np.concatenate([lex[0].numpy().view(np.uint32)[:lex[2]], np.array([0xFFFFFFFF], np.uint32)])
            ls = np.concatenate([lex[1].numpy()[:lex[2]], np.array([1.0], F32)])
        r, s, e, l = self.orc.search(self.rows, q, m, w_embed=w_embed, w_lex=w_lex, lex_rows=lr, lex_scores=ls,
                                     normalize_query=False, full_sort=True)
        self._view(out)[:] = _records(r + np.uint32(self.row0), s, e, l, m)
        out_n[0] = len(r)

    def merge(self, lists, n_lists, m, out, out_n):
        rec = self._view(lists.contiguous())
        rec = rec[rec["key"] != 0]
        rec = rec[np.argsort(rec["key"])[::-1]][:m]
        o = self._view(out)
        o[:] = 0
        o[:len(rec)] = rec
        out_n[0] = len(rec)

    def gather(self, pool, pool_n, m, emb):
        from rust_local_rag_b200 import binding as B
        rec = self._view(pool)
        emb.zero_()
        for i in range(int(pool_n[0])):
            g = int(B.key_row(rec["key"][i:i + 1])[0])
            if self.row0 <= g < self.row0 + len(self.rows):
                emb[i, :self.dim] = torch.from_numpy(self.rows[g - self.row0])

    def mmr_matrix(self, emb, pool, pool_n, p_cap, top_k, lam, sel_pos, sel_n, result):
        from rust_local_rag_b200 import binding as B
        n = int(pool_n[0])
        rec = self._view(pool)
        rel = B.key_score(rec["key"][:n])
        pos = self.orc.mmr(emb.numpy()[:n, :self.dim], rel, top_k, lam)
        sel_n[0] = len(pos)
        self._view(result)[:len(pos)] = rec[pos]

    def mmr_store(self, pool, pool_n, p_cap, top_k, lam, sel_pos, sel_n, result):
        raise AssertionError("world > 1 must not take the single-GPU path")


def _worker(rank, world, port, n, dim, cases, ret, lex_pairs=None):
    sys.path.insert(0, ROOT)
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import dist as rdist
    from oracle import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows = orc.synth_rows(n, dim, kind=1, n_clusters=16, threads=1)
        plan = rdist.ShardPlan(n, world, rank)
        shard = rows[plan.row0:plan.row0 + plan.n_local]
        backend = OracleBackend(orc, shard, plan.row0, dim)
        pitch = (dim + 31) // 32 * 32
        out = []
        for (k, lam) in cases:
            p_cap = max(rdist.pool_size(k, rdist.clamp_lambda(lam)), 1)
            bufs = rdist.Buffers(world, p_cap, pitch, torch.device("cpu"))
            q = torch.zeros(pitch + 64)
            q[:dim] = torch.from_numpy(orc.normalize(orc.synth_rows(1, dim, kind=1, seed=99, n_clusters=16)[0]))
            lex = None
            if lex_pairs is not None:       # the same global BM25 pairs on every rank; each stages its own slice
                lex = rdist.stage_lex(plan.row0, plan.row0 + plan.n_local, lex_pairs[0], lex_pairs[1], torch.device("cpu"))
            res, res_n = rdist.sharded_search(backend, dist.group.WORLD, bufs, q, k, lam, 0.7, 0.3, lex=lex)
            if rank == 0:
                r, s, e, l = rdist.decode_result(res, int(res_n[0]))
                out.append((r.tolist(), s.view(np.uint32).tolist(), e.view(np.uint32).tolist(), l.view(np.uint32).tolist()))
        if rank == 0:
            ret.put(out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_matches_unsharded_oracle(world, orc):
    n, dim = 2003, 96
    cases = [(5, 0.3), (100, 0.7), (5, 0.0), (0, 0.5), (40, 1.0)]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, dim, cases, ret)) for r in range(world)]
    [p.start() for p in procs]
    got = ret.get(timeout=180)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    rows = orc.synth_rows(n, dim, kind=1, n_clusters=16, threads=1)
    q = orc.normalize(orc.synth_rows(1, dim, kind=1, seed=99, n_clusters=16)[0])
    for (k, lam), (r, s, e, _l) in zip(cases, got):
        R, S, E, _ = orc.search_with_diversity(rows, q, k, lam, normalize_query=False, full_sort=True)
        assert r == R.tolist(), (k, lam)
        assert s == S.view(np.uint32).tolist() and e == E.view(np.uint32).tolist()


def test_sharded_search_blends_lexical_scores(orc):
    """The BM25 term of the blend (src/rag_engine.rs:505-532) on the sharded path: every rank stages its slice of
    the same global (row, score) pairs, normalised by the GLOBAL maximum; rank 0's result must equal the
    unsharded oracle's, lexical_score included (the sharded path used to drop the term)."""
    world, n, dim = 2, 2003, 96
    cases = [(5, 0.3), (100, 0.7), (7, 0.0)]
    rng = np.random.default_rng(5)
    lex_rows = rng.choice(n, 60, replace=False).astype(np.uint32)
    lex_rows[7] = lex_rows[3]                                    # a duplicate row: the later entry wins
    lex_scores = (rng.random(60) * 9 + 0.1).astype(F32)
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + 17
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, dim, cases, ret, (lex_rows, lex_scores))) for r in range(world)]
    [p.start() for p in procs]
    got = ret.get(timeout=180)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    rows = orc.synth_rows(n, dim, kind=1, n_clusters=16, threads=1)
    q = orc.normalize(orc.synth_rows(1, dim, kind=1, seed=99, n_clusters=16)[0])
    # oracle input without the duplicate (HashMap collect keeps the last value of a key)
    keep = [i for i in range(60) if i != 3]
    blended = False
    for (k, lam), (r, s, e, l) in zip(cases, got):
        R, S, E, L = orc.search_with_diversity(rows, q, k, lam, lex_rows=lex_rows[keep], lex_scores=lex_scores[keep],
                                               normalize_query=False, full_sort=True)
        assert r == R.tolist(), (k, lam)
        assert s == S.view(np.uint32).tolist() and e == E.view(np.uint32).tolist() and l == L.view(np.uint32).tolist()
        blended |= bool((L != 0).any())
    assert blended, "the lexical term never reached a result: the test does not exercise the blend"


def test_shard_plan_covers_rows_exactly():
    sys.path.insert(0, ROOT)
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200.dist import ShardPlan, pool_size, clamp_lambda
    for n in (0, 1, 7, 10_000_000, 100_000_001):
        for g in (1, 2, 3, 4, 8):
            plans = [ShardPlan(n, g, r) for r in range(g)]
            assert plans[0].row0 == 0 and sum(p.n_local for p in plans) == n
            for a, b in zip(plans, plans[1:]):
                assert a.row0 + a.n_local == b.row0
            if n >= g:
                assert plans[-1].owner(n - 1) == g - 1 and plans[0].owner(0) == 0
            if n:
                o = plans[0].owner(n // 2)
                assert plans[o].row0 <= n // 2 < plans[o].row0 + plans[o].n_local
    assert pool_size(5, 0.3) == 15 and pool_size(100, 0.7) == 300 and pool_size(0, 0.5) == 10   # :734
    assert pool_size(5, 0.0) == 5 and pool_size(0, 0.0) == 1                                    # :728, :490
    assert clamp_lambda(-1.0) == 0.0 and clamp_lambda(7.0) == 1.0 and clamp_lambda(float("nan")) != clamp_lambda(float("nan"))


def test_key_codec_roundtrip():
    sys.path.insert(0, ROOT)
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import binding as B
    s = np.array([0.0, -0.0, 1.0, -1.0, 1e-30, -1e-30, 0.5133, 3e38, -3e38], F32)
    rows = np.arange(len(s), dtype=np.uint32) * 1000
    keys = (_ord(s).astype(np.uint64) << np.uint64(32)) | (~rows).astype(np.uint64)
    assert (B.key_row(keys) == rows).all()
    back = B.key_score(keys)
    assert back.tobytes() == (s + F32(0.0)).tobytes()           # -0.0 is folded onto +0.0
    order = np.argsort(keys)[::-1]
    assert (np.diff(s[order].astype(np.float64)) <= 0).all()    # key order == score order


class OracleBatchStore:
    """TEST-ONLY stand-in for engine.DeviceStore on the batched path (CPU tensors): per-query key lists
    from the oracle's exact scan of this rank's shard, and the per-query merge."""

    def __init__(self, orc, shard_rows, row0):
        self.orc, self.rows, self.row0 = orc, shard_rows, row0

    def search_batch_device(self, queries, m, d_keys, d_cnt=None, stream=None, flags=0):
        out = d_keys.numpy().view(np.uint64)
        out[:] = 0
        for q in range(queries.shape[0]):
            r, s = self.orc.embedding_candidates(self.rows, queries[q], m, normalize_query=False)
            out[q, :len(r)] = (_ord(s).astype(np.uint64) << np.uint64(32)) | (~(r + np.uint32(self.row0))).astype(np.uint32).astype(np.uint64)

    def batch_merge(self, d_lists, n_lists, nq, m, d_out, d_out_cnt=None, stream=None):
        lists = d_lists.numpy().view(np.uint64)
        out = d_out.numpy().view(np.uint64)
        out[:] = 0
        for q in range(nq):
            k = lists[:, q, :].reshape(-1)
            k = np.sort(k[k != 0])[::-1][:m]
            out[q, :len(k)] = k
            if d_out_cnt is not None:
                d_out_cnt[q] = len(k)


def _batch_worker(rank, world, port, n, dim, nq, m, ret):
    sys.path.insert(0, ROOT)
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import dist as rdist
    from oracle import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows = orc.synth_rows(n, dim, kind=1, n_clusters=16, threads=1)
        plan = rdist.ShardPlan(n, world, rank, head_rows=n // (3 * world))       # uneven shards
        store = OracleBatchStore(orc, rows[plan.row0:plan.row0 + plan.n_local], plan.row0)
        qs = orc.synth_rows(nq, dim, kind=1, seed=7, n_clusters=16, threads=1)
        r, s, cnt = rdist.sharded_search_batch(store, dist.group.WORLD, qs, m, 0, torch.device("cpu"))
        if rank == world - 1:                                                    # valid on every rank
            ret.put((r.tolist(), s.view(np.uint32).tolist(), cnt.tolist()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_batch_matches_unsharded_oracle(orc):
    world, n, dim, nq, m = 2, 1501, 64, 9, 40
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + 11
    procs = [ctx.Process(target=_batch_worker, args=(r, world, port, n, dim, nq, m, ret)) for r in range(world)]
    [p.start() for p in procs]
    rows_g, scores_g, cnt = ret.get(timeout=180)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    rows = orc.synth_rows(n, dim, kind=1, n_clusters=16, threads=1)
    qs = orc.synth_rows(nq, dim, kind=1, seed=7, n_clusters=16, threads=1)
    for q in range(nq):
        R, S = orc.embedding_candidates(rows, qs[q], m, normalize_query=False)
        assert cnt[q] == m and rows_g[q] == R.tolist() and scores_g[q] == S.view(np.uint32).tolist()


def test_shard_plan_with_head_rows():
    from rust_local_rag_b200.dist import ShardPlan
    for n, g, head in ((10_000_000, 8, 992_666), (1000, 3, 10), (7, 2, 0), (100, 4, None), (100, 1, 5)):
        plans = [ShardPlan(n, g, r, head_rows=head) for r in range(g)]
        assert plans[0].row0 == 0 and sum(p.n_local for p in plans) == n
        for a, b in zip(plans, plans[1:]):
            assert a.row0 + a.n_local == b.row0
        if head is not None and g > 1:
            assert plans[0].n_local == head
            rest = [p.n_local for p in plans[1:]]
            assert max(rest) - min(rest) <= 1
        for row in (0, n // 2, n - 1):
            o = plans[0].owner(row)
            assert plans[o].row0 <= row < plans[o].row0 + plans[o].n_local
    # tail-balanced: scan(rank 0) + tail == scan(others)
    h = ShardPlan.balanced_head_rows(10_000_000, 8, tail_rows=300_000)
    other = (10_000_000 - h) / 7
    assert abs((h + 300_000) - other) < 2
