"""Fused exchange on ONE GPU: several row shards of one corpus live on the same device and play
the ranks of an NVSwitch box.  Each "rank" has its own store, ctx and CUDA stream and posts its
scan's list into one mailbox (rlr_topm_post_async); the root merges inside the waiting kernel
(rlr_mailbox_merge_async).  The IPC mapping itself needs >= 2 processes/GPUs and is covered by
tests/test_gpu_dist.py; everything else (kernel-side post, flags, slot ring, flow control,
rank-counting merge) is the same code and is checked here bit-exactly against the oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F32 = np.float32


class Rig:
    def __init__(self, rows, bounds, m_cap=1024, ring=4):
        import torch
        from rust_local_rag_b200 import binding as B, engine
        self.B, self.torch, self.lib = B, torch, B.load()
        self.dev = torch.device("cuda", 0)
        self.stores, self.ctxs, self.streams = [], [], []
        for lo, hi in bounds:
            s = engine.DeviceStore.from_rows(rows[lo:hi], row_base=lo)
            ctx = C.c_void_p()
            B.check(self.lib.rlr_ctx_create(s.handle, C.byref(ctx)))
            self.stores.append(s); self.ctxs.append(ctx); self.streams.append(torch.cuda.Stream(self.dev))
        self.n = len(bounds)
        self.mb = C.c_void_p()
        B.check(self.lib.rlr_mailbox_create(0, self.n, m_cap, ring, C.byref(self.mb)))
        self.q = torch.zeros(B.RLR_MAX_DIM + 64, device=self.dev)
        self.out = torch.zeros((m_cap, 2), dtype=torch.int64, device=self.dev)
        self.out_n = torch.zeros(1, dtype=torch.int32, device=self.dev)

    def set_query(self, q):
        self.q.zero_()
        self.q[:len(q)] = self.torch.from_numpy(np.ascontiguousarray(q)).to(self.dev)
        self.torch.cuda.synchronize()

    def post(self, r, seq, m, we=0.7, wl=0.3):
        self.B.check(self.lib.rlr_topm_post_async(self.ctxs[r], self.mb, r, seq, C.c_void_p(self.q.data_ptr()), we, wl,
                                                  None, None, 0, m, C.c_void_p(self.streams[r].cuda_stream)))

    def merge(self, seq, m):
        self.B.check(self.lib.rlr_mailbox_merge_async(self.ctxs[0], self.mb, seq, m, C.c_void_p(self.out.data_ptr()),
                                                      C.c_void_p(self.out_n.data_ptr()),
                                                      C.c_void_p(self.streams[0].cuda_stream)))

    def result(self):
        # wait for the ROOT's stream only: other "ranks" may legitimately still be held by the ring
        from rust_local_rag_b200 import dist as rdist
        self.streams[0].synchronize()
        return rdist.decode_result(self.out, int(self.out_n.item()))

    def status(self):
        v = C.c_uint32(7)
        self.B.check(self.lib.rlr_mailbox_status(self.mb, C.byref(v)))
        return v.value

    def close(self):
        self.torch.cuda.synchronize()
        self.lib.rlr_mailbox_close(self.mb)
        for c in self.ctxs:
            self.lib.rlr_ctx_destroy(c)
        for s in self.stores:
            s.close()


def _corpus(orc, n, dim, seed):
    rng = np.random.default_rng(seed)
    rows = orc.normalize_rows(rng.standard_normal((n, dim)).astype(F32))
    qs = orc.normalize_rows(rng.standard_normal((12, dim)).astype(F32))
    return rows, qs


@pytest.mark.parametrize("n,dim,bounds,m", [
    (9000, 768, [(0, 4500), (4500, 9000)], 300),
    (20001, 384, [(0, 1000), (1000, 9000), (9000, 20001)], 900),      # uneven (tail-balanced style) shards
    (700, 64, [(0, 100), (100, 250), (250, 500), (500, 700)], 300),   # lists shorter than m on every rank
    (5000, 1024, [(0, 5000)], 45),                                    # a single rank posting to itself
])
def test_post_and_merge_equals_single_store_and_oracle(orc, n, dim, bounds, m):
    rows, qs = _corpus(orc, n, dim, n + dim)
    rig = Rig(rows, bounds)
    seq = 0
    for qi in range(6):                      # more queries than ring slots: slots are reused
        rig.set_query(qs[qi])
        seq += 1
        for r in range(rig.n):
            rig.post(r, seq, m)
        rig.merge(seq, m)
        got_rows, got_score, got_emb, _ = rig.result()
        R, S, E, _ = orc.search(rows, qs[qi], m, normalize_query=False, full_sort=True)
        assert got_rows.tobytes() == np.asarray(R).tobytes()
        assert got_score.tobytes() == np.asarray(S).tobytes()
        assert got_emb.tobytes() == np.asarray(E).tobytes()
    assert rig.status() == 0
    rig.close()


def test_ranks_running_ahead_are_held_by_the_slot_ring(orc):
    """Rank 1 enqueues all its scans before the root enqueues anything: with ring=2 its third
    post must wait inside the kernel until the root has merged query 1, and so on.  Results
    stay exact and no wait times out."""
    n, dim, m, nq = 6000, 256, 300, 8
    rows, qs = _corpus(orc, n, dim, 99)
    rig = Rig(rows, [(0, 2500), (2500, 6000)], ring=2)
    rig.set_query(qs[0])                     # the same query every time: posts may run ahead freely
    for seq in range(1, nq + 1):
        rig.post(1, seq, m)
    R, S, E, _ = orc.search(rows, qs[0], m, normalize_query=False, full_sort=True)
    for seq in range(1, nq + 1):
        rig.post(0, seq, m)
        rig.merge(seq, m)
        got_rows, got_score, _, _ = rig.result()
        assert got_rows.tobytes() == np.asarray(R).tobytes() and got_score.tobytes() == np.asarray(S).tobytes()
    assert rig.status() == 0
    rig.close()


def test_mailbox_argument_errors(rlr):
    B, lib = rlr, rlr.load()
    mb = C.c_void_p()
    assert lib.rlr_mailbox_create(0, 0, 300, 4, C.byref(mb)) == B.RLR_ERR_INVALID_ARG
    assert lib.rlr_mailbox_create(0, 2, 5000, 4, C.byref(mb)) == B.RLR_ERR_UNSUPPORTED
    assert lib.rlr_mailbox_create(0, 2, 300, 1, C.byref(mb)) == B.RLR_ERR_INVALID_ARG
    B.check(lib.rlr_mailbox_create(0, 2, 300, 4, C.byref(mb)))
    from rust_local_rag_b200 import engine
    s = engine.DeviceStore.from_rows(np.eye(8, dtype=F32))
    ctx = C.c_void_p()
    B.check(lib.rlr_ctx_create(s.handle, C.byref(ctx)))
    import torch
    q = torch.zeros(B.RLR_MAX_DIM + 64, device="cuda")
    qp = C.c_void_p(q.data_ptr())
    assert lib.rlr_topm_post_async(ctx, mb, 0, 0, qp, 0.7, 0.3, None, None, 0, 5, None) == B.RLR_ERR_INVALID_ARG   # seq 0
    assert lib.rlr_topm_post_async(ctx, mb, 2, 1, qp, 0.7, 0.3, None, None, 0, 5, None) == B.RLR_ERR_INVALID_ARG   # rank
    assert lib.rlr_topm_post_async(ctx, mb, 0, 1, qp, 0.7, 0.3, None, None, 0, 301, None) == B.RLR_ERR_UNSUPPORTED  # m
    lib.rlr_ctx_destroy(ctx)
    lib.rlr_mailbox_close(mb)
    s.close()


def test_a_rank_that_never_delivers_yields_an_empty_result_and_a_sticky_status(orc):
    """a mailbox timeout must not degrade into a wrong answer.  Rank 1 never posts query 1: after the 4 s
    in-kernel timeout the root's merge delivers an EMPTY result (n = 0, zero keys) and sets the status word -- it
    does not merge whatever the slot happens to hold -- and a later, complete query is answered correctly again."""
    n, dim, m = 9000, 96, 40
    rows = orc.synth_rows(n, dim, kind=1, n_clusters=16)
    q = orc.normalize(orc.synth_rows(1, dim, kind=1, seed=5, n_clusters=16)[0])
    rig = Rig(rows, [(0, 4000), (4000, n)], m_cap=64, ring=2)
    rig.set_query(q)
    rig.post(0, 1, m)                      # rank 0 posts, rank 1 stays silent
    rig.merge(1, m)
    r, s, e, l = rig.result()              # returns after the timeout
    assert len(r) == 0 and rig.status() == 2
    assert (rig.out[:m, 0] == 0).all().item()
    for rk in (0, 1):                      # query 2 is complete: the mailbox works again, the status stays set
        rig.post(rk, 2, m)
    rig.merge(2, m)
    r, s, e, l = rig.result()
    R, S, E, L = orc.search(rows, q, m, normalize_query=False, full_sort=True)
    assert r.tobytes() == R.tobytes() and s.tobytes() == S.tobytes()
    assert rig.status() == 2
    rig.close()
