"""rlr_cluster_*: ONE process driving a row-sharded store on several GPUs through the C ABI
(the reference is one process, /root/reference/src/main.rs:140-167).  Bar: every result is
bit-identical to the CPU oracle's unsharded answer -- rows, blended scores, embedding and
lexical scores, MMR selections.

Two layers:
  * several shards on ONE GPU (`devices=[0, 0, 0]`): everything except the NVLink hop --
    shard plan, lexical routing, the mailbox protocol between per-shard streams, the merge
    that waits in-kernel, peer-table MMR, leases, error paths.  Runs on any GPU box.
  * one shard per GPU (`devices=[0, 1, ...]`) when the box has >= 2 GPUs: the same cases over
    real peer memory (cudaDeviceEnablePeerAccess, st.release.sys / ld.acquire.sys over NVLink).
    Skipped WITH A REASON on a 1-GPU box."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F32 = np.float32


def W(e=0.7, l=0.3):
    from rust_local_rag_b200.engine import ResolvedWeights
    return ResolvedWeights(F32(e), F32(l), F32(0.7), F32(0.3))


def same(a, b):
    return np.asarray(a).tobytes() == np.asarray(b).tobytes()


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _device_sets():
    """[0,0,0] always; one-shard-per-GPU sets when the box has them."""
    sets = [pytest.param([0, 0, 0], id="3-shards-on-gpu0")]
    for g in (2, 4, 8):
        sets.append(pytest.param(list(range(g)), id=f"{g}-gpus",
                                 marks=pytest.mark.skipif(_ngpu() < g, reason=f"needs {g} GPUs on the box (has {_ngpu()})")))
    return sets


@pytest.fixture(scope="module")
def eng():
    from rust_local_rag_b200 import engine
    return engine


@pytest.fixture(scope="module")
def corpus(orc):
    n, dim = 60011, 768
    rows = orc.synth_rows(n, dim, kind=1, n_clusters=64)
    qs = orc.synth_rows(6, dim, kind=1, seed=0x5EED0002, n_clusters=64)
    return rows, qs


@pytest.mark.parametrize("devices", _device_sets())
def test_cluster_search_with_diversity_matches_oracle(eng, orc, corpus, devices):
    rows, qs = corpus
    cl = eng.ClusterStore.from_rows(rows, devices=devices)
    ci = cl.cluster_info()
    assert ci.n_shards == len(devices) and sum(ci.shard_rows[:ci.n_shards]) == len(rows)
    assert ci.shard_rows[0] < ci.shard_rows[1]            # tail-balanced default plan: the root owns fewer rows
    for q in qs[:3]:
        for k, lam in ((100, 0.7), (5, 0.3), (0, 0.5), (1, 1.0), (7, 0.0), (100, 0.0)):
            got = cl.search_mmr(q, k, lam, W())
            ref = orc.search_with_diversity(rows, q, k, lam, threads=4)
            for a, b in zip(got, ref):
                assert same(a, b), (devices, k, lam)
    # 40 queries back to back: every mailbox slot is reused many times
    for i in range(40):
        q = qs[i % len(qs)]
        got = cl.search_mmr(q, 100, 0.7, W())
        ref = orc.search_with_diversity(rows, q, 100, 0.7, threads=4)
        assert same(got[0], ref[0]) and same(got[1], ref[1])
    assert cl.launches() > 0
    cl.close()


@pytest.mark.parametrize("devices", _device_sets())
def test_cluster_topm_candidates_and_lexical_routing(eng, orc, corpus, devices):
    rows, qs = corpus
    n = len(rows)
    # explicit, very uneven plan
    g = len(devices)
    plan = [1000] + [(n - 1000) // (g - 1)] * (g - 1)
    plan[-1] += n - sum(plan)
    cl = eng.ClusterStore.from_rows(rows, devices=devices, shard_rows=plan)
    rng = np.random.default_rng(3)
    lex_rows = rng.choice(n, 1500, replace=False).astype(np.uint32)
    lex_rows[:4] = [0, 999, 1000, n - 1]                  # shard boundaries
    lex_scores = (rng.random(1500) * 9).astype(F32)
    q = qs[0]
    for m in (15, 300, 900, 1024):
        got = cl.search_topm(q, m, W(), lex_rows, lex_scores)
        ref = orc.search(rows, q, m, lex_rows=lex_rows, lex_scores=lex_scores, threads=4)
        for a, b in zip(got, ref):
            assert same(a, b), m
    assert (got[3] > 0).any()                             # the BM25 term reached the result
    got = cl.search_mmr(q, 100, 0.7, W(), lex_rows, lex_scores)
    ref = orc.search_with_diversity(rows, q, 100, 0.7, lex_rows=lex_rows, lex_scores=lex_scores, threads=4)
    for a, b in zip(got, ref):
        assert same(a, b)
    # get_embedding_candidates (:415-461)
    r, s = cl.embedding_candidates(q, 40)
    R, S = orc.embedding_candidates(rows, q, 40, threads=4)
    assert same(r, R) and same(s, S)
    # reranker-in-the-middle (N3): search(3P) -> host blend -> mmr with caller-supplied relevance over GLOBAL rows
    cr, cc, ce, cx = cl.search_topm(q, 45, W())
    rel = (F32(0.3) * cc + F32(0.7) * rng.random(len(cc)).astype(F32)).astype(F32)
    order = np.argsort(-rel, kind="stable")
    cr, rel = cr[order], rel[order]
    sel = cl.mmr(cr, rel, 5, 0.3)
    ref_sel = orc.mmr(rows[cr], rel, 5, 0.3)
    assert same(sel, ref_sel)
    # rows come back from the owning shard
    pick = np.array([0, 999, 1000, n // 2, n - 1], np.uint32)
    assert same(cl.read_rows(pick), rows[pick])
    cl.close()


def test_cluster_f16_copy_and_tiny_store(eng, orc):
    from rust_local_rag_b200 import binding as B
    rng = np.random.default_rng(11)
    rows = orc.normalize_rows(rng.standard_normal((9000, 384)).astype(F32))
    q = rng.standard_normal(384).astype(F32)
    cl = eng.ClusterStore.from_rows(rows, devices=[0, 0], flags=B.RLR_STORE_KEEP_F16)
    got = cl.search_mmr(q, 20, 0.5, W(), flags=B.RLR_SEARCH_F16)
    ref = orc.search_with_diversity(rows.astype(np.float16).astype(F32), q, 20, 0.5)
    for a, b in zip(got, ref):
        assert same(a, b)
    cl.close()
    # fewer rows than devices: the cluster quietly uses fewer shards (every shard must own rows)
    tiny = rows[:2]
    cl = eng.ClusterStore.from_rows(tiny, devices=[0, 0, 0])
    assert cl.cluster_info().n_shards == 2
    got = cl.search_mmr(q, 5, 0.3, W())
    ref = orc.search_with_diversity(tiny, q, 5, 0.3)
    for a, b in zip(got, ref):
        assert same(a, b)
    cl.close()
    cl = eng.ClusterStore.from_rows(np.zeros((0, 384), F32), devices=[0, 0])
    assert len(cl.search_mmr(q, 5, 0.3, W())[0]) == 0      # empty store => Ok(vec![]) (:476-478)
    cl.close()


def test_cluster_errors(eng, rlr):
    rows = np.eye(8, 64, dtype=F32)
    with pytest.raises(rlr.RlrError) as ei:
        eng.ClusterStore.from_rows(rows, devices=[0, 0], shard_rows=[3, 4])
    assert ei.value.code == rlr.RLR_ERR_INVALID_ARG
    cl = eng.ClusterStore.from_rows(rows, devices=[0, 0])
    with pytest.raises(rlr.RlrError) as ei:
        cl.search_mmr(np.ones(63, F32), 5, 0.3, W())
    assert ei.value.code == rlr.RLR_ERR_DIM_MISMATCH
    bad = np.ones(64, F32); bad[5] = np.nan
    with pytest.raises(rlr.RlrError) as ei:
        cl.search_mmr(bad, 5, 0.3, W())
    assert ei.value.code == rlr.RLR_ERR_NONFINITE
    with pytest.raises(rlr.RlrError):
        cl.append(rows)
    cl.close()


@pytest.mark.parametrize("devices", _device_sets())
def test_cluster_concurrent_searches(eng, orc, corpus, devices):
    """Searches are re-entrant on one cluster like on one store (read lock, src/mcp_server.rs:89,377): each
    caller leases its own per-GPU workspaces, streams and mailbox."""
    rows, qs = corpus
    cl = eng.ClusterStore.from_rows(rows, devices=devices)
    refs = [orc.search_with_diversity(rows, q, 100, 0.7, threads=4) for q in qs]
    errs = []

    def worker(t):
        try:
            for i in range(12):
                j = (t + i) % len(qs)
                got = cl.search_mmr(qs[j], 100, 0.7, W())
                if not (same(got[0], refs[j][0]) and same(got[1], refs[j][1])):
                    errs.append((t, i))
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    cl.close()


@pytest.mark.parametrize("devices", _device_sets())
def test_engine_over_a_cluster(eng, orc, corpus, devices):
    """The host mirror with `devices=[...]`: RagEngine.search_with_diversity / search / search_documents are the
    same calls over a ClusterStore."""
    rows, qs = corpus
    e = eng.RagEngine.from_rows(rows[:20000], normalize=False, devices=devices)
    res = eng.search_documents(e, qs[0], top_k=5, diversity_factor=0.3)
    ref = orc.search_with_diversity(rows[:20000], qs[0], 5, 0.3)
    assert same(np.array([r.row for r in res], np.uint32), ref[0])
    assert same(np.array([r.score for r in res], F32), ref[1])
    res = e.search(qs[1], 30)
    ref = orc.search(rows[:20000], qs[1], 30)
    assert same(np.array([r.row for r in res], np.uint32), ref[0])
    e.store.close()


@pytest.mark.parametrize("devices", _device_sets())
def test_cluster_throughput_mode_multi_query(eng, orc, corpus, devices):
    """rlr_cluster_search_mmr_multi: 2-3 queries, ONE pass over every shard, one mailbox post per (GPU, query)."""
    rows, qs = corpus
    cl = eng.ClusterStore.from_rows(rows, devices=devices)
    refs = [orc.search_with_diversity(rows, q, 100, 0.7, threads=4) for q in qs]
    for it in range(10):                                  # slots 1..3 of the ring are reused call after call
        nq = 2 + it % 2
        idx = [(it + j) % len(qs) for j in range(nq)]
        got = cl.search_mmr_multi(qs[idx], 100, 0.7, W())
        for j, i in enumerate(idx):
            assert same(got[j][0], refs[i][0]) and same(got[j][1], refs[i][1]) and same(got[j][2], refs[i][2]), (it, j)
        one = cl.search_mmr(qs[it % len(qs)], 100, 0.7, W())           # single-query calls interleave on the same lane
        assert same(one[0], refs[it % len(qs)][0])
    for k, lam in ((5, 0.3), (7, 0.0)):
        got = cl.search_mmr_multi(qs[:3], k, lam, W())
        for j in range(3):
            ref = orc.search_with_diversity(rows, qs[j], k, lam, threads=4)
            assert same(got[j][0], ref[0]) and same(got[j][1], ref[1])
    # per-query lexical pairs (a text query each; one query without): routed per shard, each normalised by ITS global max
    rng = np.random.default_rng(31)
    n = len(rows)
    for it in range(4):
        lex = []
        for j in range(3):
            if j == it % 3:
                lex.append(None)
                continue
            cnt = int(rng.integers(41, 1500))
            top = np.argsort(-(rows @ orc.normalize_rows(qs[j:j + 1])[0]))[:40].astype(np.uint32)       # pairs that matter
            rest = np.setdiff1d(rng.choice(n, cnt, replace=False).astype(np.uint32), top)[: cnt - 40]
            lr = np.concatenate([top, rng.permutation(rest)]).astype(np.uint32)                          # unique rows, any order
            lex.append((lr, (rng.random(len(lr)) * 7).astype(F32)))
        got = cl.search_mmr_multi(qs[:3], 100, 0.7, W(), lex=lex)
        blended = False
        for j in range(3):
            lr, ls = lex[j] if lex[j] is not None else (None, None)
            ref = orc.search_with_diversity(rows, qs[j], 100, 0.7, lex_rows=lr, lex_scores=ls, threads=4)
            for a, b in zip(got[j], ref):
                assert same(a, b), (devices, it, j)
            one = cl.search_mmr(qs[j], 100, 0.7, W(), lr, ls)
            assert all(same(a, b) for a, b in zip(one, ref))
            blended |= bool((ref[3] != 0).any())
        assert blended
    cl.close()


@pytest.mark.parametrize("devices", _device_sets())
def test_cluster_batched_contraction_equals_single_store(eng, rlr, orc, corpus, devices):
    """rlr_cluster_search_batch (BASELINE configs[3] from one process): a row's tensor-core score does not depend on
    the shard it is computed in, so the cluster's answer equals the single store's bit for bit in every precision;
    with RLR_BATCH_EXACT_RESCORE it carries the oracle's exact scores."""
    rows, qs = corpus
    q = np.concatenate([qs] * 10)[:50]
    flags_store = rlr.RLR_STORE_KEEP_F16 | rlr.RLR_STORE_KEEP_BF16
    one = eng.DeviceStore.from_rows(rows, flags=flags_store)
    cl = eng.ClusterStore.from_rows(rows, devices=devices, flags=flags_store)
    for pf in (rlr.RLR_BATCH_TF32, rlr.RLR_BATCH_BF16, rlr.RLR_BATCH_F16):
        a = one.search_batch(q, 100, flags=pf)
        b = cl.search_batch(q, 100, flags=pf)
        assert (a[2] == 100).all() and same(a[2], b[2])
        assert same(a[0], b[0]) and same(a[1], b[1]), pf
    r, sc, n = cl.search_batch(q[:6], 40, flags=rlr.RLR_BATCH_EXACT_RESCORE)
    for i in range(6):
        qn = orc.normalize(q[i])
        assert sc[i].tobytes() == np.array([orc.dot(qn, rows[x]) for x in r[i]], F32).tobytes()
    one.close(); cl.close()
