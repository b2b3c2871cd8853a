"""Throughput mode (rlr_search_mmr_multi): up to 3 queries answered by ONE pass over the rows -- one group of consumer
warps per query over the same shared-memory tiles (scan_topm.cu, "QUERY GROUPS").  Concurrent searches under the
reference's read lock (/root/reference/src/mcp_server.rs:89,377) are what it models.  Bar: every query's result is
bit-identical to its single-query result and to the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F32 = np.float32


def W(e=0.7, l=0.3):
    from rust_local_rag_b200.engine import ResolvedWeights
    return ResolvedWeights(F32(e), F32(l), F32(0.7), F32(0.3))


def same(a, b):
    return np.asarray(a).tobytes() == np.asarray(b).tobytes()


@pytest.fixture(scope="module")
def eng():
    from rust_local_rag_b200 import engine
    return engine


@pytest.mark.parametrize("n,dim", [(300000, 768), (400001, 384), (270000, 1024), (1000, 96), (129, 768), (60000, 2000)])
@pytest.mark.parametrize("nq", [2, 3])
def test_multi_equals_single_and_oracle(eng, rlr, orc, n, dim, nq):
    rows = orc.synth_rows(n, dim, kind=1, n_clusters=128, seed=n + dim)
    qs = orc.synth_rows(6, dim, kind=1, seed=0x5EED0002, n_clusters=128)
    # RLR_STORE_NO_LATENCY_PATH: the small shapes must go through the same (regular) scan as the large ones
    s = eng.DeviceStore.from_rows(rows, flags=rlr.RLR_STORE_NO_LATENCY_PATH)
    rng = np.random.default_rng(n)
    lex = []
    for i in range(nq):
        if i == 1:
            lex.append(None)                               # an embedding-only query among text queries
        else:
            nl = min(n, 200)
            lr = rng.choice(n, nl, replace=False).astype(np.uint32)
            lex.append((lr, (rng.random(nl) * 5 + 0.1).astype(F32)))
    for k, lam in ((100, 0.7), (5, 0.3), (7, 0.0), (0, 0.5)):
        for lx in (None, lex):
            got = s.search_mmr_multi(qs[:nq], k, lam, W(), lex=lx)
            assert len(got) == nq
            for i in range(nq):
                pair = (None, None) if lx is None or lx[i] is None else lx[i]
                one = s.search_mmr(qs[i], k, lam, W(), pair[0], pair[1])
                ref = orc.search_with_diversity(rows, qs[i], k, lam, lex_rows=pair[0], lex_scores=pair[1], threads=8)
                for a, b, c in zip(got[i], one, ref):
                    assert same(a, b) and same(a, c), (n, dim, nq, k, lam, i, lx is not None)
    # the same query three times: three identical answers; different queries in one pass do not leak into each other
    got = s.search_mmr_multi(np.stack([qs[3]] * nq), 100, 0.7, W())
    for i in range(1, nq):
        assert same(got[i][0], got[0][0]) and same(got[i][1], got[0][1])
    s.search_mmr_multi(qs[:nq], 100, 0.7, W(), flags=rlr.RLR_WANT_TIMINGS)
    assert s.last_timings().launches == 1 + 2 * nq         # ONE scan for all queries, then a pairwise + greedy pair each
    s.close()


def test_multi_f16_store_ties_and_errors(eng, rlr, orc):
    rng = np.random.default_rng(9)
    base = orc.normalize_rows(rng.standard_normal((40, 256)).astype(F32))
    rows = np.tile(base, (3000, 1))                            # massive exact ties in every query's ranking
    qs = rng.standard_normal((3, 256)).astype(F32)
    s = eng.DeviceStore.from_rows(rows, flags=rlr.RLR_STORE_KEEP_F16)
    for flags, data in ((0, rows), (rlr.RLR_SEARCH_F16, rows.astype(np.float16).astype(F32))):
        got = s.search_mmr_multi(qs, 20, 0.5, W(), flags=flags)
        for i in range(3):
            ref = orc.search_with_diversity(data, qs[i], 20, 0.5, threads=8)
            for a, c in zip(got[i], ref):
                assert same(a, c), (flags, i)
    with pytest.raises(rlr.RlrError) as ei:
        s.search_mmr_multi(np.zeros((4, 256), F32), 5, 0.3, W())
    assert ei.value.code == rlr.RLR_ERR_UNSUPPORTED                 # nq > RLR_MAX_MULTI
    bad = qs.copy(); bad[2, 7] = np.inf
    with pytest.raises(rlr.RlrError) as ei:
        s.search_mmr_multi(bad, 5, 0.3, W())
    assert ei.value.code == rlr.RLR_ERR_NONFINITE
    one = s.search_mmr_multi(qs[:1], 5, 0.3, W())                   # nq = 1 forwards to rlr_search_mmr
    ref = orc.search_with_diversity(rows, qs[0], 5, 0.3, threads=8)
    assert same(one[0][0], ref[0])
    s.close()


def test_multi_repeated_calls_with_single_calls_in_between(eng, orc):
    """The per-launch workspace (tickets, tile counter, published bounds) is shared by launches with 1, 2 and 3 groups."""
    rows = orc.synth_rows(500000, 768, kind=1, n_clusters=256)
    qs = orc.synth_rows(12, 768, kind=1, seed=0x5EED0002, n_clusters=256)
    s = eng.DeviceStore.from_rows(rows)
    refs = [orc.search_with_diversity(rows, q, 100, 0.7, threads=8) for q in qs]
    for it in range(8):
        nq = 1 + it % 3
        idx = [(it * 3 + j) % 12 for j in range(nq)]
        got = s.search_mmr_multi(qs[idx], 100, 0.7, W())
        for j, i in enumerate(idx):
            assert same(got[j][0], refs[i][0]) and same(got[j][1], refs[i][1]), (it, j)
        one = s.search_mmr(qs[it], 100, 0.7, W())
        assert same(one[0], refs[it][0])
    s.close()
