"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle
on the same seeded inputs.  Bar: bit-exact scores (f32 arithmetic identical to the
reference), identical row lists (ties: lower row first), identical MMR selections."""
import ctypes as C
import math
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F32 = np.float32


@pytest.fixture(scope="module", params=["latency-path", "regular-path"])
def eng(request):
    """Every test of this module that builds its store with DeviceStore.from_rows runs TWICE: small stores take the
    one-launch latency path by default (rlr_b200.h), and with RLR_STORE_NO_LATENCY_PATH the same store goes through the
    copy + launch sequence that large stores use.  Both must give the oracle's bits."""
    from rust_local_rag_b200 import engine, binding as B
    engine.path_mode = request.param
    if request.param == "latency-path":
        yield engine
        return
    orig = engine.DeviceStore.from_rows.__func__

    def from_rows(cls, rows, device=0, row_base=0, flags=0):
        return orig(cls, rows, device, row_base, flags | B.RLR_STORE_NO_LATENCY_PATH)

    engine.DeviceStore.from_rows = classmethod(from_rows)
    yield engine
    engine.DeviceStore.from_rows = classmethod(orig)
    engine.path_mode = "latency-path"


def once(eng):
    """Large stores never take the latency path: run their tests once."""
    if eng.path_mode != "latency-path":
        pytest.skip("store too large for the latency path: identical to the other parametrisation")


def W(e=0.7, l=0.3):
    from rust_local_rag_b200.engine import ResolvedWeights
    return ResolvedWeights(F32(e), F32(l), F32(0.7), F32(0.3))


def same(a, b):
    return np.asarray(a).tobytes() == np.asarray(b).tobytes()


# ------------------------------------------------------------------ synthetic generator
@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("dim", [3, 100, 768])
def test_device_synth_bit_identical_to_cpu(eng, orc, kind, dim):
    n, base = 1000, 12345
    s = eng.DeviceStore.synthetic(n, dim, kind=kind, seed=0xABCDEF, centroid_seed=77, n_clusters=17, sigma=0.65,
                                  row_base=base)
    got = s.read_rows(np.arange(base, base + n))
    ref = orc.synth_rows(n, dim, kind=kind, seed=0xABCDEF, centroid_seed=77, n_clusters=17, sigma=0.65, row0=base)
    assert same(got, ref)
    s.close()


# ------------------------------------------------------------------ scan + top-m
@pytest.mark.parametrize("n,dim,m", [
    (1, 3, 1), (7, 3, 5), (127, 32, 127), (128, 33, 10), (129, 100, 129), (1000, 384, 45),
    (4097, 768, 300), (20000, 768, 900), (5000, 1024, 1024), (300, 2000, 64), (60000, 64, 1000),
])
def test_search_topm_parity(eng, orc, n, dim, m):
    rng = np.random.default_rng(n * 31 + dim)
    rows = orc.normalize_rows(rng.standard_normal((n, dim)).astype(F32))
    q = rng.standard_normal(dim).astype(F32)
    s = eng.DeviceStore.from_rows(rows)
    for (we, wl) in ((0.7, 0.3), (1.0, 0.0)):
        r, c, e, l = s.search_topm(q, m, W(we, wl))
        R, Cc, E, L = orc.search(rows, q, m, w_embed=we, w_lex=wl, full_sort=n <= 5000, threads=4)
        assert same(r, R), (n, dim, m)
        assert same(c, Cc) and same(e, E) and same(l, L)
    s.close()


def test_ties_lower_row_first_and_duplicates(eng, orc):
    rng = np.random.default_rng(5)
    base = orc.normalize_rows(rng.standard_normal((50, 768)).astype(F32))
    rows = np.tile(base, (200, 1))                      # every row appears 200 times: massive exact ties
    q = rng.standard_normal(768).astype(F32)
    s = eng.DeviceStore.from_rows(rows)
    r, c, e, l = s.search_topm(q, 900, W())
    R, Cc, E, L = orc.search(rows, q, 900, full_sort=True)
    assert same(r, R) and same(c, Cc) and same(e, E)
    # all rows identical: every score ties, result must be rows 0..m-1
    rows2 = np.tile(base[:1], (70000, 1))
    s2 = eng.DeviceStore.from_rows(rows2)
    r2, c2, _, _ = s2.search_topm(q, 300, W())
    assert (r2 == np.arange(300)).all() and len(set(c2.tolist())) == 1
    s.close(); s2.close()


def test_lexical_blend_parity(eng, orc):
    rng = np.random.default_rng(8)
    n, dim = 30000, 768
    rows = orc.normalize_rows(rng.standard_normal((n, dim)).astype(F32))
    q = rng.standard_normal(dim).astype(F32)
    s = eng.DeviceStore.from_rows(rows)
    lex_rows = rng.choice(n, 1500, replace=False).astype(np.uint32)
    lex_scores = (rng.random(1500) * 9).astype(F32)
    for m in (15, 300, 900):
        got = s.search_topm(q, m, W(), lex_rows, lex_scores)
        ref = orc.search(rows, q, m, lex_rows=lex_rows, lex_scores=lex_scores, full_sort=False, threads=4)
        for a, b in zip(got, ref):
            assert same(a, b), m
    assert (got[3] > 0).any()                           # lexical rows do make it into the list
    # out-of-range lexical rows are ignored like a missed HashMap lookup (:525)
    got = s.search_topm(q, 50, W(), np.array([5, n + 7], np.uint32), np.array([1.0, 50.0], F32))
    ref = orc.search(rows, q, 50, lex_rows=np.array([5, n + 7], np.uint32), lex_scores=np.array([1.0, 50.0], F32))
    for a, b in zip(got, ref):
        assert same(a, b)
    s.close()


def test_unnormalised_and_zero_rows(eng, orc):
    rng = np.random.default_rng(9)
    rows = (rng.standard_normal((2000, 96)) * 3).astype(F32)    # store is used as given
    rows[17] = 0                                                 # zero row scores exactly 0.0 (:1765)
    q = rng.standard_normal(96).astype(F32)
    s = eng.DeviceStore.from_rows(rows)
    got = s.search_topm(q, 1024, W())
    ref = orc.search(rows, q, 1024, full_sort=True)
    for a, b in zip(got, ref):
        assert same(a, b)
    assert 17 in got[0].tolist() and got[2][got[0].tolist().index(17)] == 0.0
    s.close()


def test_embedding_candidates_parity(eng, orc):
    rng = np.random.default_rng(10)
    rows = orc.normalize_rows(rng.standard_normal((9000, 384)).astype(F32))
    q = rng.standard_normal(384).astype(F32)
    s = eng.DeviceStore.from_rows(rows)
    r, sc = s.embedding_candidates(q, 30)
    R, S = orc.embedding_candidates(rows, q, 30)
    assert same(r, R) and same(sc, S)
    s.close()


# ------------------------------------------------------------------ MMR
def _mmr_gpu(eng, cands, top_k, lam):
    if not cands:
        return []
    dim = max(len(c[2]) for c in cands)
    emb = np.zeros((len(cands), dim), F32)
    for i, c in enumerate(cands):
        emb[i, :len(c[2])] = c[2]
    s = eng.DeviceStore.from_rows(emb)                  # rows are NOT normalised, like the reference's tests
    pos = s.mmr(np.arange(len(cands)), np.array([c[1] for c in cands], F32), top_k, lam)
    s.close()
    return [cands[p][0] for p in pos]


def test_reference_mmr_kats_on_gpu(eng):
    """/root/reference/src/rag_engine.rs:2877-3038, through rlr_mmr."""
    g = lambda c, k, lam: _mmr_gpu(eng, c, k, lam)
    assert g([("chunk1", 0.9, [1, 0, 0])], 5, 0.3) == ["chunk1"]
    assert len(g([("chunk1", 0.9, [1, 0, 0]), ("chunk2", 0.8, [0, 1, 0])], 10, 0.3)) == 2
    assert g([("chunk1", 0.9, [1, 0.1, 0]), ("chunk2", 0.8, [1, 0.2, 0]), ("chunk3", 0.7, [1, 0.3, 0])], 3, 0.0) == \
        ["chunk1", "chunk2", "chunk3"]
    assert g([("chunk1", 0.9, [1, 0, 0]), ("chunk2", 0.85, [0.99, 0.1, 0]), ("chunk3", 0.7, [0, 1, 0])], 2, 0.9) == \
        ["chunk1", "chunk3"]
    r = g([("chunk1", 0.9, [1, 0, 0]), ("chunk_nan", math.nan, [0, 1, 0]), ("chunk3", 0.7, [0, 0, 1])], 3, 0.3)
    assert len(r) == 2 and "chunk_nan" not in r
    r = g([("chunk1", 0.9, [1, 0, 0]), ("chunk_inf", math.inf, [0, 1, 0]), ("chunk3", 0.7, [0, 0, 1])], 3, 0.3)
    assert len(r) == 2 and "chunk_inf" not in r
    assert g([("a", 0.9, [1, 0, 0, 0]), ("b", 0.8, [0, 1, 0, 0]), ("c", 0.7, [0, 0, 1, 0]), ("d", 0.6, [0, 0, 0, 1])],
             4, 0.3) == ["a", "b", "c", "d"]
    c = [("selected", 0.9, [1, 0, 0]), ("similar", 0.8, [1, 0, 0]), ("diverse", 0.6, [0, 1, 0])]
    assert g(c, 2, 0.5) == ["selected", "diverse"]
    assert g(c, 3, 0.5) == ["selected", "diverse", "similar"]
    assert g([("a", 0.9, [1, 0]), ("b", 0.8, [0, 1])], 0, 0.3) == ["a"]          # top_k = 0 -> first only
    t = [(f"t{i}", 0.5, [1.0 if j == i else 0.0 for j in range(5)]) for i in range(5)]
    assert g(t, 5, 0.0) == ["t0", "t4", "t3", "t2", "t1"]                         # swap_remove tie order


@pytest.mark.parametrize("p,dim,k,lam", [(15, 768, 5, 0.3), (300, 768, 100, 0.7), (300, 1024, 100, 0.3),
                                         (45, 384, 45, 1.0), (333, 96, 50, 0.5), (1024, 64, 100, 0.7),
                                         (2, 768, 5, 0.9), (300, 768, 100, float("nan"))])
def test_mmr_parity_clustered(eng, orc, p, dim, k, lam):
    rng = np.random.default_rng(p + dim)
    cent = rng.standard_normal((8, dim)).astype(F32)
    rows = orc.normalize_rows(cent[rng.integers(0, 8, p)] + 0.5 * rng.standard_normal((p, dim)).astype(F32))
    rel = np.sort(rng.random(p).astype(F32))[::-1].copy()
    rel[p // 2] = rel[p // 2 - 1]                         # an exact relevance tie
    s = eng.DeviceStore.from_rows(rows)
    got = s.mmr(np.arange(p), rel, k, lam)
    ref = orc.mmr(rows, rel, k, lam, threads=4)
    assert same(got, ref), (got, ref)
    s.close()


def test_mmr_nonfinite_inputs(eng, orc):
    rng = np.random.default_rng(3)
    rows = orc.normalize_rows(rng.standard_normal((40, 32)).astype(F32))
    rel = np.sort(rng.random(40).astype(F32))[::-1].copy()
    rel[3] = np.nan; rel[7] = np.inf; rel[9] = -np.inf
    s = eng.DeviceStore.from_rows(rows)
    got = s.mmr(np.arange(40), rel, 40, 0.4)
    ref = orc.mmr(rows, rel, 40, 0.4)
    assert same(got, ref) and len(got) == 37
    rel[:] = np.nan                                       # nothing valid after the first pick
    assert same(s.mmr(np.arange(40), rel, 10, 0.4), orc.mmr(rows, rel, 10, 0.4))
    s.close()


# ------------------------------------------------------------------ fused search_with_diversity
@pytest.mark.parametrize("kind", [0, 1])
def test_config1_search_documents_10k(eng, orc, kind):
    """BASELINE config 1: top_k=5 diversity=0.3 over 10k synthetic 768-d chunks."""
    n, dim = 10000, 768
    rows = orc.synth_rows(n, dim, kind=kind, n_clusters=64)
    s = eng.DeviceStore.from_rows(rows)
    for qi in range(6):
        q = orc.synth_rows(1, dim, kind=kind, seed=0x5EED0002, n_clusters=64, row0=qi)[0] * F32(3.0)
        for (k, lam) in ((5, 0.3), (5, 0.0), (100, 0.7), (0, 0.3), (1, 1.0)):
            got = s.search_mmr(q, k, lam, W())
            ref = orc.search_with_diversity(rows, q, k, lam, full_sort=True)
            for a, b in zip(got, ref):
                assert same(a, b), (kind, qi, k, lam)
    s.close()


def test_config2_1m_x_768_top100_mmr(eng, orc):
    """BASELINE config 2: single-query top_k=100 diversity=0.7 MMR over 1M x 768 f32 chunks."""
    once(eng)
    n, dim = 1_000_000, 768
    s = eng.DeviceStore.synthetic(n, dim, kind=1, n_clusters=4096)
    rows = orc.synth_rows(n, dim, kind=1, n_clusters=4096)         # bit-identical twin on the host
    assert same(s.read_rows([0, 1, n // 2, n - 1]), rows[[0, 1, n // 2, n - 1]])
    for qi in range(3):
        q = orc.synth_rows(1, dim, kind=1, seed=0x5EED0002, n_clusters=4096, row0=qi * 17)[0]
        got = s.search_mmr(q, 100, 0.7, W())
        ref = orc.search_with_diversity(rows, q, 100, 0.7, full_sort=False, threads=orc.max_threads())
        for a, b in zip(got, ref):
            assert same(a, b), qi
        assert len(got[0]) == 100 and len(set(got[0].tolist())) == 100
        top = s.search_topm(q, 900, W())
        topr = orc.search(rows, q, 900, full_sort=False, threads=orc.max_threads())
        assert same(top[0], topr[0]) and same(top[1], topr[1])
    s.close()


# ------------------------------------------------------------------ size-independent properties
def test_properties_full_size(eng):
    once(eng)
    n, dim = 1_000_000, 768
    s = eng.DeviceStore.synthetic(n, dim, kind=0)
    q = np.random.default_rng(1).standard_normal(dim).astype(F32)
    r900, c900, e900, _ = s.search_topm(q, 900, W())
    r300, c300, _, _ = s.search_topm(q, 300, W())
    assert same(r900[:300], r300) and same(c900[:300], c300)            # prefix property
    assert (np.diff(c900.astype(np.float64)) <= 0).all()                # sorted descending
    assert len(set(r900.tolist())) == 900
    again = s.search_topm(q, 900, W())
    assert same(again[0], r900) and same(again[1], c900)                # idempotent
    # the scores are the exact sequential dot of the stored rows
    from oracle import orc as O
    back = s.read_rows(r900[:20])
    qn = O.normalize(q)
    assert same(e900[:20], np.array([O.dot(qn, b) for b in back], F32))
    # querying with a stored row returns that row first with score w_e * ~1
    probe = s.read_rows([123456])[0]
    r, c, e, _ = s.search_topm(probe, 5, W(1.0, 0.0))
    assert r[0] == 123456 and abs(e[0] - 1.0) < 1e-5
    # MMR with lambda -> 0+ keeps relevance order; selection is a subset of the pool
    sel = s.search_mmr(q, 100, 1e-9, W())
    assert same(sel[0], r300[:100])
    sel7 = s.search_mmr(q, 100, 0.7, W())
    assert set(sel7[0].tolist()) <= set(r300.tolist()) and sel7[0][0] == r300[0]
    s.close()


# ------------------------------------------------------------------ errors, empties, re-entrancy
def test_errors_and_empty(eng, rlr):
    s = eng.DeviceStore.from_rows(np.eye(8, dtype=F32))
    with pytest.raises(rlr.RlrError) as ei:
        s.search_topm(np.ones(7, F32), 3, W())
    assert ei.value.code == rlr.RLR_ERR_DIM_MISMATCH
    bad = np.ones(8, F32); bad[2] = np.nan
    with pytest.raises(rlr.RlrError) as ei:
        s.search_topm(bad, 3, W())
    assert ei.value.code == rlr.RLR_ERR_NONFINITE
    with pytest.raises(rlr.RlrError) as ei:
        s.search_topm(np.ones(8, F32), 5000, W())
    assert ei.value.code == rlr.RLR_ERR_UNSUPPORTED
    r, c, e, l = s.search_topm(np.ones(8, F32), 100, W())               # m > n_rows
    assert len(r) == 8
    s.close()
    empty = eng.DeviceStore.from_rows(np.zeros((0, 16), F32))
    assert len(empty.search_topm(np.ones(16, F32), 5, W())[0]) == 0      # :476-478 Ok(vec![])
    assert len(empty.search_mmr(np.ones(16, F32), 5, 0.3, W())[0]) == 0
    empty.close()


def test_concurrent_searches_one_store(eng, orc):
    """Searches hold only the read lock in the reference (src/mcp_server.rs:89): re-entrant."""
    rng = np.random.default_rng(4)
    rows = orc.normalize_rows(rng.standard_normal((50000, 256)).astype(F32))
    qs = rng.standard_normal((8, 256)).astype(F32)
    s = eng.DeviceStore.from_rows(rows)
    refs = [orc.search_with_diversity(rows, q, 20, 0.5, threads=2) for q in qs]
    errs = []

    def work(i):
        try:
            for _ in range(5):
                got = s.search_mmr(qs[i], 20, 0.5, W())
                if not (same(got[0], refs[i][0]) and same(got[1], refs[i][1])):
                    errs.append(i)
        except Exception as ex:  # pragma: no cover
            errs.append(repr(ex))

    th = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in th]; [t.join() for t in th]
    assert not errs
    s.close()


# ------------------------------------------------------------------ in-kernel cross-CTA merge, adversarial layouts
@pytest.mark.parametrize("n,m", [(2000, 300), (70000, 300), (70000, 900), (70000, 1024), (200000, 15), (19000, 1000),
                                 (148 * 128 * 3, 700)])
def test_merge_paths_on_sorted_and_skewed_data(eng, orc, n, m):
    """Scores that decrease with the row index put the whole top-m into a few CTAs' lists, which
    defeats the sampled merge and forces the 'extras' and the bisection paths of final_merge."""
    dim = 64
    rng = np.random.default_rng(n + m)
    q = orc.normalize(rng.standard_normal(dim).astype(F32))
    noise = rng.standard_normal((n, dim)).astype(F32)
    scale = (np.arange(n, dtype=F32) / F32(n))[:, None]
    rows = orc.normalize_rows(q[None, :] + scale * noise)            # row 0 == q, similarity decays with the row index
    s = eng.DeviceStore.from_rows(rows)
    ref = orc.search(rows, q, m, w_embed=1.0, w_lex=0.0, full_sort=False, threads=4)
    for _ in range(4):                                   # tile hand-out is dynamic: repeat to shake out races
        got = s.search_topm(q, m, W(1.0, 0.0))
        for a, b in zip(got, ref):
            assert same(a, b)
    # reversed: the best rows are the LAST ones (last tiles, whichever CTAs grab them)
    rows_r = rows[::-1].copy()
    s2 = eng.DeviceStore.from_rows(rows_r)
    ref = orc.search(rows_r, q, m, full_sort=False, threads=4)
    for _ in range(4):
        got = s2.search_topm(q, m, W())
        for a, b in zip(got, ref):
            assert same(a, b)
    s.close(); s2.close()


@pytest.mark.parametrize("dim,n,m", [(32, 300000, 300), (32, 300000, 1024), (64, 150000, 15), (100, 90000, 900)])
def test_low_dim_many_tiles_per_stage_ring(eng, orc, dim, n, m):
    """dim <= 64 is one pipeline stage per tile, so the TMA producer runs several TILES ahead of
    the consumers: exercises the per-stage tile-id mailbox of the dynamic tile scheduler."""
    rng = np.random.default_rng(dim * n + m)
    rows = orc.normalize_rows(rng.standard_normal((n, dim)).astype(F32))
    q = rng.standard_normal(dim).astype(F32)
    s = eng.DeviceStore.from_rows(rows)
    ref = orc.search(rows, q, m, full_sort=False, threads=4)
    for _ in range(5):
        got = s.search_topm(q, m, W())
        for a, b in zip(got, ref):
            assert same(a, b)
    s.close()


# ------------------------------------------------------------------ rlr_merge_async (the multi-GPU exchange merge)
@pytest.mark.parametrize("n_lists,m,fill", [(2, 300, 1.0), (8, 300, 1.0), (8, 1024, 1.0), (3, 15, 0.5), (8, 300, 0.1),
                                            (1, 100, 1.0), (64, 1024, 0.7), (40, 900, 1.0)])
def test_merge_async_matches_numpy(eng, rlr, n_lists, m, fill):
    import torch
    rng = np.random.default_rng(n_lists * 1000 + m)
    s = eng.DeviceStore.from_rows(np.eye(4, dtype=F32))
    lib = rlr.load()
    ctx = C.c_void_p()
    rlr.check(lib.rlr_ctx_create(s.handle, C.byref(ctx)))
    rec = np.zeros((n_lists, m), rlr.CAND_DTYPE)
    all_keys = np.unique(rng.integers(1, 1 << 40, size=4 * n_lists * m, dtype=np.uint64))
    all_keys = (rng.permutation(all_keys)[:n_lists * m] << np.uint64(20)).reshape(n_lists, m)
    for j in range(n_lists):
        cnt = int(round(m * fill)) if j % 2 == 0 else m
        k = np.sort(all_keys[j, :cnt])[::-1]
        rec["key"][j, :cnt] = k
        rec["emb"][j, :cnt] = (k % np.uint64(1000)).astype(F32)
        rec["lex"][j, :cnt] = (k % np.uint64(7)).astype(F32)
    d_in = torch.from_numpy(rec.view(np.int64).reshape(n_lists, m, 2)).cuda()
    d_out = torch.zeros((m, 2), dtype=torch.int64, device="cuda")
    d_n = torch.zeros(1, dtype=torch.int32, device="cuda")
    rlr.check(lib.rlr_merge_async(ctx, C.c_void_p(d_in.data_ptr()), n_lists, m, C.c_void_p(d_out.data_ptr()),
                                  C.c_void_p(d_n.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().reshape(-1).view(rlr.CAND_DTYPE)
    flat = rec.reshape(-1)
    flat = flat[flat["key"] != 0]
    ref = flat[np.argsort(flat["key"])[::-1]][:m]
    assert int(d_n.item()) == len(ref)
    assert got[:len(ref)].tobytes() == ref.tobytes()
    assert (got["key"][len(ref):] == 0).all()
    lib.rlr_ctx_destroy(ctx)
    s.close()


# ------------------------------------------------------------------ binary16 store copy (BASELINE config 5)
def _round16(rows):
    return rows.astype(np.float16).astype(F32)


@pytest.mark.parametrize("n,dim,m", [(5000, 768, 300), (20000, 100, 45), (3000, 64, 900), (129, 3, 10), (40000, 1024, 300)])
def test_f16_only_store_is_exact_on_rounded_rows(eng, rlr, orc, n, dim, m):
    """An f16 store is 'the reference run on the binary16-rounded embeddings': rows are rounded once
    (RN-even), widened exactly, and every sum is still the sequential f32 one -> bit-identical."""
    rng = np.random.default_rng(n + dim)
    rows = orc.normalize_rows(rng.standard_normal((n, dim)).astype(F32))
    r16 = _round16(rows)
    q = rng.standard_normal(dim).astype(F32)
    s = eng.DeviceStore.from_rows(rows, flags=rlr.RLR_STORE_F16_ONLY)
    assert same(s.read_rows(np.arange(min(n, 50))), r16[:50])
    for a, b in zip(s.search_topm(q, m, W()), orc.search(r16, q, m, full_sort=n <= 5000, threads=4)):
        assert same(a, b)
    for (k, lam) in ((5, 0.3), (100, 0.7)):
        for a, b in zip(s.search_mmr(q, k, lam, W()), orc.search_with_diversity(r16, q, k, lam, threads=4)):
            assert same(a, b)
    s.close()


def test_keep_f16_store_serves_both_precisions_and_states_tolerance(eng, rlr, orc):
    n, dim = 200_000, 768
    rows = orc.synth_rows(n, dim, kind=1, n_clusters=256)
    r16 = _round16(rows)
    s = eng.DeviceStore.from_rows(rows, flags=rlr.RLR_STORE_KEEP_F16)
    q = orc.synth_rows(1, dim, kind=1, seed=0x5EED0002, n_clusters=256)[0]
    f32 = s.search_topm(q, 300, W())
    f16 = s.search_topm(q, 300, W(), flags=rlr.RLR_SEARCH_F16)
    for a, b in zip(f32, orc.search(rows, q, 300, threads=4)):
        assert same(a, b)
    for a, b in zip(f16, orc.search(r16, q, 300, threads=4)):
        assert same(a, b)
    mm16 = s.search_mmr(q, 100, 0.7, W(), flags=rlr.RLR_SEARCH_F16)
    for a, b in zip(mm16, orc.search_with_diversity(r16, q, 100, 0.7, threads=4)):
        assert same(a, b)
    # STATED f16 tolerance vs the f32 store (unit vectors, dim 768): every score of a common row within
    # 2e-4 absolute (binary16 has 11 significant bits: |dx| <= 2^-11 |x| per element, errors average out
    # over 768 terms), and the top-300 sets overlap >= 95 %.
    common = set(f32[0].tolist()) & set(f16[0].tolist())
    assert len(common) >= 285
    e32 = dict(zip(f32[0].tolist(), f32[2].tolist())); e16 = dict(zip(f16[0].tolist(), f16[2].tolist()))
    dev = max(abs(e32[r] - e16[r]) for r in common)
    assert dev < 2e-4, dev
    s.close()
    with pytest.raises(rlr.RlrError):
        s2 = eng.DeviceStore.from_rows(rows[:100])
        try:
            s2.search_topm(q, 5, W(), flags=rlr.RLR_SEARCH_F16)      # no f16 copy in this store
        finally:
            s2.close()


def test_f16_synthetic_store_matches_rounded_cpu_rows(eng, rlr, orc):
    n, dim = 3000, 768
    s = eng.DeviceStore.synthetic(n, dim, kind=1, n_clusters=64, flags=rlr.RLR_STORE_F16_ONLY, row_base=500)
    ref = _round16(orc.synth_rows(n, dim, kind=1, n_clusters=64, row0=500))
    assert same(s.read_rows(np.arange(500, 500 + n)), ref)
    s.close()


def _host_ram_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 1e9
    except Exception:
        return 0.0


def test_config3_10m_x_768_full_size(eng, orc):
    """BASELINE configs[2] / the bench workload at FULL size: 10M x 768 f32 (30.7 GB), top_k=100, diversity 0.7:
    two row shards posting through a mailbox (the fused multi-GPU exchange, here on one GPU) must reproduce the
    single-store result exactly, every returned score must be the oracle's sequential dot of the regenerated row,
    and the size-independent properties hold.  The comparison with the oracle's scan of all 10M rows is the next
    test (it needs 31 GB of host RAM and says so when it cannot run)."""
    once(eng)
    import ctypes as C
    import torch
    from rust_local_rag_b200 import binding as B, dist as rdist
    if torch.cuda.mem_get_info()[0] < 70e9:
        pytest.skip("needs ~62 GB of free HBM")
    n, dim, k, lam = 10_000_000, 768, 100, 0.7
    kw = dict(kind=1, seed=0x5EED0001, centroid_seed=0x5EED00C0, n_clusters=4096, sigma=0.65)
    s = eng.DeviceStore.synthetic(n, dim, **kw)
    qs = orc.synth_rows(3, dim, **{**kw, "seed": 0x5EED0002})
    got = [s.search_mmr(q, k, lam, W(), flags=B.RLR_QUERY_PRENORMALIZED) for q in qs]
    pools = [s.search_topm(q, 300, W(), flags=B.RLR_QUERY_PRENORMALIZED) for q in qs]
    for (rows, score, emb, lex), pool, q in zip(got, pools, qs):
        assert len(rows) == k and len(set(rows.tolist())) == k
        assert set(rows.tolist()) <= set(pool[0].tolist()) and rows[0] == pool[0][0]      # MMR picks from the pool, best first
        assert (np.diff(pool[1].astype(np.float64)) <= 0).all()
        for r, e in list(zip(rows, emb))[::7]:                                           # exact scores of regenerated rows
            assert np.float32(orc.dot(q, orc.synth_rows(1, dim, row0=int(r), **kw)[0])).tobytes() == np.float32(e).tobytes()
    # (b) two shards (uneven) + mailbox on the same GPU
    lib = B.load()
    bounds = [(0, 4_200_000), (4_200_000, n)]
    shards = [eng.DeviceStore.synthetic(hi - lo, dim, row_base=lo, **kw) for lo, hi in bounds]
    ctxs = []
    for sh in shards:
        c = C.c_void_p()
        B.check(lib.rlr_ctx_create(sh.handle, C.byref(c)))
        ctxs.append(c)
    mb = C.c_void_p()
    B.check(lib.rlr_mailbox_create(0, 2, 300, 2, C.byref(mb)))
    qd = torch.zeros(B.RLR_MAX_DIM + 64, device="cuda")
    out = torch.zeros((300, 2), dtype=torch.int64, device="cuda")
    out_n = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for seq, (q, pool) in enumerate(zip(qs, pools), start=1):
        qd.zero_(); qd[:dim] = torch.from_numpy(q).cuda()
        for r in range(2):
            B.check(lib.rlr_topm_post_async(ctxs[r], mb, r, seq, C.c_void_p(qd.data_ptr()), 0.7, 0.3, None, None, 0, 300, st))
        B.check(lib.rlr_mailbox_merge_async(ctxs[0], mb, seq, 300, C.c_void_p(out.data_ptr()), C.c_void_p(out_n.data_ptr()), st))
        torch.cuda.synchronize()
        mr, ms, me, _ = rdist.decode_result(out, int(out_n.item()))
        assert same(mr, pool[0]) and same(ms, pool[1]) and same(me, pool[2])
    lib.rlr_mailbox_close(mb)
    for c in ctxs:
        lib.rlr_ctx_destroy(c)
    for sh in shards:
        sh.close()
    s.close()


def test_config3_10m_x_768_against_the_oracle_scan_of_all_rows(eng, orc):
    """The bench workload at FULL size against the CPU oracle scanning the same 10M rows (bit-identical synthetic
    twin, 30.7 GB of host RAM): every row, score and MMR pick bit for bit.  Skips WITH THE REASON when the box cannot
    hold the host copy; bench.py makes the same comparison in every N=1 run (`parity.oracle_full_results_compared`)."""
    import torch
    from rust_local_rag_b200 import binding as B
    once(eng)
    if torch.cuda.mem_get_info()[0] < 40e9:
        pytest.skip("needs ~31 GB of free HBM")
    ram, thr = _host_ram_gb(), orc.max_threads()
    if ram <= 45 or thr < 8:
        pytest.skip(f"oracle scan of 10M x 768 needs > 45 GB of free host RAM and >= 8 threads (box has {ram:.0f} GB, {thr} threads)")
    n, dim, k, lam = 10_000_000, 768, 100, 0.7
    kw = dict(kind=1, seed=0x5EED0001, centroid_seed=0x5EED00C0, n_clusters=4096, sigma=0.65)
    s = eng.DeviceStore.synthetic(n, dim, **kw)
    qs = orc.synth_rows(2, dim, **{**kw, "seed": 0x5EED0002})
    host = orc.synth_rows(n, dim, **kw)
    for q in qs:
        got = s.search_mmr(q, k, lam, W(), flags=B.RLR_QUERY_PRENORMALIZED)
        ref = orc.search_with_diversity(host, q, k, lam, normalize_query=False, full_sort=False, threads=thr)
        for a, b in zip(got, ref):
            assert same(a, b)
    del host
    s.close()
