"""Batched-query path (tcgen05 GEMM + fused per-query threshold filter), BASELINE config 4.

Floating-point kernel: compared against a plain fp32/fp64 reference of the same contraction on
the same binary16-rounded inputs.  Stated tolerance: |score - ref| <= 2e-6 absolute (f16 x f16
products are exact in f32; only the f32 accumulation order differs), top-m sets identical except
for rows whose reference score is within that tolerance of the cut.  With
RLR_BATCH_EXACT_RESCORE the returned scores are bit-identical to the sequential-f32 oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F32 = np.float32
TOL = 2e-6


def _ref_topm(rows16, q16, m):
    s = q16.astype(np.float64) @ rows16.astype(np.float64).T            # [nq, n]
    order = np.argsort(-s, axis=1, kind="stable")[:, :m]
    return s, order


@pytest.mark.parametrize("n,dim,nq,m", [(4096, 768, 256, 100), (20000, 768, 300, 100), (10000, 1024, 1024, 100),
                                         (5000, 64, 17, 10), (70000, 384, 512, 300), (3000, 768, 5, 900)])
def test_batch_matches_fp_reference(n, dim, nq, m):
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import binding as B, engine
    from oracle import orc
    rng = np.random.default_rng(n + nq)
    rows = orc.normalize_rows(rng.standard_normal((n, dim)).astype(F32))
    qs = rng.standard_normal((nq, dim)).astype(F32)
    s = engine.DeviceStore.from_rows(rows, flags=B.RLR_STORE_KEEP_F16)
    got_rows, got_scores, got_n = s.search_batch(qs, m)
    rows16 = rows.astype(np.float16).astype(F32)
    q16 = np.stack([orc.normalize(q) for q in qs]).astype(np.float16).astype(F32)
    ref, order = _ref_topm(rows16, q16, m)
    m_eff = min(m, n)
    assert (got_n == m_eff).all()
    for q in range(nq):
        r = got_rows[q, :m_eff]
        assert len(set(r.tolist())) == m_eff
        # scores are the contraction of the rounded inputs
        assert np.abs(got_scores[q, :m_eff].astype(np.float64) - ref[q, r]).max() <= TOL, q
        assert (np.diff(got_scores[q, :m_eff].astype(np.float64)) <= 0).all()
        # the set is the reference top-m up to ties at the cut
        cut = ref[q, order[q, m_eff - 1]]
        missing = set(order[q].tolist()) - set(r.tolist())
        assert all(ref[q, x] <= cut + 2 * TOL for x in missing), q
        assert all(ref[q, x] >= cut - 2 * TOL for x in r.tolist()), q
    s.close()


def test_batch_exact_rescore_is_bit_identical_to_single_query_path():
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import binding as B, engine
    from oracle import orc
    n, dim, nq, m = 50000, 768, 64, 100
    rows = orc.synth_rows(n, dim, kind=1, n_clusters=128)
    qs = orc.synth_rows(nq, dim, kind=1, seed=0x5EED0002, n_clusters=128)
    s = engine.DeviceStore.from_rows(rows, flags=B.RLR_STORE_KEEP_F16)
    got_rows, got_scores, _ = s.search_batch(qs, m, flags=B.RLR_BATCH_EXACT_RESCORE)
    exact_hits = 0
    for q in range(nq):
        r, sc = s.embedding_candidates(qs[q], m)                       # exact single-query path (f32 store)
        # every returned score is the exact sequential dot of that row
        qn = orc.normalize(qs[q])
        back = rows[got_rows[q]]
        assert got_scores[q].tobytes() == np.array([orc.dot(qn, b) for b in back], F32).tobytes()
        exact_hits += len(set(r.tolist()) & set(got_rows[q].tolist()))
    # shortlist came from f16 scores: recall of the exact top-100 must be essentially complete
    assert exact_hits >= 0.99 * nq * m
    s.close()


def test_batch_operand_copies_are_required_for_the_16_bit_kinds():
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import binding as B, engine
    s = engine.DeviceStore.from_rows(np.eye(64, dtype=F32))
    for flag in (B.RLR_BATCH_F16, B.RLR_BATCH_BF16):
        with pytest.raises(B.RlrError) as ei:
            s.search_batch(np.ones((2, 64), F32), 3, flags=flag)
        assert ei.value.code == B.RLR_ERR_INVALID_ARG
    rows, _, n = s.search_batch(np.eye(2, 64, dtype=F32), 3)          # plain f32 store: tf32 over the rows themselves
    assert (n == 3).all() and rows[0, 0] == 0 and rows[1, 0] == 1
    s.close()
    s = engine.DeviceStore.from_rows(np.eye(64, dtype=F32), flags=B.RLR_STORE_F16_ONLY)
    with pytest.raises(B.RlrError):
        s.search_batch(np.ones((2, 64), F32), 3, flags=B.RLR_BATCH_TF32)
    s.close()


def _tf32_trunc(x):
    return (np.ascontiguousarray(x, F32).view(np.uint32) & np.uint32(0xFFFFE000)).view(F32)


def _bf16_rne(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x, F32)).to(torch.bfloat16).float().numpy()


# stated tolerances of the operand precisions against the EXACT scores on unit vectors (asserted below):
#   binary16 2e-4, tf32 1e-3 (10-bit mantissas, truncated), bfloat16 4e-3 (7-bit mantissas)
TOL_VS_EXACT = {"f16": 2e-4, "tf32": 1e-3, "bf16": 4e-3}


@pytest.mark.parametrize("prec", ["tf32", "bf16", "f16"])
@pytest.mark.parametrize("n,dim,nq,m", [(4096, 768, 256, 100), (20000, 1024, 300, 100), (5000, 96, 17, 10), (70000, 384, 512, 300)])
def test_batch_precisions_match_the_contraction_of_their_rounded_inputs(prec, n, dim, nq, m):
    """north_star kernel (3): "tcgen05/TMEM tiles with tf32/bf16 inputs".  Each precision is the f64 contraction of
    the inputs rounded the way that precision rounds them (tf32: truncation to 10 mantissa bits by the tensor core,
    straight over the f32 store -- no copy; bf16 / binary16: round to nearest even into the store's copy), within
    TOL absolute (the products are exact in f32, only the f32 accumulation order differs), and within the STATED
    tolerance of the exact scores."""
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import binding as B, engine
    from oracle import orc
    rng = np.random.default_rng(n + nq + len(prec))
    rows = orc.normalize_rows(rng.standard_normal((n, dim)).astype(F32))
    qs = rng.standard_normal((nq, dim)).astype(F32)
    qn = np.stack([orc.normalize(q) for q in qs])
    store_flags = {"tf32": 0, "bf16": B.RLR_STORE_KEEP_BF16, "f16": B.RLR_STORE_KEEP_F16}[prec]
    flag = {"tf32": B.RLR_BATCH_TF32, "bf16": B.RLR_BATCH_BF16, "f16": B.RLR_BATCH_F16}[prec]
    rnd = {"tf32": _tf32_trunc, "bf16": _bf16_rne, "f16": lambda x: x.astype(np.float16).astype(F32)}[prec]
    s = engine.DeviceStore.from_rows(rows, flags=store_flags)
    got_rows, got_scores, got_n = s.search_batch(qs, m, flags=flag)
    ref, order = _ref_topm(rnd(rows), rnd(qn), m)
    exact = qn.astype(np.float64) @ rows.astype(np.float64).T
    m_eff = min(m, n)
    assert (got_n == m_eff).all()
    worst_model, worst_exact = 0.0, 0.0
    for q in range(nq):
        r = got_rows[q, :m_eff]
        assert len(set(r.tolist())) == m_eff
        g = got_scores[q, :m_eff].astype(np.float64)
        worst_model = max(worst_model, np.abs(g - ref[q, r]).max())
        worst_exact = max(worst_exact, np.abs(g - exact[q, r]).max())
        assert (np.diff(g) <= 0).all()
        cut = ref[q, order[q, m_eff - 1]]
        missing = set(order[q].tolist()) - set(r.tolist())
        assert all(ref[q, x] <= cut + 2 * TOL for x in missing), (prec, q)
        assert all(ref[q, x] >= cut - 2 * TOL for x in r.tolist()), (prec, q)
    assert worst_model <= TOL, (prec, worst_model)
    assert worst_exact <= TOL_VS_EXACT[prec], (prec, worst_exact)
    # exact re-score restores the single-query path's bits in every precision
    got_rows, got_scores, _ = s.search_batch(qs[:8], m, flags=flag | B.RLR_BATCH_EXACT_RESCORE)
    for q in range(8):
        back = rows[got_rows[q, :m_eff]]
        assert got_scores[q, :m_eff].tobytes() == np.array([orc.dot(qn[q], b) for b in back], F32).tobytes()
    s.close()


def test_bf16_copy_follows_store_mutation():
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import binding as B, engine
    from oracle import orc
    rng = np.random.default_rng(77)
    rows = orc.normalize_rows(rng.standard_normal((3000, 128)).astype(F32))
    extra = orc.normalize_rows(rng.standard_normal((500, 128)).astype(F32))
    s = engine.DeviceStore.from_rows(rows, flags=B.RLR_STORE_KEEP_BF16)
    s.remove_rows(np.arange(100, 400))
    s.append(extra)
    n = s.info().n_rows
    now = s.read_rows(np.arange(n))
    qs = rng.standard_normal((16, 128)).astype(F32)
    qn = np.stack([orc.normalize(q) for q in qs])
    got_rows, got_scores, _ = s.search_batch(qs, 50, flags=B.RLR_BATCH_BF16)
    ref = _bf16_rne(qn).astype(np.float64) @ _bf16_rne(now).astype(np.float64).T
    for q in range(16):
        assert np.abs(got_scores[q].astype(np.float64) - ref[q, got_rows[q]]).max() <= TOL
    s.close()


def test_batch_on_skewed_rows_takes_the_checked_path():
    """Scores that grow with the row index make every later row beat the frozen thresholds: the
    unchecked fast pass overflows its candidate lists and the batch is redone phase by phase."""
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import binding as B, engine
    from oracle import orc
    n, dim, nq, m = 30000, 64, 40, 100
    rng = np.random.default_rng(1)
    q = orc.normalize(rng.standard_normal(dim).astype(F32))
    noise = rng.standard_normal((n, dim)).astype(F32)
    scale = (np.arange(n, dtype=F32)[::-1] / F32(n))[:, None]            # last rows are the most similar to q
    rows = orc.normalize_rows(q[None, :] + scale * noise)
    qs = np.tile(q, (nq, 1)) + 0.01 * rng.standard_normal((nq, dim)).astype(F32)
    s = engine.DeviceStore.from_rows(rows, flags=B.RLR_STORE_KEEP_F16)
    got_rows, got_scores, got_n = s.search_batch(qs, m)
    rows16 = rows.astype(np.float16).astype(F32)
    q16 = np.stack([orc.normalize(x) for x in qs]).astype(np.float16).astype(F32)
    ref, order = _ref_topm(rows16, q16, m)
    assert (got_n == m).all()
    for qi in range(nq):
        r = got_rows[qi]
        assert np.abs(got_scores[qi].astype(np.float64) - ref[qi, r]).max() <= TOL
        cut = ref[qi, order[qi, m - 1]]
        assert all(ref[qi, x] >= cut - 2 * TOL for x in r.tolist())
        assert all(ref[qi, x] <= cut + 2 * TOL for x in set(order[qi].tolist()) - set(r.tolist()))
    s.close()


@pytest.mark.parametrize("flags_name", ["tensor", "exact"])
def test_sharded_batch_merge_equals_single_store(flags_name):
    """BASELINE config 4 composition on one GPU: three row shards (uneven, one shorter than m) answer
    the same query batch (rlr_search_batch_device), their key lists are merged per query on the
    device (rlr_batch_merge_async).  Must equal the unsharded batch search bit for bit: a row's
    tensor-core score does not depend on which tile or shard it is computed in."""
    import torch
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import binding as B, engine
    from oracle import orc
    n, dim, nq, m = 30000, 768, 200, 100
    flags = B.RLR_BATCH_EXACT_RESCORE if flags_name == "exact" else 0
    rows = orc.synth_rows(n, dim, kind=1, n_clusters=64)
    qs = orc.synth_rows(nq, dim, kind=1, seed=0x5EED0002, n_clusters=64)
    full = engine.DeviceStore.from_rows(rows, flags=B.RLR_STORE_KEEP_F16)
    want_rows, want_scores, want_n = full.search_batch(qs, m, flags=flags)
    bounds = [(0, 60), (60, 17000), (17000, n)]
    shards = [engine.DeviceStore.from_rows(rows[lo:hi], row_base=lo, flags=B.RLR_STORE_KEEP_F16) for lo, hi in bounds]
    lists = torch.zeros((len(shards), nq, m), dtype=torch.int64, device="cuda")
    for j, sh in enumerate(shards):
        sh.search_batch_device(qs, m, lists[j], None, None, flags)
    merged = torch.zeros((nq, m), dtype=torch.int64, device="cuda")
    cnt = torch.zeros(nq, dtype=torch.int32, device="cuda")
    full.batch_merge(lists, len(shards), nq, m, merged, cnt)
    torch.cuda.synchronize()
    keys = merged.cpu().numpy().view(np.uint64)
    assert (cnt.cpu().numpy() == want_n).all()
    got_rows, got_scores = B.key_row(keys), B.key_score(keys)
    if flags_name == "tensor":
        assert got_rows.tobytes() == want_rows.tobytes()
        assert got_scores.tobytes() == want_scores.tobytes()
    else:
        # exact re-score: every shard re-scores ITS shortlist, so the merged list is the exact top-m of a
        # superset of the unsharded shortlist -- scores are the oracle's bits, order is exact, and the set
        # can only differ from the unsharded one at the tensor-core cut (2e-4)
        for q in range(0, nq, 13):
            R, S = orc.embedding_candidates(rows, qs[q], m)
            exact = {int(r): s for r, s in zip(R, S)}
            for r, sc in zip(got_rows[q], got_scores[q]):
                want = exact.get(int(r))
                if want is None:
                    want = np.float32(orc.dot(orc.normalize(qs[q]), rows[int(r)]))
                    assert want >= S[-1] - 2e-4
                assert np.float32(sc).tobytes() == np.float32(want).tobytes(), (q, r)
            assert (np.diff(got_scores[q].astype(np.float64)) <= 0).all()
            assert len(set(got_rows[q].tolist()) ^ set(R.tolist())) <= 4
    for sh in shards:
        sh.close()
    full.close()


def test_batch_equal_scores_at_the_cut_are_ordered_by_row():
    """Every row appears 20 times, so equal tensor-core scores straddle the top-m cut of every query: the
    prune's select must fall back to the row half of the keys (lower row first), exactly like the full sort."""
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import binding as B, engine
    from oracle import orc
    rng = np.random.default_rng(77)
    base = orc.normalize_rows(rng.standard_normal((50, 256)).astype(F32))
    rows = np.tile(base, (20, 1))                                   # 1000 rows
    qs = rng.standard_normal((40, 256)).astype(F32)
    s = engine.DeviceStore.from_rows(rows, flags=B.RLR_STORE_KEEP_F16)
    full_rows, full_scores, full_n = s.search_batch(qs, 1000)      # complete ranking (CTA prune: m > 128)
    assert (full_n == 1000).all()
    for q in range(40):
        order = np.lexsort((full_rows[q], -full_scores[q].astype(np.float64)))
        assert (order == np.arange(1000)).all()                     # (score desc, row asc)
    for m in (100, 7, 128):
        r, sc, n = s.search_batch(qs, m)                            # warp prune
        assert (n == m).all()
        assert r.tobytes() == np.ascontiguousarray(full_rows[:, :m]).tobytes()
        assert sc.tobytes() == np.ascontiguousarray(full_scores[:, :m]).tobytes()
    s.close()


def test_batch_queries_are_normalised_and_checked_on_the_device():
    """The batch's normalize (:494) and NaN/Inf check run on the device: un-normalised queries must give the
    single-query path's exact bits after the exact re-score (that path normalises on the host), and a NaN
    anywhere in the batch is RLR_ERR_NONFINITE."""
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import binding as B, engine
    from oracle import orc
    rng = np.random.default_rng(3)
    n, dim, nq, m = 20000, 384, 33, 50
    rows = orc.normalize_rows(rng.standard_normal((n, dim)).astype(F32))
    qs = (rng.standard_normal((nq, dim)) * 7.5).astype(F32)                # far from unit norm
    s = engine.DeviceStore.from_rows(rows, flags=B.RLR_STORE_KEEP_F16)
    got_rows, got_scores, _ = s.search_batch(qs, m, flags=B.RLR_BATCH_EXACT_RESCORE)
    for q in range(nq):
        R, S = orc.embedding_candidates(rows, qs[q], m)                     # normalises the query like :425
        exact = {int(r): sc for r, sc in zip(R, S)}
        for r, sc in zip(got_rows[q], got_scores[q]):
            want = exact.get(int(r))
            if want is not None:
                assert np.float32(sc).tobytes() == np.float32(want).tobytes(), (q, r)
        assert len(set(got_rows[q].tolist()) & set(R.tolist())) >= m - 3
    bad = qs.copy()
    bad[17, 5] = np.nan
    with pytest.raises(B.RlrError) as ei:
        s.search_batch(bad, m)
    assert ei.value.code == B.RLR_ERR_NONFINITE
    bad[17, 5] = np.inf
    with pytest.raises(B.RlrError) as ei:
        s.search_batch(bad, m)
    assert ei.value.code == B.RLR_ERR_NONFINITE
    again = s.search_batch(qs, m, flags=B.RLR_BATCH_EXACT_RESCORE)          # the store is still usable
    assert again[0].tobytes() == got_rows.tobytes()
    s.close()
