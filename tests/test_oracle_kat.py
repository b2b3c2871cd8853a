"""Pins the CPU oracle (oracle/rlr_oracle.c) against every known-answer test the reference
holds for the hot path: /root/reference/src/rag_engine.rs:2674-2799 (cosine/dot/normalize),
:2877-3038 (MMR, run there on a verbatim copy of mmr_diversify) and :3044-3226 (weights),
plus the derived vectors listed in SURVEY.md section 4."""
import math

import numpy as np
import pytest

F32 = np.float32


# ---------------------------------------------------------------- cosine / dot / normalize
def test_cosine_identical(orc):          # :2674-2683
    assert abs(orc.cosine([1, 0, 0], [1, 0, 0]) - 1.0) < 1e-6


def test_cosine_orthogonal(orc):         # :2685-2694
    assert abs(orc.cosine([1, 0, 0], [0, 1, 0])) < 1e-6


def test_cosine_opposite(orc):           # :2696-2705
    assert abs(orc.cosine([1, 0, 0], [-1, 0, 0]) + 1.0) < 1e-6


def test_cosine_zero_vectors(orc):       # :2707-2716
    assert orc.cosine([0, 0, 0], [1, 2, 3]) == 0.0
    assert orc.cosine([0, 0, 0], [0, 0, 0]) == 0.0


def test_cosine_near_zero(orc):          # :2718-2728
    assert orc.cosine([1e-12, 1e-12, 1e-12], [1, 2, 3]) == 0.0


def test_cosine_mismatched_length(orc):  # :2730-2739
    assert orc.cosine([1, 2, 3], [1, 2]) == 0.0


def test_cosine_empty(orc):              # :2741-2747
    assert orc.cosine([], []) == 0.0


def test_cosine_clamping(orc):           # :2749-2759
    s = orc.cosine([1, 1, 1], [1, 1, 1])
    assert -1.0 <= s <= 1.0


def _ramps(dim=384):
    a = np.array([F32(i) / F32(dim) for i in range(dim)], dtype=F32)
    b = np.array([F32(i + 10) / F32(dim) for i in range(dim)], dtype=F32)
    return a, b


def test_cosine_realistic(orc):          # :2761-2774
    a, b = _ramps()
    s = orc.cosine(a, b)
    assert 0.9 < s < 1.0


def test_dot_equals_cosine_when_normalized(orc):  # :2776-2799
    a, b = _ramps()
    cos = orc.cosine(a, b)
    d = orc.dot(orc.normalize(a), orc.normalize(b))
    assert abs(cos - d) < 1e-6


def test_dot_truncates_to_shorter(orc):  # zip semantics, :1778
    assert orc.dot([1, 2, 3], [4, 5]) == 14.0


def test_dot_is_strict_sequential_f32(orc):
    """Independent restatement in numpy scalars: one rounding per mul and per add, in index order."""
    rng = np.random.default_rng(7)
    for dim in (1, 3, 31, 384, 768, 1024):
        a = rng.standard_normal(dim).astype(F32)
        b = rng.standard_normal(dim).astype(F32)
        acc = F32(0.0)
        for x, y in zip(a, b):
            acc = F32(acc + F32(x * y))
        got = orc.dot(a, b)
        assert np.float32(got).tobytes() == acc.tobytes(), dim


def test_normalize_semantics(orc):       # :1763-1771
    v = orc.normalize([3.0, 4.0])
    assert np.allclose(v, [0.6, 0.8], atol=1e-7)
    z = orc.normalize([0.0, 0.0, 0.0])
    assert (z == 0).all()
    tiny = np.array([1e-11, 0, 0], dtype=F32)     # norm_sq = 1e-22 <= 1e-20: untouched
    assert orc.normalize(tiny).tobytes() == tiny.tobytes()
    rng = np.random.default_rng(3)
    x = rng.standard_normal(768).astype(F32)
    s = F32(0)
    for e in x:
        s = F32(s + F32(e * e))
    ref = np.array([F32(e / np.sqrt(s, dtype=F32)) for e in x], dtype=F32)
    assert orc.normalize(x).tobytes() == ref.tobytes()


# ---------------------------------------------------------------- MMR
def _mmr(orc, cands, top_k, lam):
    """cands: list of (id, score, embedding); returns selected ids like test_mmr_diversify (:2824-2875)."""
    if not cands:
        return []
    dim = max(len(c[2]) for c in cands)
    emb = np.zeros((len(cands), dim), F32)
    for i, c in enumerate(cands):
        emb[i, :len(c[2])] = c[2]
    rel = np.array([c[1] for c in cands], F32)
    pos = orc.mmr(emb, rel, top_k, lam)
    return [cands[p][0] for p in pos]


def test_mmr_empty(orc):                 # :2877-2884
    assert _mmr(orc, [], 5, 0.3) == []


def test_mmr_single(orc):                # :2886-2897
    assert _mmr(orc, [("chunk1", 0.9, [1, 0, 0])], 5, 0.3) == ["chunk1"]


def test_mmr_topk_larger(orc):           # :2899-2914
    r = _mmr(orc, [("chunk1", 0.9, [1, 0, 0]), ("chunk2", 0.8, [0, 1, 0])], 10, 0.3)
    assert len(r) == 2


def test_mmr_zero_diversity(orc):        # :2916-2935 (+ derived full list)
    r = _mmr(orc, [("chunk1", 0.9, [1, 0.1, 0]), ("chunk2", 0.8, [1, 0.2, 0]), ("chunk3", 0.7, [1, 0.3, 0])], 3, 0.0)
    assert len(r) == 3 and r[0] == "chunk1"
    assert r == ["chunk1", "chunk2", "chunk3"]


def test_mmr_high_diversity(orc):        # :2937-2961
    r = _mmr(orc, [("chunk1", 0.9, [1, 0, 0]), ("chunk2", 0.85, [0.99, 0.1, 0]), ("chunk3", 0.7, [0, 1, 0])], 2, 0.9)
    assert r == ["chunk1", "chunk3"]


def test_mmr_nan_score(orc):             # :2963-2980
    r = _mmr(orc, [("chunk1", 0.9, [1, 0, 0]), ("chunk_nan", math.nan, [0, 1, 0]), ("chunk3", 0.7, [0, 0, 1])], 3, 0.3)
    assert len(r) == 2 and "chunk_nan" not in r


def test_mmr_inf_score(orc):             # :2982-2999
    r = _mmr(orc, [("chunk1", 0.9, [1, 0, 0]), ("chunk_inf", math.inf, [0, 1, 0]), ("chunk3", 0.7, [0, 0, 1])], 3, 0.3)
    assert len(r) == 2 and "chunk_inf" not in r


def test_mmr_orthogonal_keeps_order(orc):  # :3001-3018 (+ derived)
    r = _mmr(orc, [("a", 0.9, [1, 0, 0, 0]), ("b", 0.8, [0, 1, 0, 0]), ("c", 0.7, [0, 0, 1, 0]),
                   ("d", 0.6, [0, 0, 0, 1])], 4, 0.3)
    assert len(r) == 4 and r[0] == "a"
    assert r == ["a", "b", "c", "d"]


def test_mmr_formula(orc):               # :3020-3044 (+ derived top_k=3)
    c = [("selected", 0.9, [1, 0, 0]), ("similar", 0.8, [1, 0, 0]), ("diverse", 0.6, [0, 1, 0])]
    assert _mmr(orc, c, 2, 0.5) == ["selected", "diverse"]
    assert _mmr(orc, c, 3, 0.5) == ["selected", "diverse", "similar"]


def test_mmr_topk_zero_returns_first(orc):  # first is pushed before the loop check, :782-788
    assert _mmr(orc, [("a", 0.9, [1, 0]), ("b", 0.8, [0, 1])], 0, 0.3) == ["a"]


def test_mmr_tie_order_is_swap_remove_order(orc):  # SURVEY 4 derived vector
    c = [(f"t{i}", 0.5, [1.0 if j == i else 0.0 for j in range(5)]) for i in range(5)]
    assert _mmr(orc, c, 5, 0.0) == ["t0", "t4", "t3", "t2", "t1"]


def test_mmr_threads_do_not_change_result(orc):
    rng = np.random.default_rng(11)
    emb = orc.normalize_rows(rng.standard_normal((60, 48)).astype(F32) + 0.7)
    rel = np.sort(rng.random(60).astype(F32))[::-1].copy()
    a = orc.mmr(emb, rel, 20, 0.7, threads=1)
    b = orc.mmr(emb, rel, 20, 0.7, threads=4)
    assert (a == b).all()


# ---------------------------------------------------------------- weights
def test_resolve_weight_override(orc):   # :3044-3049
    assert orc.resolve_weight(0.5, 0.7) == F32(0.5)
    assert orc.resolve_weight(0.9, 0.3) == F32(0.9)


def test_resolve_weight_none(orc):       # :3051-3056
    assert orc.resolve_weight(None, 0.7) == F32(0.7)
    assert orc.resolve_weight(None, 0.3) == F32(0.3)


def test_resolve_weight_boundaries(orc):  # :3058-3064
    assert orc.resolve_weight(0.0, 0.5) == 0.0
    assert orc.resolve_weight(1.0, 0.5) == 1.0


@pytest.mark.parametrize("bad", [math.nan, math.inf, -math.inf, -0.1, 1.5, 2.0, -1e-6, 1.000001])
def test_resolve_weight_rejects(orc, bad):  # :3066-3104, :3196-3218
    assert orc.resolve_weight(bad, 0.5) == 0.5


def test_resolve_weight_negative_zero(orc):  # :3220-3225
    r = orc.resolve_weight(-0.0, 0.5)
    assert r == 0.0 and math.copysign(1.0, r) == -1.0


# ---------------------------------------------------------------- search (source-text only; no reference test exists)
def _data(orc, n, dim, seed):
    rng = np.random.default_rng(seed)
    return orc.normalize_rows(rng.standard_normal((n, dim)).astype(F32)), rng.standard_normal(dim).astype(F32)


def test_search_heap_path_equals_literal_full_sort(orc):
    rows, q = _data(orc, 3000, 64, 5)
    lex_rows = np.array([5, 17, 2999, 100], np.uint32)
    lex_scores = np.array([2.5, 0.1, 7.0, 3.3], F32)
    for k in (1, 5, 100, 2000, 5000):
        a = orc.search(rows, q, k, lex_rows=lex_rows, lex_scores=lex_scores, full_sort=True)
        b = orc.search(rows, q, k, lex_rows=lex_rows, lex_scores=lex_scores, full_sort=False, threads=3)
        for x, y in zip(a, b):
            assert x.tobytes() == y.tobytes(), k


def test_search_matches_numpy_restatement(orc):
    rows, q = _data(orc, 500, 96, 9)
    qn = orc.normalize(q)
    emb = np.array([orc.dot(qn, r) for r in rows], F32)
    comb = (F32(0.7) * emb).astype(F32) + F32(0.3) * F32(0.0)
    order = np.lexsort((np.arange(len(comb)), -comb.astype(np.float64)))[:15]
    r, s, e, l = orc.search(rows, q, 15, full_sort=True)
    assert (r == order).all()
    assert s.tobytes() == comb[order].astype(F32).tobytes()
    assert e.tobytes() == emb[order].tobytes()
    assert (l == 0).all()


def test_search_ties_lower_row_first(orc):
    rows = np.tile(orc.normalize(np.arange(1, 9, dtype=F32)), (40, 1))
    r, s, _, _ = orc.search(rows, rows[0], 10, full_sort=True)
    assert (r == np.arange(10)).all()
    r2, _, _, _ = orc.search(rows, rows[0], 10, full_sort=False)
    assert (r2 == np.arange(10)).all()


def test_search_edge_cases(orc):
    rows, q = _data(orc, 7, 16, 1)
    assert len(orc.search(rows[:0].reshape(0, 16), q, 5)[0]) == 0          # empty store, :476
    assert len(orc.search(rows, q, 0)[0]) == 1                               # top_k.max(1), :490
    assert len(orc.search(rows, q, 100)[0]) == 7                             # fewer rows than top_k


def test_lexical_blend(orc):             # :511-532
    rows, q = _data(orc, 50, 32, 2)
    lex_rows = np.array([49, 3], np.uint32)
    lex_scores = np.array([4.0, 1.0], F32)
    r, s, e, l = orc.search(rows, q, 50, lex_rows=lex_rows, lex_scores=lex_scores, full_sort=True)
    li = {int(rr): float(ll) for rr, ll in zip(r, l)}
    assert li[49] == 1.0 and li[3] == 0.25 and sum(v != 0 for v in li.values()) == 2
    i = list(r).index(49)
    assert s[i] == F32(F32(0.7) * e[i] + F32(F32(0.3) * F32(1.0)))


def test_search_with_diversity_pool_and_shortcut(orc):  # :725-759
    rows, q = _data(orc, 400, 32, 4)
    a = orc.search_with_diversity(rows, q, 5, 0.0, full_sort=True)
    b = orc.search(rows, q, 5, full_sort=True)
    assert a[0].tobytes() == b[0].tobytes()                                   # lambda == 0 -> search(top_k)
    pool = orc.search(rows, q, 15, full_sort=True)                            # max(3*5, 5+10) = 15
    pos = orc.mmr(rows[pool[0]], pool[1], 5, 0.3)
    c = orc.search_with_diversity(rows, q, 5, 0.3, full_sort=True)
    assert (c[0] == pool[0][pos]).all()
    d = orc.search_with_diversity(rows, q, 2, 0.3, full_sort=True)            # pool = max(6, 12) = 12
    pool12 = orc.search(rows, q, 12, full_sort=True)
    assert (d[0] == pool12[0][orc.mmr(rows[pool12[0]], pool12[1], 2, 0.3)]).all()
    e = orc.search_with_diversity(rows, q, 3, -5.0, full_sort=True)           # clamp to 0
    assert (e[0] == orc.search(rows, q, 3, full_sort=True)[0]).all()


def test_embedding_candidates_is_raw_dot_order(orc):  # :415-461
    rows, q = _data(orc, 300, 48, 6)
    r, s = orc.embedding_candidates(rows, q, 20)
    qn = orc.normalize(q)
    emb = np.array([orc.dot(qn, x) for x in rows], F32)
    order = np.lexsort((np.arange(300), -emb.astype(np.float64)))[:20]
    assert (r == order).all() and s.tobytes() == emb[order].tobytes()


def test_synth_rows_are_unit_and_reproducible(orc):
    a = orc.synth_rows(64, 768, threads=1)
    b = orc.synth_rows(64, 768, threads=4)
    assert a.tobytes() == b.tobytes()
    assert np.allclose(np.linalg.norm(a.astype(np.float64), axis=1), 1.0, atol=1e-5)
    c = orc.synth_rows(32, 768, row0=32, threads=2)
    assert c.tobytes() == a[32:].tobytes()                                    # shard-able by row0
    cl = orc.synth_rows(4096 * 2, 64, kind=1, n_clusters=4096)
    same = float(np.dot(cl[0], cl[4096])); diff = float(np.dot(cl[0], cl[1]))
    assert same > 0.5 and abs(diff) < 0.5
