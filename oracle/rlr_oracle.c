/*
 * rlr_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
 *
 * A plain-C, CPU restatement of the arithmetic on rust-local-rag's retrieval hot
 * path (CrashCartCapital/rust-local-rag, src/rag_engine.rs).  It exists so that the
 * CUDA path can be checked for parity.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this file's library.
 * The product library (librlr_b200.so) never links, loads or calls anything here.
 *
 * Build flags are part of the contract: -O2 -ffp-contract=off, no -ffast-math.
 * Every float expression below is written one IEEE-754 binary32 operation at a time
 * so that the result is what Rust/LLVM computes for the cited lines (Rust never
 * contracts a*b+c into an FMA and never re-associates an f32 `.sum()`).
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - orc_dot / orc_normalize / orc_cosine : pinned by the reference's own KATs,
 *     src/rag_engine.rs:2674-2799 (tests/test_oracle_kat.py encodes all ten).
 *   - orc_mmr : pinned by the reference's nine MMR tests, :2877-3038.
 *   - orc_resolve_weight : pinned by the reference's weight tests, :3044-3226.
 *   - orc_search / orc_search_with_diversity / orc_embedding_candidates : the
 *     reference holds NO test for search()'s blend/sort/cut/fallback or the pool
 *     sizing; these follow the source text only (":470-701", ":717-759", ":415-461").
 *   The reference itself cannot be compiled here (no cargo/rustc in the image).
 *
 * Determinism rule for exact-score ties: the reference's tie order is a per-process
 * random HashMap/HashSet iteration order fed to a stable sort (:508,:524,:543,:668);
 * this oracle feeds rows in ascending row order, i.e. "lower row first", which is one
 * of the reference's valid outcomes.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* similarity primitives                                                      */
/* ------------------------------------------------------------------------- */

/* src/rag_engine.rs:1776-1779  dot_product: zip truncates to the shorter slice,
 * `.sum()` is a strict left-to-right f32 fold, one rounding per mul and per add. */
ORC_API float orc_dot(const float *a, size_t na, const float *b, size_t nb)
{
    size_t n = na < nb ? na : nb;
    volatile float acc = 0.0f; /* volatile: forbid any vectorised re-association */
    for (size_t i = 0; i < n; ++i) {
        float p = a[i] * b[i];
        acc = acc + p;
    }
    return acc;
}

/* Same arithmetic without the volatile (gcc -O2 without -ffast-math may not
 * re-associate float adds, so this is bit-identical; tests assert that). */
static inline float dot_seq(const float *a, const float *b, size_t n)
{
    float acc = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float p = a[i] * b[i];
        acc = acc + p;
    }
    return acc;
}

/* src/rag_engine.rs:1763-1771  normalize: s = sum x*x (sequential); if s > 1e-20
 * divide every element by sqrt(s) (a true division each, not a reciprocal mul). */
ORC_API void orc_normalize(float *v, size_t n)
{
    float norm_sq = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float p = v[i] * v[i];
        norm_sq = norm_sq + p;
    }
    if (norm_sq > 1e-20f) {
        float norm = sqrtf(norm_sq);
        for (size_t i = 0; i < n; ++i)
            v[i] = v[i] / norm;
    }
}

/* src/rag_engine.rs:1742-1759  cosine_similarity (tests / legacy only). */
ORC_API float orc_cosine(const float *a, size_t na, const float *b, size_t nb)
{
    if (na != nb)
        return 0.0f;
    const float EPSILON = 1e-10f;
    float dot = 0.0f, sa = 0.0f, sb = 0.0f;
    for (size_t i = 0; i < na; ++i) { float p = a[i] * b[i]; dot = dot + p; }
    for (size_t i = 0; i < na; ++i) { float p = a[i] * a[i]; sa = sa + p; }
    for (size_t i = 0; i < na; ++i) { float p = b[i] * b[i]; sb = sb + p; }
    float norm_a = sqrtf(sa), norm_b = sqrtf(sb);
    if (norm_a < EPSILON || norm_b < EPSILON)
        return 0.0f;
    float d = norm_a * norm_b;
    float c = dot / d;
    /* f32::clamp(-1.0, 1.0); NaN stays NaN */
    if (c < -1.0f) c = -1.0f;
    if (c > 1.0f) c = 1.0f;
    return c;
}

/* ------------------------------------------------------------------------- */
/* weights                                                                    */
/* ------------------------------------------------------------------------- */

/* src/rag_engine.rs:1869-1873  resolve_weight: Some(w) kept iff finite and in
 * [0.0, 1.0] (RangeInclusive::contains, so -0.0 is accepted), else the default. */
ORC_API float orc_resolve_weight(int has_override, float override_w, float dflt)
{
    if (has_override && isfinite(override_w) && override_w >= 0.0f && override_w <= 1.0f)
        return override_w;
    return dflt;
}

/* src/rag_engine.rs:1801-1804 defaults. */
ORC_API void orc_default_weights(float out4[4])
{
    out4[0] = 0.7f; out4[1] = 0.3f; out4[2] = 0.7f; out4[3] = 0.3f;
}

/* ------------------------------------------------------------------------- */
/* search (reranker-absent branch)                                            */
/* ------------------------------------------------------------------------- */

typedef struct {
    float combined;
    float emb;
    float lex;
    uint32_t row;
} orc_cand;

/* Ordering::reverse of partial_cmp with unwrap_or(Equal) (:543).  Returns <0 if a
 * must come before b in a descending sort. */
static inline int cmp_desc(float a, float b)
{
    if (a > b) return -1;
    if (a < b) return 1;
    return 0; /* equal or unordered */
}

/* stable merge sort, descending by .combined (slice::sort_by is stable, :543). */
static void stable_sort_desc(orc_cand *v, orc_cand *tmp, size_t n)
{
    if (n < 2) return;
    size_t h = n / 2;
    stable_sort_desc(v, tmp, h);
    stable_sort_desc(v + h, tmp, n - h);
    size_t i = 0, j = h, k = 0;
    while (i < h && j < n) {
        /* take right only if strictly better: keeps equal elements in input order */
        if (cmp_desc(v[j].combined, v[i].combined) < 0) tmp[k++] = v[j++];
        else tmp[k++] = v[i++];
    }
    while (i < h) tmp[k++] = v[i++];
    while (j < n) tmp[k++] = v[j++];
    memcpy(v, tmp, n * sizeof(orc_cand));
}

/* "a ranks before b" under (combined desc, row asc) -- what a stable descending
 * sort of row-ordered input produces when no score is NaN. */
static inline int ranks_before(const orc_cand *a, const orc_cand *b)
{
    if (a->combined > b->combined) return 1;
    if (a->combined < b->combined) return 0;
    return a->row < b->row;
}

/* bounded selection: keep the best `cap` under ranks_before using a heap whose root
 * is the WORST kept element.  Result is then stably sorted.  Used for large n where
 * sorting n fat tuples like the reference does (:543) is pointless for the oracle;
 * tests check it against the literal full stable sort at small n. */
static void heap_sift_down(orc_cand *h, size_t n, size_t i)
{
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, w = i;
        if (l < n && ranks_before(&h[w], &h[l])) w = l; /* l is worse than w */
        if (r < n && ranks_before(&h[w], &h[r])) w = r;
        if (w == i) return;
        orc_cand t = h[i]; h[i] = h[w]; h[w] = t;
        i = w;
    }
}

static int cand_rank_cmp(const void *pa, const void *pb)
{
    const orc_cand *a = pa, *b = pb;
    if (ranks_before(a, b)) return -1;
    if (ranks_before(b, a)) return 1;
    return 0;
}

typedef struct { uint32_t row; float score; } lex_ent;
static int lex_cmp(const void *a, const void *b)
{
    uint32_t ra = ((const lex_ent *)a)->row, rb = ((const lex_ent *)b)->row;
    return ra < rb ? -1 : (ra > rb);
}

/*
 * orc_search -- src/rag_engine.rs:470-701 with self.reranker == None and
 * self.ann_index == None (the state after load_from_disk, :185,:1392).
 *
 *   rows      n x dim f32, row stride `pitch` floats; rows are used as stored
 *             (the reference normalises at insert :359 and at load :1678-1680 --
 *             the caller does that with orc_normalize).
 *   q_raw     the query embedding as returned by the embedding service; it is
 *             copied and normalised here when normalize_query != 0 (:493-494).
 *   lex_*     the (row, BM25 score) pairs LexicalIndex::score returned (:505-506);
 *             unique rows; may be empty.  Rows >= n are ignored (:525 `if let Some`).
 *   full_sort != 0 : literal path -- score every row, stable-sort all n (:543).
 *             == 0 : heap selection of the same result (large n).
 *   threads   OpenMP threads for the scoring loop (per-row arithmetic unchanged).
 * Returns the number of results written (<= max(top_k,1)); results are in final
 * order: the fallback fill (:667-698) re-sorts the initial_k candidates by
 * initial_score and takes top_k, score = initial_score.
 */
ORC_API int64_t orc_search(const float *rows, uint64_t n, uint32_t dim, uint64_t pitch,
                           const float *q_raw, int normalize_query, uint64_t top_k,
                           float w_embed, float w_lex,
                           const uint32_t *lex_rows, const float *lex_scores, uint64_t n_lex,
                           int full_sort, int threads,
                           uint32_t *out_rows, float *out_score, float *out_emb, float *out_lex)
{
    if (n == 0) return 0;                      /* :476-478 */
    if (top_k < 1) top_k = 1;                  /* :490 */

    float *q = malloc(sizeof(float) * dim);
    memcpy(q, q_raw, sizeof(float) * dim);
    if (normalize_query) orc_normalize(q, dim); /* :494 */

    /* :511-515  max_lexical = fold(0.0, f32::max).max(f32::EPSILON) */
    lex_ent *lex = NULL;
    float max_lexical = 0.0f;
    if (n_lex) {
        lex = malloc(sizeof(lex_ent) * n_lex);
        for (uint64_t i = 0; i < n_lex; ++i) {
            lex[i].row = lex_rows[i];
            lex[i].score = lex_scores[i];
            max_lexical = fmaxf(max_lexical, lex_scores[i]);
        }
        qsort(lex, n_lex, sizeof(lex_ent), lex_cmp);
    }
    max_lexical = fmaxf(max_lexical, FLT_EPSILON);

    /* :544  initial_k = min(len, max(3*top_k, top_k)) */
    uint64_t initial_k = 3 * top_k > top_k ? 3 * top_k : top_k;
    if (initial_k > n) initial_k = n;

    orc_cand *cand;
    uint64_t n_cand;

    if (full_sort) {
        cand = malloc(sizeof(orc_cand) * n);
        orc_cand *tmp = malloc(sizeof(orc_cand) * n);
        uint64_t li = 0;
        for (uint64_t r = 0; r < n; ++r) {       /* :524-541 */
            float e = dot_seq(q, rows + r * pitch, dim);
            float l = 0.0f;                      /* .unwrap_or(0.0) */
            while (li < n_lex && lex[li].row < r) ++li;
            if (li < n_lex && lex[li].row == r) l = lex[li].score / max_lexical;
            float a = w_embed * e;
            float b = w_lex * l;
            cand[r].combined = a + b;            /* :531-532 */
            cand[r].emb = e; cand[r].lex = l; cand[r].row = (uint32_t)r;
        }
        stable_sort_desc(cand, tmp, n);          /* :543 */
        free(tmp);
        n_cand = initial_k;                      /* :546-548 take(initial_k) */
    } else {
        float *emb = malloc(sizeof(float) * n);
#ifdef _OPENMP
        if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
        for (int64_t r = 0; r < (int64_t)n; ++r)
            emb[r] = dot_seq(q, rows + (uint64_t)r * pitch, dim);
        cand = malloc(sizeof(orc_cand) * (initial_k + 1));
        uint64_t hn = 0, li = 0;
        for (uint64_t r = 0; r < n; ++r) {
            float l = 0.0f;
            while (li < n_lex && lex[li].row < r) ++li;
            if (li < n_lex && lex[li].row == r) l = lex[li].score / max_lexical;
            float a = w_embed * emb[r];
            float b = w_lex * l;
            orc_cand c = { a + b, emb[r], l, (uint32_t)r };
            if (hn < initial_k) {
                cand[hn++] = c;
                if (hn == initial_k)
                    for (int64_t i = (int64_t)hn / 2 - 1; i >= 0; --i) heap_sift_down(cand, hn, (size_t)i);
            } else if (ranks_before(&c, &cand[0])) {
                cand[0] = c;
                heap_sift_down(cand, hn, 0);
            }
        }
        free(emb);
        qsort(cand, hn, sizeof(orc_cand), cand_rank_cmp);
        n_cand = hn;
    }

    /* :667-698 fallback fill (no reranker => ordered_results is empty): candidates
     * re-sorted by initial_score desc, first top_k taken.  The stable re-sort of an
     * already (combined desc, row asc)-ordered list is the identity. */
    uint64_t n_out = n_cand < top_k ? n_cand : top_k;
    for (uint64_t i = 0; i < n_out; ++i) {
        out_rows[i] = cand[i].row;
        out_score[i] = cand[i].combined;
        if (out_emb) out_emb[i] = cand[i].emb;
        if (out_lex) out_lex[i] = cand[i].lex;
    }
    free(cand); free(lex); free(q);
    return (int64_t)n_out;
}

/*
 * orc_embedding_candidates -- src/rag_engine.rs:415-461 with ann_index == None:
 * raw dot per row, stable sort descending by the raw dot (:445), take(count).
 */
ORC_API int64_t orc_embedding_candidates(const float *rows, uint64_t n, uint32_t dim, uint64_t pitch,
                                         const float *q_raw, int normalize_query, uint64_t count,
                                         int threads, uint32_t *out_rows, float *out_score)
{
    if (n == 0 || count == 0) return 0;
    /* identical to orc_search with w_embed = 1, no lexical, top_k = count, except that
     * there is no 3x cut; 1.0f * e + 0.0f * 0.0f == e exactly, so reuse is bit-exact
     * (modulo the sign of a zero, which no comparison can see). */
    uint64_t k = count > n ? n : count;
    return orc_search(rows, n, dim, pitch, q_raw, normalize_query, k, 1.0f, 0.0f,
                      NULL, NULL, 0, 0, threads, out_rows, out_score, NULL, NULL);
}

/* ------------------------------------------------------------------------- */
/* MMR                                                                        */
/* ------------------------------------------------------------------------- */

/*
 * orc_mmr -- src/rag_engine.rs:767-839, literally: O(k^2 * P * D), every
 * candidate x selected dot recomputed every round, swap_remove bookkeeping,
 * fold(0.0, max) floor, strict '>' argmax over the CURRENT order of `remaining`.
 *
 *   emb        p x dim (row stride pitch): candidate embeddings in `search` order
 *   relevance  p       : result.score of each candidate (:794)
 *   out_pos    indices into the ORIGINAL candidate list, in selection order
 * Returns the number selected.  threads > 1 parallelises the per-candidate loop of
 * one round (each candidate's arithmetic unchanged; the argmax is then resolved
 * sequentially in `remaining` order, so the result is identical).
 */
ORC_API int64_t orc_mmr(const float *emb, uint64_t pitch, const float *relevance, uint64_t p,
                        uint32_t dim, uint64_t top_k, float lambda, int threads, uint32_t *out_pos)
{
    if (p == 0) return 0;                                 /* :773-775 */
    uint32_t *remaining = malloc(sizeof(uint32_t) * p);
    uint32_t *selected = malloc(sizeof(uint32_t) * p);
    float *mmr = malloc(sizeof(float) * p);
    uint8_t *ok = malloc(p);
    uint64_t n_rem = p, n_sel = 0;
    for (uint64_t i = 0; i < p; ++i) remaining[i] = (uint32_t)i;

    /* :782-785  swap_remove(0): element 0 out, last element moves into slot 0 */
    selected[n_sel++] = remaining[0];
    remaining[0] = remaining[n_rem - 1];
    --n_rem;

    float one_minus = 1.0f - lambda;
    while (n_sel < top_k && n_rem > 0) {                  /* :788 */
#ifdef _OPENMP
        if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
#endif
        for (int64_t idx = 0; idx < (int64_t)n_rem; ++idx) {
            uint32_t c = remaining[idx];
            float rel = relevance[c];
            ok[idx] = 0;
            if (!isfinite(rel)) continue;                 /* :794-797 */
            float max_sim = 0.0f;                         /* fold(0.0_f32, max) :800-804 */
            for (uint64_t s = 0; s < n_sel; ++s) {
                float sim = dot_seq(emb + (uint64_t)c * pitch, emb + (uint64_t)selected[s] * pitch, dim);
                if (isfinite(sim)) max_sim = fmaxf(max_sim, sim);
            }
            float a = one_minus * rel;                    /* :808-809 three roundings */
            float b = lambda * max_sim;
            float m = a - b;
            mmr[idx] = m;
            ok[idx] = isfinite(m) ? 1 : 0;
        }
        float best = -INFINITY;
        uint64_t best_idx = 0;
        for (uint64_t idx = 0; idx < n_rem; ++idx)        /* :812 strict '>' */
            if (ok[idx] && mmr[idx] > best) { best = mmr[idx]; best_idx = idx; }
        if (best == -INFINITY) break;                     /* :819-822 */
        selected[n_sel++] = remaining[best_idx];          /* :825 swap_remove(best_idx) */
        remaining[best_idx] = remaining[n_rem - 1];
        --n_rem;
    }
    for (uint64_t i = 0; i < n_sel; ++i) out_pos[i] = selected[i];
    free(remaining); free(selected); free(mmr); free(ok);
    return (int64_t)n_sel;
}

/*
 * orc_search_with_diversity -- src/rag_engine.rs:717-759 (+ the API clamps of
 * src/mcp_server.rs:85-86 are the caller's business).
 * lambda is clamped to [0,1] (:725); lambda == 0 => plain search(top_k) (:728-730);
 * else pool = max(3*top_k, top_k+10) (:734), search(pool), gather embeddings
 * (:742-753), MMR to top_k (:756).
 */
ORC_API int64_t orc_search_with_diversity(const float *rows, uint64_t n, uint32_t dim, uint64_t pitch,
                                          const float *q_raw, int normalize_query, uint64_t top_k,
                                          float lambda, float w_embed, float w_lex,
                                          const uint32_t *lex_rows, const float *lex_scores, uint64_t n_lex,
                                          int full_sort, int threads,
                                          uint32_t *out_rows, float *out_score, float *out_emb, float *out_lex)
{
    /* f32::clamp: NaN stays NaN */
    if (lambda < 0.0f) lambda = 0.0f;
    if (lambda > 1.0f) lambda = 1.0f;
    if (lambda == 0.0f)
        return orc_search(rows, n, dim, pitch, q_raw, normalize_query, top_k, w_embed, w_lex,
                          lex_rows, lex_scores, n_lex, full_sort, threads,
                          out_rows, out_score, out_emb, out_lex);
    uint64_t pool = 3 * top_k > top_k + 10 ? 3 * top_k : top_k + 10;
    uint32_t *p_rows = malloc(sizeof(uint32_t) * pool);
    float *p_score = malloc(sizeof(float) * pool);
    float *p_emb = malloc(sizeof(float) * pool);
    float *p_lex = malloc(sizeof(float) * pool);
    int64_t np = orc_search(rows, n, dim, pitch, q_raw, normalize_query, pool, w_embed, w_lex,
                            lex_rows, lex_scores, n_lex, full_sort, threads,
                            p_rows, p_score, p_emb, p_lex);
    int64_t n_sel = 0;
    if (np > 0) {
        float *pe = malloc(sizeof(float) * (uint64_t)np * dim);
        for (int64_t i = 0; i < np; ++i)
            memcpy(pe + (uint64_t)i * dim, rows + (uint64_t)p_rows[i] * pitch, sizeof(float) * dim);
        uint32_t *pos = malloc(sizeof(uint32_t) * (uint64_t)np);
        n_sel = orc_mmr(pe, dim, p_score, (uint64_t)np, dim, top_k, lambda, threads, pos);
        for (int64_t i = 0; i < n_sel; ++i) {
            out_rows[i] = p_rows[pos[i]];
            out_score[i] = p_score[pos[i]];
            if (out_emb) out_emb[i] = p_emb[pos[i]];
            if (out_lex) out_lex[i] = p_lex[pos[i]];
        }
        free(pe); free(pos);
    }
    free(p_rows); free(p_score); free(p_emb); free(p_lex);
    return n_sel;
}

/* ------------------------------------------------------------------------- */
/* synthetic embeddings (bench / parity inputs)               */
/* ------------------------------------------------------------------------- */

static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* top 24 bits -> exact f32 in [-1, 1): (u24 - 2^23) / 2^23 */
static inline float hash_uniform(uint64_t seed, uint64_t idx)
{
    uint32_t u = (uint32_t)(splitmix64(seed ^ splitmix64(idx)) >> 40);
    return (float)((int32_t)u - 8388608) * (1.0f / 8388608.0f);
}

/*
 * kind 0 (iid):       x[row][c] = U(seed, row*dim + c)
 * kind 1 (clustered): x[row][c] = U(centroid_seed, (row % n_clusters)*dim + c)
 *                                 + sigma * U(seed, row*dim + c)
 * then normalize() exactly as the reference does at insert/load.
 * row0 lets a shard generate global rows [row0, row0+n).
 */
ORC_API void orc_synth_rows(float *out, uint64_t pitch, uint64_t row0, uint64_t n, uint32_t dim,
                            int kind, uint64_t seed, uint64_t centroid_seed, uint32_t n_clusters,
                            float sigma, int threads)
{
#ifdef _OPENMP
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        uint64_t row = row0 + (uint64_t)i;
        float *v = out + (uint64_t)i * pitch;
        for (uint32_t c = 0; c < dim; ++c) {
            float u = hash_uniform(seed, row * dim + c);
            if (kind == 1) {
                uint64_t cl = row % n_clusters;
                float ce = hash_uniform(centroid_seed, cl * dim + c);
                float s = sigma * u;
                u = ce + s;
            }
            v[c] = u;
        }
        for (uint64_t c = dim; c < pitch; ++c) v[c] = 0.0f;
        orc_normalize(v, dim);
    }
}

ORC_API int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
