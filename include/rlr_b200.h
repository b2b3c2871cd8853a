/*
 * rlr_b200.h -- C ABI of the B200-native retrieval hot path for rust-local-rag.
 *
 * This is the whole drop-in boundary.  The reference (CrashCartCapital/rust-local-rag)
 * has no FFI/plugin interface for this path: the boundary is three methods of
 * `RagEngine` (src/rag_engine.rs:470 `search`, :717 `search_with_diversity`, :415
 * `get_embedding_candidates`).  A `-sys` crate binds the symbols below and
 * `rag_engine.rs` keeps its method signatures; INTEGRATION.md shows the binding.
 * Each entry point cites the reference lines it replaces.
 *
 * Conventions
 *   - every function returns an `int` status (RLR_OK == 0); the message of the last
 *     failure on the calling thread is `rlr_last_error()`.  Nothing unwinds or aborts
 *     across the boundary; CUDA failures become RLR_ERR_CUDA.
 *   - plain pointers and sizes only.  Host buffers are borrowed for the call; output
 *     buffers are caller-allocated with the stated capacity.  The library owns all
 *     device memory.
 *   - rows are dense `uint32_t` positions into the caller's `row -> chunk_id` table
 *     (the library never sees chunk ids, text or metadata).
 *   - threading mirrors the reference's RwLock (src/mcp_server.rs:89,377 read lock;
 *     src/worker.rs:397-437 write lock): every `rlr_search*` / `rlr_mmr*` /
 *     `rlr_embedding_candidates` call is re-entrant on one store handle and may run
 *     concurrently from several host threads; `rlr_store_*` mutators and
 *     `rlr_store_destroy` require exclusivity.
 *   - there is NO CPU fallback: without a usable sm_100 device every compute entry
 *     point fails with RLR_ERR_NO_DEVICE.
 *   - arithmetic: scores are bit-identical to the reference's f32 arithmetic
 *     (strict left-to-right sums, separate multiply and add roundings -- see
 *     DESIGN.md); exact-score ties are ordered "lower row first".
 */
#ifndef RLR_B200_H
#define RLR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLR_ABI_VERSION 1

/* status codes */
#define RLR_OK                 0
#define RLR_ERR_INVALID_ARG    1
#define RLR_ERR_NO_DEVICE      2  /* no CUDA device / not sm_100 / extension unusable   */
#define RLR_ERR_CUDA           3  /* a CUDA runtime/driver call or kernel failed         */
#define RLR_ERR_OOM            4
#define RLR_ERR_DIM_MISMATCH   5  /* deliberate deviation from :1778 (zip truncates)     */
#define RLR_ERR_UNSUPPORTED    6
#define RLR_ERR_NONFINITE      7  /* NaN/Inf in a query or in uploaded rows              */

/* limits */
#define RLR_MAX_TOP_K          100   /* src/mcp_server.rs:364 MAX_TOP_K                   */
#define RLR_MAX_M              1024  /* max candidates one scan can return (>= 3*pool=900)*/
#define RLR_MAX_DIM            4096

typedef struct rlr_store rlr_store; /* opaque: one embedding model's chunk store on one GPU */

/* QueryWeights (src/rag_engine.rs:1846-1863): four Option<f32>.  Bit i of `has`
 * set means field i is Some(..): 0 embedding, 1 lexical, 2 reranker, 3 initial. */
typedef struct rlr_query_weights {
    float    embedding;
    float    lexical;
    float    reranker;
    float    initial;
    uint32_t has;
} rlr_query_weights;

/* ResolvedWeights (src/rag_engine.rs:1877-1896) */
typedef struct rlr_resolved_weights {
    float embedding;
    float lexical;
    float reranker;
    float initial;
} rlr_resolved_weights;

typedef struct rlr_store_info {
    uint64_t n_rows;        /* rows in this store (this shard)                           */
    uint64_t row_base;      /* global row index of local row 0 (multi-GPU shards)        */
    uint32_t dim;
    uint32_t pitch;         /* floats per stored row (dim rounded up to 32, zero padded) */
    int32_t  device;        /* CUDA ordinal                                              */
    uint32_t flags;
    uint64_t bytes_device;  /* device bytes held by the store                            */
} rlr_store_info;

typedef struct rlr_device_info {
    int32_t  device;
    int32_t  sm_count;
    int32_t  cc_major, cc_minor;
    uint64_t total_mem;
    char     name[128];
} rlr_device_info;

/* per-call stage timings, from CUDA events recorded on the call's own stream */
typedef struct rlr_timings {
    float scan_ms;          /* fused scan + per-CTA top-M kernel                          */
    float merge_ms;         /* cross-CTA merge + finalize                                 */
    float mmr_ms;           /* pairwise-similarity + greedy selection kernels             */
    float total_ms;         /* first launch to last launch on the device                  */
    uint32_t launches;      /* kernels launched by the call                               */
} rlr_timings;

/* store flags */
#define RLR_STORE_KEEP_F16      0x1u  /* keep the f32 rows AND a binary16 copy of them      */
#define RLR_STORE_CHECK_FINITE  0x2u  /* scan uploaded rows for NaN/Inf on the device      */
#define RLR_STORE_NORMALIZE_ON_UPLOAD 0x8u /* run normalize (:1763-1771) on every uploaded/appended row on the
                                            * device (same sequential arithmetic, same bits): bulk loads need not
                                            * normalise 10M rows on one host core (apply_loaded_state, :1678-1680) */
#define RLR_STORE_F16_ONLY      0x4u  /* keep only the binary16 copy (half the HBM, config 5) */
#define RLR_STORE_NO_LATENCY_PATH 0x20u /* never take the small-store latency path (below): always the copy + launch
                                         sequence + stream synchronisation that large stores use */
#define RLR_STORE_KEEP_BF16     0x10u /* keep the f32 rows AND a bfloat16 copy (operand of rlr_search_batch with
                                         RLR_BATCH_BF16; the single-query paths never read it)  */

/* Small stores (<= 262,144 rows of <= 1024 dims, f32 -- the reference's real operating point, BASELINE configs[0]) are
 * served on a LATENCY PATH by rlr_search_topm / rlr_search_mmr / rlr_embedding_candidates: the normalised query and
 * the lexical pairs travel in the kernel's parameter block (no H2D copy), rows are tiled so that every SM gets work,
 * for pools <= 32 the scan's last CTA runs merge + pairwise + greedy MMR itself (ONE launch per request), and the
 * result lands in mapped pinned host memory followed by a system-scope flag that the host polls (no D2H copy, no
 * cudaStreamSynchronize).  Same arithmetic, same bits.  RLR_NO_LATENCY_PATH=1 in the environment or
 * RLR_STORE_NO_LATENCY_PATH on the store turn it off. */

/* search flags */
#define RLR_QUERY_PRENORMALIZED 0x1u  /* skip the normalize(&mut q) of :494                */
#define RLR_WANT_TIMINGS        0x2u  /* record CUDA events; read with rlr_last_timings    */
#define RLR_SEARCH_F16          0x4u  /* scan + MMR on the binary16 copy (implied for F16_ONLY
                                         stores).  Rows are rounded to binary16 once (round to
                                         nearest even) and widened exactly; all arithmetic stays
                                         the reference's sequential f32, so results are bit-identical
                                         to the reference run on the rounded rows.  Measured
                                         deviation from the f32 store: DESIGN.md "f16 store".   */

/* synthetic fill kinds (bench / parity inputs) */
#define RLR_SYNTH_IID           0
#define RLR_SYNTH_CLUSTERED     1

/* ---- library ----------------------------------------------------------------- */

int         rlr_abi_version(void);
const char *rlr_last_error(void);                 /* thread-local, never NULL             */
int         rlr_device_count(int *out_count);
int         rlr_device_query(int device, rlr_device_info *out);

/* ---- host-side scalar helpers (bit-exact mirrors, for glue and tests) ---------- */

/* normalize, src/rag_engine.rs:1763-1771 (in place). */
int rlr_normalize(float *v, size_t n);

/* ResolvedWeights::from_query_weights, src/rag_engine.rs:1869-1896, defaults
 * :1801-1804, env cache RAG_*_WEIGHT :1806-1841 (read once per process). */
int rlr_resolve_weights(const rlr_query_weights *overrides /* nullable */,
                        rlr_resolved_weights *out);

/* ---- chunk store ------------------------------------------------------------- */

/*
 * Create a device-resident store from host rows.  Replaces the embedding half of
 * `HashMap<String, DocumentChunk>` (src/rag_engine.rs:105) after `apply_loaded_state`
 * (:1655-1696).  `rows` is n_rows x dim f32, row stride `host_pitch` floats (0 =>
 * dim), ALREADY normalised by the host exactly as :1678-1680 does (so device rows
 * are bit-identical to what the reference would scan).  rows may be NULL to create
 * an uninitialised store that is then filled by rlr_store_upload / rlr_store_fill_synthetic.
 * row_base: global index of row 0 when this store is one shard of a row-sharded corpus.
 */
int rlr_store_create(int device, uint32_t dim, uint64_t n_rows, const float *rows,
                     uint64_t host_pitch, uint64_t row_base, uint32_t flags,
                     rlr_store **out);
int rlr_store_destroy(rlr_store *s);
int rlr_store_info_get(const rlr_store *s, rlr_store_info *out);

/* overwrite rows [row0, row0+n) from host memory (normalised by the caller). */
int rlr_store_upload(rlr_store *s, uint64_t row0, uint64_t n, const float *rows,
                     uint64_t host_pitch);

/*
 * Store mutation for `add_document` (src/rag_engine.rs:347-386: `chunks.retain(..)` drops the
 * document's old chunks, then the new normalised chunks are inserted), called under the
 * reference's write lock -- exclusivity required.
 *   rlr_store_reserve      pre-size the device allocations (amortises appends).
 *   rlr_store_append       append n normalised rows; *out_first_row = global row of the first.
 *   rlr_store_remove_rows  delete the listed global rows.  The store stays dense: each hole is
 *                          filled with a surviving row from the tail.  The moves are reported as
 *                          (out_moved_from[i] -> out_moved_to[i]), i < *out_n_moved <= n, so that
 *                          the caller can patch its `row -> chunk_id` table; out arrays have capacity n.
 */
int rlr_store_reserve(rlr_store *s, uint64_t capacity_rows);
int rlr_store_append(rlr_store *s, uint64_t n, const float *rows, uint64_t host_pitch,
                     uint64_t *out_first_row);
int rlr_store_remove_rows(rlr_store *s, const uint32_t *rows, uint64_t n,
                          uint32_t *out_moved_from, uint32_t *out_moved_to, uint64_t *out_n_moved);

/* copy rows[i] (local indices) back to host: out is n x dim, stride dim. */
int rlr_store_read_rows(const rlr_store *s, const uint32_t *rows, uint64_t n, float *out);

/* Fill local rows [0,n_rows) with synthetic unit vectors for GLOBAL rows
 * row_base.. (bit-reproducible on the CPU; normalised on the device with the
 * reference's sequential arithmetic). */
int rlr_store_fill_synthetic(rlr_store *s, int kind, uint64_t seed, uint64_t centroid_seed,
                             uint32_t n_clusters, float sigma);

/* ---- the hot path ------------------------------------------------------------ */

/*
 * rlr_search_topm -- RagEngine::search up to the candidate cut, reranker absent:
 * src/rag_engine.rs:476-565 (+ the fallback ordering :667-698).
 *   query      dim f32 as returned by the embedding service; normalised here (:494)
 *              unless RLR_QUERY_PRENORMALIZED
 *   w          resolved weights (embedding, lexical used)
 *   lex_rows / lex_scores / n_lex
 *              what LexicalIndex::score(query, 5*top_k) returned (:505-506): unique
 *              local rows with raw BM25 scores; normalised by max (:511-530) here
 *   m          number of candidates wanted, 1..RLR_MAX_M.  search(top_k) semantics:
 *              m = top_k for the reranker-absent result, m = 3*top_k to feed a host
 *              reranker (:544)
 * Outputs (capacity m each; any may be NULL except out_rows/out_n): rows in
 * (combined desc, row asc) order with combined (= initial_score), embedding_score,
 * lexical_score.  *out_n = min(m, n_rows).  Empty store => RLR_OK, *out_n = 0 (:476).
 */
int rlr_search_topm(rlr_store *s, const float *query, uint32_t dim, uint32_t flags,
                    const rlr_resolved_weights *w,
                    const uint32_t *lex_rows, const float *lex_scores, uint32_t n_lex,
                    uint32_t m,
                    uint32_t *out_rows, float *out_combined, float *out_emb, float *out_lex,
                    uint32_t *out_n);

/*
 * rlr_mmr -- RagEngine::mmr_diversify, src/rag_engine.rs:767-839, on rows resident
 * in the store (replaces the embedding lookup :742-753 as well).
 *   cand_rows  p local rows in `search` order;  relevance: their result.score (:794;
 *              caller-supplied so that a host reranker can sit in between)
 *   top_k      selections wanted (the first candidate is always selected, :782-785)
 *   lambda     diversity_factor, used as given (clamp is the caller's, :725)
 * out_sel_pos (capacity min(p, max(top_k,1))): positions into cand_rows, selection
 * order.  p <= RLR_MAX_M.
 */
int rlr_mmr(rlr_store *s, const uint32_t *cand_rows, const float *relevance, uint32_t p,
            uint32_t top_k, float lambda, uint32_t flags,
            uint32_t *out_sel_pos, uint32_t *out_n);

/*
 * rlr_search_mmr -- RagEngine::search_with_diversity with no reranker,
 * src/rag_engine.rs:717-759: lambda clamped to [0,1]; lambda == 0 => search(top_k);
 * else pool = max(3*top_k, top_k+10), search(pool), MMR to top_k.  One launch
 * sequence, one host<->device round trip.  This is the benchmarked entry point.
 * Outputs have capacity max(top_k, 1).
 */
int rlr_search_mmr(rlr_store *s, const float *query, uint32_t dim, uint32_t flags,
                   uint32_t top_k, float diversity_factor,
                   const rlr_resolved_weights *w,
                   const uint32_t *lex_rows, const float *lex_scores, uint32_t n_lex,
                   uint32_t *out_rows, float *out_score, float *out_emb, float *out_lex,
                   uint32_t *out_n);

/*
 * rlr_search_mmr_multi -- THROUGHPUT MODE: 2..RLR_MAX_MULTI independent search_with_diversity calls (:717-759)
 * answered by ONE pass over the rows.  What it models: several searches in flight under the reference's read lock
 * (src/mcp_server.rs:89,377).  The scan kernel runs one group of consumer warps per query over the SAME shared-
 * memory tiles, so every row crosses HBM once per call instead of once per query; each (row, query) dot is still
 * its own strict left-to-right f32 chain, and every query's result is bit-identical to its rlr_search_mmr result.
 *   queries      nq x dim f32 (row stride dim), each normalised here unless RLR_QUERY_PRENORMALIZED
 *   lex_*        per query q: n_lex[q] pairs at lex_rows[q] / lex_scores[q] (any may be NULL / 0); lex_rows, lex_scores
 *                and n_lex themselves may be NULL for embedding-only queries
 * Outputs: out_* are nq x max(top_k,1) (row stride max(top_k,1)), out_n[q] the count of query q.
 * The latency of one call is that of a single query plus (nq - 1) MMR tails; use it when queries queue up.
 */
#define RLR_MAX_MULTI 3
int rlr_search_mmr_multi(rlr_store *s, const float *queries, uint32_t nq, uint32_t dim, uint32_t flags,
                         uint32_t top_k, float diversity_factor, const rlr_resolved_weights *w,
                         const uint32_t *const *lex_rows, const float *const *lex_scores, const uint32_t *n_lex,
                         uint32_t *out_rows, float *out_score, float *out_emb, float *out_lex,
                         uint32_t *out_n);

/*
 * rlr_embedding_candidates -- RagEngine::get_embedding_candidates,
 * src/rag_engine.rs:415-461 (ann_index == None): top `count` by raw dot product.
 */
int rlr_embedding_candidates(rlr_store *s, const float *query, uint32_t dim, uint32_t flags,
                             uint32_t count, uint32_t *out_rows, float *out_score,
                             uint32_t *out_n);

/*
 * rlr_search_batch -- many queries at once (not in the reference, which answers one query
 * per call; BASELINE config 4 / north_star kernel (3)).  For each of n_queries embeddings:
 * the top m rows by embedding score (the ordering of get_embedding_candidates, :445), computed
 * as ONE dense contraction on the tensor cores: tcgen05.mma over the binary16 / bfloat16 copy of
 * the store or (tf32) over the f32 rows themselves, f32 accumulation in TMEM, a per-query
 * threshold filter fused into the epilogue.
 *   queries     n_queries x dim f32, row stride dim; normalised here (:494) unless
 *               RLR_QUERY_PRENORMALIZED; converted to the operand precision for the contraction
 *   m           1..RLR_MAX_M
 *   flags       RLR_BATCH_EXACT_RESCORE: re-score the m shortlisted rows of every query with the
 *               reference's sequential f32 arithmetic (on the f32 rows when the store has them)
 *               and order by that; out_scores are then bit-identical to the single-query path.
 *               Without it out_scores are the tensor-core values (<= ~2e-4 absolute from exact
 *               on unit vectors) and the order is by those.
 * Outputs: out_rows / out_scores are n_queries x m (row stride m), out_n[q] = min(m, n_rows).
 */
#define RLR_BATCH_EXACT_RESCORE 0x8u
/* operand precision of the contraction (at most one; default: binary16 when the store keeps that copy,
 * else tf32).  All accumulate in f32 in TMEM.
 *   RLR_BATCH_F16   tcgen05.mma kind::f16 over the binary16 copy: 10-bit mantissas, |score - exact| <= ~2e-4
 *   RLR_BATCH_BF16  kind::f16 with bfloat16 operands over the RLR_STORE_KEEP_BF16 copy: 7-bit mantissas,
 *                   same speed as binary16, <= ~2e-3 absolute on unit vectors
 *   RLR_BATCH_TF32  kind::tf32 straight over the f32 rows (NO second copy of the store): the tensor core reads
 *                   the upper 19 bits of every f32 (10-bit mantissas, truncated), <= ~1e-3 absolute on unit
 *                   vectors; half the tensor rate and twice the operand bytes of the 16-bit kinds
 * Tolerances are asserted in tests/test_gpu_batch.py; RLR_BATCH_EXACT_RESCORE restores exact bits in every mode. */
#define RLR_BATCH_F16           0x10u
#define RLR_BATCH_BF16          0x20u
#define RLR_BATCH_TF32          0x40u
int rlr_search_batch(rlr_store *s, const float *queries, uint32_t n_queries, uint32_t dim,
                     uint32_t flags, uint32_t m,
                     uint32_t *out_rows, float *out_scores, uint32_t *out_n);

/* Multi-GPU composition of the batched path (BASELINE config 4: rows sharded over the GPUs of
 * one box).  rlr_search_batch_device is rlr_search_batch with the result left in HBM: d_keys is
 * n_queries x m u64 rank keys ((ordered(score) << 32) | ~global_row, 0 = empty) in rank order,
 * d_cnt (nullable) n_queries u32; both are valid on `stream` when the call returns.  After an
 * all-gather of every rank's d_keys into [n_lists][n_queries][m], rlr_batch_merge_async merges
 * them per query (one CTA per query) into the global best m. */
int rlr_search_batch_device(rlr_store *s, const float *queries, uint32_t n_queries, uint32_t dim,
                            uint32_t flags, uint32_t m, void *d_keys, void *d_cnt /* nullable */,
                            void *stream);
int rlr_batch_merge_async(rlr_store *s, const void *d_lists, uint32_t n_lists, uint32_t n_queries,
                          uint32_t m, void *d_out_keys, void *d_out_cnt /* nullable */, void *stream);

/* ---- BM25 on the device: LexicalIndex (src/rag_engine.rs:2083-2237) for text queries -- SURVEY.md 8(f) N4 ----------
 *
 * RagEngine::search blends a BM25 score into every text query (:505-532).  On the host that pass costs 18 ms per query
 * at 10k chunks and 0.2 s at 100k (tools/bm25_cost.py: hash-map postings like the reference's) -- two to three orders
 * of magnitude more than the whole GPU search.  An rlr_bm25 index keeps the NUMERIC half of LexicalIndex on the device
 * and scores it there; strings stay on the host: the caller tokenizes (`tokenize`, :2242-2247) and owns the
 * term -> id dictionary.
 *   rlr_bm25_set_doc      add_chunk (:2106-2138) for the chunk at `row` (a global row of the store): its distinct term
 *                         ids with their counts; replaces what the row held; no terms => the row is not indexed
 *   rlr_bm25_remove_doc   remove_chunk (:2140-2167)
 *   rlr_bm25_move_doc     follow rlr_store_remove_rows: `to_row` takes over the document of `from_row`
 *   rlr_bm25_score        LexicalIndex::score(query, limit) (:2169-2227): the `limit` best (row, score) by score desc.
 *                         query_terms: the ids of the query's tokens that exist in the dictionary, in BYTEWISE ORDER OF
 *                         THE TERM STRINGS (that order fixes the f32 summation order per document; the reference's is a
 *                         random HashSet order); duplicates are ignored.  Exact-score ties go to the lower row.
 *   rlr_search_text_topm / rlr_search_text_mmr
 *                         RagEngine::search / search_with_diversity for a TEXT query: BM25 scoring, the top 5*top_k
 *                         selection, max-normalisation (:511-530), the embedding scan with the blend, top-k and MMR all
 *                         run on the device on one stream with no host round trip in between.
 * Scores are bit-identical to the reference's f32 formula evaluated in that term order (idf's ln() is the C library's
 * logf, computed on the host per query term).  Mutators need exclusivity like the store's; scoring is re-entrant.
 * The device image is rebuilt lazily by the first query after a mutation (a forward index for stores of up to 262,144
 * rows, postings by term beyond that: 0.42 ms for LexicalIndex::score over 10M chunks / 500M postings on one GPU).
 * An index is bound to its store: use it only while the store lives (destroying it afterwards is allowed).
 * rlr_bm25 serves one single-GPU store; rlr_cluster_bm25 (below) is the same index over a cluster. */
typedef struct rlr_bm25 rlr_bm25;
int rlr_bm25_create(rlr_store *s, rlr_bm25 **out);
int rlr_bm25_destroy(rlr_bm25 *ix);
int rlr_bm25_set_doc(rlr_bm25 *ix, uint32_t row, const uint32_t *term_ids, const uint32_t *term_freqs, uint32_t n_terms);
/* bulk rlr_bm25_set_doc for rows [row0, row0 + n_docs): CSR, offsets[n_docs + 1] index term_ids / term_freqs
 * (what a loader does for every chunk at start-up, validate_index_sync :1375-1389) */
int rlr_bm25_set_docs(rlr_bm25 *ix, uint32_t row0, uint32_t n_docs, const uint64_t *offsets,
                      const uint32_t *term_ids, const uint32_t *term_freqs);
int rlr_bm25_remove_doc(rlr_bm25 *ix, uint32_t row);
int rlr_bm25_move_doc(rlr_bm25 *ix, uint32_t from_row, uint32_t to_row);
int rlr_bm25_stats(const rlr_bm25 *ix, uint64_t *total_docs, uint64_t *total_length, uint64_t *n_terms);
int rlr_bm25_score(rlr_bm25 *ix, const uint32_t *query_terms, uint32_t n_terms, uint32_t limit,
                   uint32_t *out_rows, float *out_scores, uint32_t cap, uint32_t *out_n);
int rlr_search_text_topm(rlr_store *s, rlr_bm25 *ix, const float *query, uint32_t dim, uint32_t flags,
                         const rlr_resolved_weights *w, const uint32_t *query_terms, uint32_t n_terms, uint32_t m,
                         uint32_t *out_rows, float *out_combined, float *out_emb, float *out_lex, uint32_t *out_n);
int rlr_search_text_mmr(rlr_store *s, rlr_bm25 *ix, const float *query, uint32_t dim, uint32_t flags,
                        uint32_t top_k, float diversity_factor, const rlr_resolved_weights *w,
                        const uint32_t *query_terms, uint32_t n_terms,
                        uint32_t *out_rows, float *out_score, float *out_emb, float *out_lex, uint32_t *out_n);

/* (The host-side BM25 twin that the non-Rust host mirrors use for text queries lives in
 * include/rlr_hostmirror.h / librlr_hostmirror.so: host-mirror support, not part of this boundary.) */

/* stage timings of the calling thread's most recent call made with RLR_WANT_TIMINGS */
int rlr_last_timings(rlr_timings *out);

/* ---- one process, every GPU of the box: the same calls over a row-sharded store ---------------
 *
 * The reference is ONE process holding ONE `Arc<RwLock<RagEngine>>` (src/main.rs:140-167; callers take
 * `rag_state.read()`, src/mcp_server.rs:89,377).  An rlr_cluster is that engine's embedding store sharded
 * by contiguous row blocks over several GPUs of one NVSwitch box, driven from the calling host thread:
 * no second process, no torchrun, no NCCL.  Every rlr_cluster_* search has exactly the arguments,
 * semantics and results (bit for bit, rows are GLOBAL positions) of its rlr_* single-store namesake:
 *   - per query the host thread enqueues one scan kernel per GPU; each kernel's last CTA stores the
 *     shard's top-m list straight into a mailbox in shard 0's HBM (peer stores over NVLink, enabled with
 *     cudaDeviceEnablePeerAccess) and publishes a sequence number with a system-scope release;
 *   - shard 0's GPU merges inside a kernel that waits for the flags, and its MMR kernels load the pool
 *     rows from the owning GPUs' HBM through peer pointers -- no collective, no host round trip between
 *     the stages, one stream synchronisation per query;
 *   - the lexical pairs (LexicalIndex::score output, :505-506) are normalised by their GLOBAL maximum
 *     (:511-515) and routed to the shard that owns each row.
 * Re-entrant like the single-store calls (each concurrent caller leases its own per-GPU workspaces,
 * streams and mailbox).  A cluster is a bulk-loaded snapshot (load_from_disk, :1520-1696): there is no
 * rlr_cluster_append / remove; mutate a single-GPU store (rlr_store_append / rlr_store_remove_rows) or
 * rebuild the cluster.
 *   devices     n_devices CUDA ordinals; entry 0 is the root (mailbox, merge, MMR).  The same ordinal may
 *               appear more than once (several shards on one GPU: how the path is tested on a 1-GPU box).
 *   shard_rows  rows per shard (sum == n_rows, every entry > 0) or NULL for the default plan: even blocks,
 *               except that the root gets fewer rows so that its scan + merge/MMR tail takes as long as the
 *               other GPUs' scans (tail-balanced sharding).
 * Fails with RLR_ERR_UNSUPPORTED when the GPUs cannot reach each other's memory (no peer access). */
#define RLR_MAX_SHARDS 16
typedef struct rlr_cluster rlr_cluster;
typedef struct rlr_cluster_info {
    uint64_t n_rows;
    uint32_t dim;
    uint32_t pitch;
    uint32_t flags;
    uint32_t n_shards;
    int32_t  device[RLR_MAX_SHARDS];
    uint64_t row_base[RLR_MAX_SHARDS];
    uint64_t shard_rows[RLR_MAX_SHARDS];
} rlr_cluster_info;

int rlr_cluster_create(const int *devices, uint32_t n_devices, uint32_t dim, uint64_t n_rows,
                       const float *rows /* nullable */, uint64_t host_pitch, uint32_t flags,
                       const uint64_t *shard_rows /* nullable */, rlr_cluster **out);
int rlr_cluster_destroy(rlr_cluster *c);
int rlr_cluster_info_get(const rlr_cluster *c, rlr_cluster_info *out);
/* overwrite GLOBAL rows [row0, row0+n) (routed to the owning shards); rows as for rlr_store_upload */
int rlr_cluster_upload(rlr_cluster *c, uint64_t row0, uint64_t n, const float *rows, uint64_t host_pitch);
int rlr_cluster_read_rows(const rlr_cluster *c, const uint32_t *rows, uint64_t n, float *out);
int rlr_cluster_fill_synthetic(rlr_cluster *c, int kind, uint64_t seed, uint64_t centroid_seed,
                               uint32_t n_clusters, float sigma);
/* RagEngine::search (:476-565 + :667-698), as rlr_search_topm */
int rlr_cluster_search_topm(rlr_cluster *c, const float *query, uint32_t dim, uint32_t flags,
                            const rlr_resolved_weights *w,
                            const uint32_t *lex_rows, const float *lex_scores, uint32_t n_lex, uint32_t m,
                            uint32_t *out_rows, float *out_combined, float *out_emb, float *out_lex,
                            uint32_t *out_n);
/* RagEngine::mmr_diversify (:767-839) over GLOBAL candidate rows, as rlr_mmr */
int rlr_cluster_mmr(rlr_cluster *c, const uint32_t *cand_rows, const float *relevance, uint32_t p,
                    uint32_t top_k, float lambda, uint32_t flags, uint32_t *out_sel_pos, uint32_t *out_n);
/* RagEngine::search_with_diversity (:717-759), as rlr_search_mmr: the benchmarked multi-GPU entry point */
int rlr_cluster_search_mmr(rlr_cluster *c, const float *query, uint32_t dim, uint32_t flags,
                           uint32_t top_k, float diversity_factor, const rlr_resolved_weights *w,
                           const uint32_t *lex_rows, const float *lex_scores, uint32_t n_lex,
                           uint32_t *out_rows, float *out_score, float *out_emb, float *out_lex,
                           uint32_t *out_n);
/* throughput mode over the cluster, as rlr_search_mmr_multi (same arguments, per-query lexical pairs included): every
 * GPU scans its shard ONCE for the nq queries and posts nq lists; the root merges and diversifies each */
int rlr_cluster_search_mmr_multi(rlr_cluster *c, const float *queries, uint32_t nq, uint32_t dim, uint32_t flags,
                                 uint32_t top_k, float diversity_factor, const rlr_resolved_weights *w,
                                 const uint32_t *const *lex_rows, const float *const *lex_scores, const uint32_t *n_lex,
                                 uint32_t *out_rows, float *out_score, float *out_emb, float *out_lex,
                                 uint32_t *out_n);
/* the batched contraction over the cluster, as rlr_search_batch (flags: operand precision, exact re-score): every GPU
 * contracts the batch against its shard, the root pulls the per-shard key lists over NVLink and merges them per query */
int rlr_cluster_search_batch(rlr_cluster *c, const float *queries, uint32_t n_queries, uint32_t dim, uint32_t flags,
                             uint32_t m, uint32_t *out_rows, float *out_scores, uint32_t *out_n);
/* RagEngine::get_embedding_candidates (:415-461), as rlr_embedding_candidates */
int rlr_cluster_embedding_candidates(rlr_cluster *c, const float *query, uint32_t dim, uint32_t flags,
                                     uint32_t count, uint32_t *out_rows, float *out_score, uint32_t *out_n);
/* per-shard scan-kernel durations (ms) of the calling thread's most recent rlr_cluster_search_* call made
 * with RLR_WANT_TIMINGS (rlr_last_timings has the maximum as scan_ms and the root's merge / MMR stages) */
int rlr_cluster_last_scan_ms(float *out_ms, uint32_t cap, uint32_t *out_n);
/* kernels launched by this cluster's searches since creation (bench `gpu_launches`) */
int rlr_cluster_launch_count(const rlr_cluster *c, uint64_t *out);

/* BM25 over the cluster (LexicalIndex, src/rag_engine.rs:2083-2237, for a corpus sharded over several GPUs): one
 * device index per shard.  A query is scored on EVERY shard's GPU at once with the statistics of the whole corpus
 * (N, average length, df -- so a document's score does not depend on the sharding), each shard ranks its own `limit`
 * best on the device, and the host merges those short lists (<= 16 x limit records) into the global `limit` best:
 * exactly what one index over all rows returns.  rlr_cluster_search_text_* then run the cluster search with those
 * pairs -- the results equal rlr_search_text_* on one store holding the same rows, bit for bit.  `row` arguments are
 * global rows.  With one shard everything is forwarded to the single-GPU entry points. */
typedef struct rlr_cluster_bm25 rlr_cluster_bm25;
int rlr_cluster_bm25_create(rlr_cluster *c, rlr_cluster_bm25 **out);
int rlr_cluster_bm25_destroy(rlr_cluster_bm25 *ix);
int rlr_cluster_bm25_set_doc(rlr_cluster_bm25 *ix, uint32_t row, const uint32_t *term_ids, const uint32_t *term_freqs, uint32_t n_terms);
int rlr_cluster_bm25_set_docs(rlr_cluster_bm25 *ix, uint32_t row0, uint32_t n_docs, const uint64_t *offsets,
                              const uint32_t *term_ids, const uint32_t *term_freqs);
int rlr_cluster_bm25_remove_doc(rlr_cluster_bm25 *ix, uint32_t row);
int rlr_cluster_bm25_stats(const rlr_cluster_bm25 *ix, uint64_t *total_docs, uint64_t *total_length, uint64_t *n_terms);
int rlr_cluster_bm25_score(rlr_cluster_bm25 *ix, const uint32_t *query_terms, uint32_t n_terms, uint32_t limit,
                           uint32_t *out_rows, float *out_scores, uint32_t cap, uint32_t *out_n);
int rlr_cluster_search_text_topm(rlr_cluster *c, rlr_cluster_bm25 *ix, const float *query, uint32_t dim, uint32_t flags,
                                 const rlr_resolved_weights *w, const uint32_t *query_terms, uint32_t n_terms, uint32_t m,
                                 uint32_t *out_rows, float *out_combined, float *out_emb, float *out_lex, uint32_t *out_n);
int rlr_cluster_search_text_mmr(rlr_cluster *c, rlr_cluster_bm25 *ix, const float *query, uint32_t dim, uint32_t flags,
                                uint32_t top_k, float diversity_factor, const rlr_resolved_weights *w,
                                const uint32_t *query_terms, uint32_t n_terms,
                                uint32_t *out_rows, float *out_score, float *out_emb, float *out_lex, uint32_t *out_n);

/* ---- device-level building blocks (multi-GPU composition, bench `value`) -------
 *
 * Same kernels, but inputs/outputs stay in device memory and work is enqueued on a
 * caller-supplied CUDA stream (`cudaStream_t` passed as void*; device pointers as
 * void*).  No host synchronisation.  Used by the one-process-per-GPU path to place
 * the NCCL all-gather between the local top-M and the merge + MMR, and by bench.py
 * to time the path with inputs already resident in HBM.
 */
typedef struct rlr_ctx rlr_ctx; /* per-caller workspace bound to one store */

int rlr_ctx_create(rlr_store *s, rlr_ctx **out);
int rlr_ctx_destroy(rlr_ctx *c);

/* Candidate record exchanged between GPUs: 16 bytes. */
typedef struct rlr_cand {
    uint64_t key;       /* (ordered(combined) << 32) | ~global_row : larger == ranks earlier */
    float    emb;       /* embedding_score                                                   */
    float    lex;       /* lexical_score                                                     */
} rlr_cand;

/* scan + local top-m: d_query = dim normalised f32 on the device; d_lex_* sorted by
 * row (may be NULL); writes m records (rank order, padded with key 0) to d_out and
 * the valid count to d_out_n.  Rows in the keys are GLOBAL (row_base + local). */
int rlr_topm_async(rlr_ctx *c, const void *d_query, float w_embed, float w_lex,
                   const void *d_lex_rows, const void *d_lex_norm, uint32_t n_lex,
                   uint32_t m, void *d_out /* rlr_cand[m] */, void *d_out_n /* u32 */,
                   void *stream);

/* merge n_lists lists of m records each (e.g. the all-gathered per-GPU lists) into
 * the best m, same record format. */
int rlr_merge_async(rlr_ctx *c, const void *d_lists, uint32_t n_lists, uint32_t m,
                    void *d_out, void *d_out_n, void *stream);

/* gather the embeddings of the records that this store owns into d_out
 * (m x pitch f32, zero rows for records owned by another shard). */
int rlr_gather_async(rlr_ctx *c, const void *d_cands, const void *d_n, uint32_t m,
                     void *d_out, void *stream);

/* MMR over a dense candidate matrix in device memory (p x pitch f32) with relevance
 * = combined score decoded from d_cands.  Writes selected positions / count and, when
 * d_result != NULL, the selected records in selection order. */
int rlr_mmr_async(rlr_ctx *c, const void *d_emb, uint32_t pitch, uint32_t dim,
                  const void *d_cands, const void *d_n, uint32_t p_cap,
                  uint32_t top_k, float lambda,
                  void *d_sel_pos /* u32[top_k] */, void *d_sel_n /* u32 */,
                  void *d_result /* rlr_cand[top_k], nullable */, void *stream);

/* MMR over candidates that are rows of THIS store (keys carry global rows): no gather. */
int rlr_mmr_store_async(rlr_ctx *c, const void *d_cands, const void *d_n, uint32_t p_cap,
                        uint32_t top_k, float lambda, void *d_sel_pos, void *d_sel_n,
                        void *d_result /* nullable */, void *stream);

/* ---- peer memory: MMR on rank 0 reads pool rows straight from the owning GPUs' HBM ------------
 *
 * Row shards of one corpus on the GPUs of one NVSwitch box.  Each process exports its shard
 * (CUDA IPC), rank 0 opens all of them, and rlr_mmr_peers_async runs MMR over GLOBAL candidate
 * rows: the pairwise-similarity kernel dereferences peer pointers (NVLink loads), so no gather
 * collective and no staging copy are needed.  Shards must agree on dim and precision. */
#define RLR_IPC_HANDLE_BYTES 64
typedef struct rlr_peer_set rlr_peer_set;

int rlr_store_ipc_export(const rlr_store *s, uint32_t search_flags, void *handle_out /* 64 bytes */);
int rlr_peer_set_open(rlr_store *local, uint32_t my_index, uint32_t n_shards,
                      const void *handles /* n_shards x 64 bytes */, const uint64_t *row_base,
                      const uint64_t *n_rows, uint32_t search_flags, rlr_peer_set **out);
int rlr_peer_set_close(rlr_peer_set *p);
int rlr_mmr_peers_async(rlr_ctx *c, rlr_peer_set *p, const void *d_cands, const void *d_n, uint32_t p_cap,
                        uint32_t top_k, float lambda, void *d_sel_pos, void *d_sel_n,
                        void *d_result /* nullable */, void *stream);

/* ---- fused exchange: the scan kernel delivers its list into the root GPU's HBM ------------
 * Replaces the all-gather of SURVEY.md 8(e): the root rank owns a mailbox (a ring of slots,
 * one list of m_cap records per rank and slot); every other rank maps it through CUDA IPC.
 * rlr_topm_post_async runs the same scan + top-m as rlr_topm_async, but the kernel's last CTA
 * stores the merged list straight into slot (seq % ring) of the mailbox -- NVLink peer stores
 * issued by the compute kernel itself -- and then publishes `seq` with a system-scope release.
 * rlr_mailbox_merge_async (root only) waits inside its kernel for the n_ranks flags of `seq`,
 * merges the lists to the global best m and frees the slot.  No collective call, no host
 * round trip; ranks other than the root are done as soon as their scan is.
 * Sequence numbers start at 1, increase by 1 per query and must agree across ranks; a slot is
 * reused after `ring` queries, and a posting kernel waits (bounded, 4 s) for the root to have
 * consumed it.  Ranks that mapped the mailbox must rlr_mailbox_close it before the root frees it.
 * (One `consumed` word per slot, so queries may be in flight on several streams /
 * ctxs of a rank at once as long as `ring` is a multiple of the number of such lanes).  rlr_mailbox_status reads a sticky word: non-zero once any wait timed out. */
typedef struct rlr_mailbox rlr_mailbox;
int rlr_mailbox_create(int device, uint32_t n_ranks, uint32_t m_cap, uint32_t ring, rlr_mailbox **out);
int rlr_mailbox_ipc_export(const rlr_mailbox *mb, void *handle_out /* 64 bytes */);
int rlr_mailbox_open(int device, const void *handle /* 64 bytes */, uint32_t n_ranks, uint32_t m_cap,
                     uint32_t ring, rlr_mailbox **out);
int rlr_mailbox_close(rlr_mailbox *mb);
int rlr_mailbox_status(rlr_mailbox *mb, uint32_t *out);
int rlr_topm_post_async(rlr_ctx *c, rlr_mailbox *mb, uint32_t my_rank, uint64_t seq, const void *d_query,
                        float w_embed, float w_lex, const void *d_lex_rows, const void *d_lex_norm,
                        uint32_t n_lex, uint32_t m, void *stream);
int rlr_mailbox_merge_async(rlr_ctx *c, rlr_mailbox *mb, uint64_t seq, uint32_t m, void *d_out,
                            void *d_out_n, void *stream);

/* fused single-GPU search_with_diversity on the device: results stay in HBM.
 * d_result: rlr_cand[max(top_k,1)] in selection order, d_result_n: u32. */
int rlr_search_mmr_async(rlr_ctx *c, const void *d_query, uint32_t top_k, float diversity_factor,
                         float w_embed, float w_lex,
                         void *d_result, void *d_result_n, void *stream);

/* device-level form of rlr_search_mmr_multi: nq ctxs of ONE store (ctx q holds query q's pool / MMR buffers; the scan
 * uses ctx 0's launch workspace), d_queries[q] normalised on the device, everything enqueued on `stream`.
 * d_results[q]: rlr_cand[max(top_k,1)], d_result_ns[q]: u32. */
int rlr_search_mmr_multi_async(rlr_ctx *const *ctxs, uint32_t nq, const void *const *d_queries, uint32_t top_k,
                               float diversity_factor, float w_embed, float w_lex,
                               void *const *d_results, void *const *d_result_ns, void *stream);

/* search flags (RLR_SEARCH_F16) used by the device-level entry points of this ctx */
int rlr_ctx_set_flags(rlr_ctx *c, uint32_t search_flags);

/* launches enqueued by this ctx since creation (bench `gpu_launches`) */
int rlr_ctx_launch_count(const rlr_ctx *c, uint64_t *out);

/* Time the scan kernel alone: `iters` back-to-back launches on `stream`, CUDA
 * events around them on that stream; returns mean ms per launch (roofline). */
int rlr_time_scan(rlr_ctx *c, const void *d_query, uint32_t m, uint32_t iters,
                  void *stream, float *out_ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* RLR_B200_H */
