/*
 * rlr_hostmirror.h -- HOST-MIRROR SUPPORT (librlr_hostmirror.so, plain C++, no CUDA): a twin of the reference's
 * LexicalIndex / tokenize, /root/reference/src/rag_engine.rs:2083-2247.
 *
 * NOT part of the drop-in boundary (include/rlr_b200.h).  The reference keeps its BM25 index on the host and a
 * Rust integration keeps using it: its output -- what `lexical_index.score(query, top_k * 5)` returned, :505 --
 * is the `lex_rows` / `lex_scores` input of rlr_search_topm / rlr_search_mmr / rlr_cluster_*.  This library only
 * lets the non-Rust host mirrors shipped with this repository (include/rlr_engine.hpp, engine.py) answer text
 * queries in their tests.  Chunks are identified by caller-chosen u64 keys.  Deterministic where the reference is
 * not: query terms are summed in bytewise order, score ties go to the smaller key.
 * rlr_tokenize returns the tokens of `text` joined by '\n' (:2242-2247).
 */
#ifndef RLR_HOSTMIRROR_H
#define RLR_HOSTMIRROR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLR_HM_OK               0
#define RLR_HM_ERR_INVALID_ARG  1
#define RLR_HM_ERR_UNSUPPORTED  6

typedef struct rlr_lexical rlr_lexical;
const char *rlr_hostmirror_last_error(void);
int rlr_lexical_create(rlr_lexical **out);
int rlr_lexical_destroy(rlr_lexical *lx);
int rlr_lexical_add_chunk(rlr_lexical *lx, uint64_t chunk_key, const char *text_utf8, size_t len);
int rlr_lexical_remove_chunk(rlr_lexical *lx, uint64_t chunk_key);
int rlr_lexical_contains(const rlr_lexical *lx, uint64_t chunk_key, int *out);
int rlr_lexical_stats(const rlr_lexical *lx, uint64_t *total_docs, uint64_t *total_length, uint64_t *n_terms);
int rlr_lexical_score(const rlr_lexical *lx, const char *query_utf8, size_t len, uint32_t limit,
                      uint64_t *out_keys, float *out_scores, uint32_t cap, uint32_t *out_n);
int rlr_tokenize(const char *text_utf8, size_t len, char *out, size_t out_cap, size_t *out_len,
                 uint32_t *out_tokens);
/* test hook: is_alphanumeric and the full lowercase mapping (3 code points, 0 padded) of code points [0, n_cp) */
int rlr_hostmirror_unicode_dump(uint8_t *out_alnum, uint32_t *out_lower, uint32_t n_cp);

#ifdef __cplusplus
}
#endif
#endif /* RLR_HOSTMIRROR_H */
