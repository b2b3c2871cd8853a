// kernels.cuh -- host-side launch interface of the sm_100a kernels (internal).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rlr_b200.h"

namespace rlr {

constexpr int kScanRows = 128;        // rows per tile == consumer threads per CTA
constexpr int kScanChunks = 2;        // 128-byte column chunks per pipeline stage
constexpr int kScanThreads = kScanRows + 64;   // 4 consumer warps (per query group) + TMA producer warp + threshold warp
constexpr int kTopBuf = 2048;         // per-CTA candidate buffer (entries)
constexpr int kChunkFloats = 32;      // 128 B : the TMA SWIZZLE_128B span
constexpr int kQueryCap = RLR_MAX_DIM + 128; // floats, zero padded (a stage spans <= 128 elements)

// Optional delivery of the scan's merged list into a mailbox slot that may live in another
// GPU's HBM (peer mapping over NVLink): the last CTA waits until the slot is free, writes the
// list through `d_out`/`d_out_n` (peer pointers), then publishes `seq` with a system-scope
// release store.  See mailbox_merge_launch.
struct ScanPost {
    unsigned long long *flag;            // null: no post (plain local scan)
    const unsigned long long *consumed;  // this SLOT's word: free once *consumed + ring >= seq (the root has merged
                                         // the query that used the slot `ring` sequence numbers ago)
    unsigned long long seq;
    uint32_t ring;
    uint32_t *status;                    // local word, set non-zero if the wait timed out
};

// Latency path for small stores (BASELINE configs[0]: 10k chunks, top_k = 5 -- the reference's real operating
// point): the WHOLE request is one kernel launch with no copy engine involved.  The normalised query (and the few
// lexical pairs) travel in the kernel's parameter block, the last CTA runs merge + pairwise + greedy MMR itself when
// the pool is small, and the result is written straight into mapped pinned host memory followed by a system-scope
// flag that the host polls -- no H2D copy, no D2H copy, no cudaStreamSynchronize.
constexpr uint32_t kLatQFloats = 1024;     // query floats a parameter block carries (pitch <= 1024)
constexpr uint32_t kLatLex = 192;          // lexical pairs a parameter block carries (5 * pool for pool <= 32, rounded up)
constexpr uint32_t kLatFusePool = 32;      // pools up to this size are diversified by the scan's last CTA itself
struct LatParams {
    alignas(16) float q[kLatQFloats];      // normalised query, zero beyond dim (read 16 bytes at a time)
    uint32_t lex_rows[kLatLex];            // sorted local rows
    float lex_norm[kLatLex];
    const uint32_t *d_lex_rows;            // non-null: the lexical pairs are DEVICE arrays instead (written by the device
    const float *d_lex_norm;               // BM25 stage right before this launch); the two arrays above are ignored
    uint32_t keep_l2;                      // 1: the whole store fits L2 -- load it evict_last so that it stays resident between queries
    uint32_t mode;                         // 0: off (plain scan) | 1: deliver the merged top-m | 2: + fused MMR of the pool
    uint32_t top_k;
    float lambda;
    uint32_t pitch;                        // floats per stored f32 row
    const float *g_rows;                   // the store's f32 rows (pool rows are staged from here for the fused MMR)
    uint32_t *d_sel_pos;                   // device scratch, >= kLatFusePool words
    rlr_cand *result;                      // mapped pinned host memory (mode != 0): records ...
    uint32_t *result_n;                    // ... count ...
    unsigned long long *flag;              // ... and the completion word, which receives `seq` last (release, system scope)
    unsigned long long seq;
};

// Query groups (throughput mode): one launch answers up to kMaxQueryGroups queries in a single pass over the rows.
constexpr int kMaxQueryGroups = 3;
struct ScanGroupIO {
    const float *query;        // kQueryCap floats on the device, zero beyond dim
    const uint32_t *lex_rows;  // this query's lexical pairs: sorted local rows (may be null) ...
    const float *lex_norm;     // ... their lexical_score (already / max_lexical)
    uint32_t n_lex;
    rlr_cand *out;             // best m records over all CTAs (null: skip the in-kernel merge)
    uint32_t *out_n;
    ScanPost post;             // zeroed: off
};
struct ScanGroups { ScanGroupIO g[kMaxQueryGroups]; };

struct ScanArgs {
    const CUtensorMap *tmap;   // host pointer; copied into the kernel's param space (f32 or f16 map)
    int half;                  // 1: the map describes the binary16 copy of the store
    const float *d_query;      // kQueryCap floats, zero beyond dim
    uint32_t n_rows;
    uint32_t row_base;
    uint32_t pitch;            // elements per stored row: multiple of 32 (f32) / 64 (f16)
    float w_embed, w_lex;
    const uint32_t *d_lex_rows; // sorted ascending, local rows (may be null)
    const float *d_lex_norm;    // lexical_score per entry (already / max_lexical)
    uint32_t n_lex;
    uint32_t m;                // 1..RLR_MAX_M
    rlr_cand *d_lists;         // n_groups x grid x m records
    uint32_t *d_counts;        // n_groups x grid
    uint32_t *d_ticket;        // 8 words, zero-initialised: [g] finish ticket of query group g, [4] dynamic tile
                               // counter; the last CTA to finish (per group) merges and resets them
    uint32_t *d_pub;           // n_groups x grid words, zero-initialised: per-CTA published r-th best score
    rlr_cand *d_out;           // best m records over all CTAs (null: skip the in-kernel merge)
    uint32_t *d_out_n;
    int grid;                  // CTAs to launch (<= SM count)
    int smem_bytes;            // dynamic shared memory to request
    int n_stages;
    uint32_t buf_cap;          // 0: derive from m
    unsigned long long *d_trace; // dev-only: phase timestamps (5*grid + 16 words) or null
    ScanPost post;             // zeroed: off
    uint32_t n_groups;         // 0 / 1: one query (d_query, d_out, d_out_n, post above); 2..kMaxQueryGroups: `groups`
    ScanGroups groups;         // per-group query / outputs / post when n_groups > 1
    uint32_t rows_per_tile;    // 0 / kScanRows: 128-row tiles.  Small stores use fewer rows per tile (a multiple of 8, with
                               // a tensor map whose box has that many rows) so that the rows spread over every SM
    const LatParams *lat;      // host pointer or null: latency path (f32 stores only); d_query / d_lex_* are then ignored
};

// Pick grid / stages / smem for a store on a device.
void scan_plan(int sm_count, int max_smem_optin, uint32_t n_rows, uint32_t pitch, int half, ScanArgs *a, uint32_t n_groups = 1);
// rows per tile for a store of n_rows on sm_count SMs: kScanRows for stores that fill every SM with 128-row tiles,
// else the least multiple of 8 that gives every SM at most one tile
uint32_t scan_rows_per_tile(int sm_count, uint64_t n_rows);
// plan for the latency path: every SM, `rows_per_tile`-row tiles, room for the lexical pairs in shared memory
void scan_plan_small(int sm_count, int max_smem_optin, uint32_t n_rows, uint32_t pitch, uint32_t rows_per_tile, ScanArgs *a);
cudaError_t scan_configure(); // one-time cudaFuncSetAttribute
cudaError_t scan_launch(const ScanArgs &a, cudaStream_t stream);

// Merge n_lists lists of m records (zero-key padded) into the best m.  Uses d_tmp
// (capacity >= merge_tmp_records(n_lists, m) records) for intermediate levels.
size_t merge_tmp_records(uint32_t n_lists, uint32_t m);
cudaError_t merge_launch(const rlr_cand *d_lists, uint32_t n_lists, uint32_t m, rlr_cand *d_tmp,
                         rlr_cand *d_out, uint32_t *d_out_n, cudaStream_t stream, uint32_t *launches);

// Root side of the fused exchange: wait (acquire, system scope) until every rank's list for
// `seq` has landed in the mailbox slot, merge the n_lists lists (each `m` of `list_stride`
// records) by rank counting, then release the slot (`*consumed = seq`).
cudaError_t mailbox_merge_launch(const rlr_cand *d_slot, uint32_t list_stride, const unsigned long long *d_flags,
                                 unsigned long long seq, unsigned long long *d_consumed, uint32_t n_lists, uint32_t m,
                                 rlr_cand *d_out, uint32_t *d_out_n, uint32_t *d_status, cudaStream_t stream);
constexpr unsigned long long kMailboxTimeoutNs = 4000000000ull;   // a dead peer must not hang the GPU

// Batched path, multi-GPU: per query, merge n_lists rank-ordered key lists (d_lists is
// [n_lists][n_queries][m] u64 keys, zero padded) into the best m (rank counting, one CTA per query).
cudaError_t batch_merge_launch(const unsigned long long *d_lists, uint32_t n_lists, uint32_t n_queries, uint32_t m,
                               unsigned long long *d_out, uint32_t *d_out_cnt, cudaStream_t stream);

// MMR: pairwise similarities (upper triangle) + greedy selection.
//   d_emb/pitch : matrix the candidate embeddings live in
//   d_cands     : p_cap candidate records (rank order); row index of candidate i is
//                 (key_row(cand.key) - row_base) when use_rows != 0, else i
//   d_rel       : optional explicit relevance (else decoded from the keys)
// Row shards of one corpus living on several GPUs of an NVSwitch box, reachable from this GPU
// through peer mappings (CUDA IPC): the MMR pairwise kernel loads candidate rows straight
// from the owning GPU's HBM over NVLink.
constexpr int kMaxPeers = 16;
struct PeerTable {
    const void *base[kMaxPeers];
    uint32_t row_base[kMaxPeers];
    uint32_t n_rows[kMaxPeers];
    uint32_t n;                // 0: no peers, rows come from d_emb
};

struct MmrArgs {
    const void *d_emb;         // f32 rows, or binary16 rows when half != 0
    int half;
    uint32_t pitch, dim;       // pitch in elements
    const rlr_cand *d_cands;
    const uint32_t *d_n;       // number of valid candidates (device)
    const uint32_t *d_rows;    // optional explicit local rows (overrides keys)
    const float *d_rel;        // optional explicit relevance
    uint32_t row_base;
    int use_rows;
    uint32_t p_cap;            // launch bound for p
    uint32_t top_k;
    float lambda;
    float *d_tri;              // p_cap*(p_cap-1)/2 floats
    uint32_t *d_sel_pos;       // top_k (>=1)
    uint32_t *d_sel_n;
    rlr_cand *d_result;        // optional: selected records in selection order
    int max_smem_optin;
    const PeerTable *peers;    // optional (host pointer): candidate rows are global and live on these shards
    unsigned long long *done_flag;   // optional: (mapped pinned) word that receives done_seq, system scope, after the
    unsigned long long done_seq;     // result and count have been written -- the host polls it instead of synchronising
    void *d_gather;            // with peers: p_cap x pitch elements of scratch; the rows are gathered into it once (NVLink)
                               // and the pairwise kernel reads local memory.  null: the pairwise kernel loads from the peers itself
};
cudaError_t mmr_configure();
cudaError_t mmr_launch(const MmrArgs &a, cudaStream_t stream, uint32_t *launches);

// gather rows owned by this shard into a dense p x pitch matrix (zeros otherwise)
cudaError_t gather_launch(const void *d_store, int half, uint32_t pitch, uint32_t n_rows, uint32_t row_base,
                          const rlr_cand *d_cands, const uint32_t *d_n, uint32_t p_cap, float *d_out,
                          uint32_t out_pitch, cudaStream_t stream);
cudaError_t gather_rows_launch(const void *d_store, int half, uint32_t pitch, const uint32_t *d_rows, uint32_t n,
                               float *d_out, uint32_t out_pitch, cudaStream_t stream);

// synthetic rows (f32 and/or binary16 output; either pointer may be null) + finite check + f32->f16
cudaError_t synth_launch(float *d_rows, uint32_t pitch, void *d_rows16, uint32_t pitch16, uint32_t dim, uint64_t row_base, uint32_t n_rows,
                         int kind, uint64_t seed, uint64_t centroid_seed, uint32_t n_clusters, float sigma,
                         cudaStream_t stream);
// in-place reference normalize of n_rows stored f32 rows (sequential arithmetic per row)
cudaError_t normalize_rows_launch(float *d_rows, uint32_t pitch, uint32_t dim, uint64_t n_rows, cudaStream_t stream);
cudaError_t finite_check_launch(const float *d_rows, uint64_t n_floats, uint32_t *d_flag, cudaStream_t stream);
// rows[to[i]] = rows[from[i]] (disjoint sources and destinations), row_bytes a multiple of 16
cudaError_t move_rows_launch(void *d_rows, uint32_t row_bytes, const uint32_t *d_from, const uint32_t *d_to, uint32_t n,
                             cudaStream_t stream);
cudaError_t to_half_launch(const float *d_src, uint32_t src_pitch, void *d_dst, uint32_t dst_pitch, uint32_t dim,
                           uint64_t n_rows, cudaStream_t stream);

// batched-query path (tcgen05 GEMM + per-query threshold filter), batch_gemm.cu
size_t batch_smem_bytes(uint32_t nq_pad);
cudaError_t batch_configure(int smem_optin);
cudaError_t batch_init_launch(float *tau, uint32_t *state_cnt, uint32_t *app_cnt, uint32_t nq, uint32_t nq_pad,
                              uint32_t *overflow, cudaStream_t st);
// operand precision of the batched contraction: what a 128-byte k-chunk of the operand tiles holds
constexpr int kPrecF16 = 0, kPrecBF16 = 1, kPrecTF32 = 2;
cudaError_t batch_queries_to_operand_launch(const float *d_q, uint32_t dim, void *d_qop, uint32_t pitch_elems, int prec,
                                            uint32_t nq, uint32_t nq_pad, uint32_t *d_nonfinite /* nullable */, cudaStream_t st);
cudaError_t to_bf16_launch(const float *d_src, uint32_t src_pitch, void *d_dst, uint32_t dst_pitch, uint32_t dim,
                           uint64_t n_rows, cudaStream_t stream);
cudaError_t batch_gemm_launch(const CUtensorMap *tmapA, const CUtensorMap *tmapQ, int grid, uint32_t n_rows,
                              uint32_t row_base, uint32_t tile0, uint32_t tile1, uint32_t nq_pad, uint32_t pitch_elems, int prec,
                              const float *tau, unsigned long long *app_keys, uint32_t *app_cnt, uint32_t cap,
                              uint32_t *overflow, cudaStream_t st);
cudaError_t batch_gemm2_launch(const CUtensorMap *tmapA, const CUtensorMap *tmapQ128, int sm_count, uint32_t n_rows,
                               uint32_t row_base, uint32_t tile0, uint32_t tile1, uint32_t nq_pad, uint32_t pitch_elems, int prec,
                               const float *tau, unsigned long long *app_keys, uint32_t *app_cnt, uint32_t cap,
                               uint32_t *overflow, int dense, cudaStream_t st);
cudaError_t batch_set_cnt_launch(uint32_t *app_cnt, uint32_t nq, uint32_t value, cudaStream_t st);
cudaError_t batch_prune_launch(unsigned long long *state_keys, uint32_t *state_cnt, uint32_t m, unsigned long long *app_keys,
                               uint32_t *app_cnt, uint32_t cap, float *tau, uint32_t nq, cudaStream_t st);
cudaError_t batch_rescore_launch(const void *d_rows, int half, uint32_t pitch, uint32_t dim, uint32_t row_base,
                                 const float *d_q32, unsigned long long *state_keys, const uint32_t *state_cnt, uint32_t m,
                                 uint32_t nq, cudaStream_t st);

// copy a selected subset of records into SoA output arrays (host-facing results)
cudaError_t unpack_launch(const rlr_cand *d_cands, const uint32_t *d_n, uint32_t cap, uint32_t *d_rows,
                          float *d_score, float *d_emb, float *d_lex, cudaStream_t stream);

} // namespace rlr
