//! Raw bindings to `include/rlr_b200.h`.  NOT COMPILED HERE (no Rust toolchain in the image).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const RLR_OK: c_int = 0;
pub const RLR_ERR_INVALID_ARG: c_int = 1;
pub const RLR_ERR_NO_DEVICE: c_int = 2;
pub const RLR_ERR_CUDA: c_int = 3;
pub const RLR_ERR_OOM: c_int = 4;
pub const RLR_ERR_DIM_MISMATCH: c_int = 5;
pub const RLR_ERR_UNSUPPORTED: c_int = 6;
pub const RLR_ERR_NONFINITE: c_int = 7;
pub const RLR_MAX_M: u32 = 1024;
pub const RLR_QUERY_PRENORMALIZED: u32 = 0x1;
pub const RLR_WANT_TIMINGS: u32 = 0x2;
pub const RLR_SEARCH_F16: u32 = 0x4;
pub const RLR_STORE_KEEP_F16: u32 = 0x1;
pub const RLR_STORE_F16_ONLY: u32 = 0x4;
pub const RLR_STORE_CHECK_FINITE: u32 = 0x2;
pub const RLR_STORE_NORMALIZE_ON_UPLOAD: u32 = 0x8;
pub const RLR_BATCH_EXACT_RESCORE: u32 = 0x8;
pub const RLR_BATCH_F16: u32 = 0x10;
pub const RLR_BATCH_BF16: u32 = 0x20;
pub const RLR_BATCH_TF32: u32 = 0x40;
pub const RLR_STORE_KEEP_BF16: u32 = 0x10;
pub const RLR_STORE_NO_LATENCY_PATH: u32 = 0x20;
pub const RLR_MAX_TOP_K: u32 = 100;
pub const RLR_MAX_DIM: u32 = 4096;
pub const RLR_IPC_HANDLE_BYTES: usize = 64;
pub const RLR_SYNTH_IID: c_int = 0;
pub const RLR_SYNTH_CLUSTERED: c_int = 1;
pub const RLR_ABI_VERSION: c_int = 1;
pub const RLR_MAX_SHARDS: usize = 16;
pub const RLR_MAX_MULTI: u32 = 3;

#[repr(C)] pub struct rlr_store { _p: [u8; 0] }
#[repr(C)] pub struct rlr_ctx { _p: [u8; 0] }
#[repr(C)] pub struct rlr_peer_set { _p: [u8; 0] }
#[repr(C)] pub struct rlr_mailbox { _p: [u8; 0] }
#[repr(C)] pub struct rlr_cluster { _p: [u8; 0] }
#[repr(C)] pub struct rlr_bm25 { _p: [u8; 0] }
#[repr(C)] pub struct rlr_cluster_bm25 { _p: [u8; 0] }

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct rlr_query_weights { pub embedding: f32, pub lexical: f32, pub reranker: f32, pub initial: f32, pub has: u32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct rlr_resolved_weights { pub embedding: f32, pub lexical: f32, pub reranker: f32, pub initial: f32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct rlr_store_info { pub n_rows: u64, pub row_base: u64, pub dim: u32, pub pitch: u32, pub device: i32, pub flags: u32, pub bytes_device: u64 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct rlr_device_info { pub device: i32, pub sm_count: i32, pub cc_major: i32, pub cc_minor: i32, pub total_mem: u64, pub name: [c_char; 128] }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct rlr_timings { pub scan_ms: f32, pub merge_ms: f32, pub mmr_ms: f32, pub total_ms: f32, pub launches: u32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct rlr_cluster_info { pub n_rows: u64, pub dim: u32, pub pitch: u32, pub flags: u32, pub n_shards: u32, pub device: [i32; RLR_MAX_SHARDS], pub row_base: [u64; RLR_MAX_SHARDS], pub shard_rows: [u64; RLR_MAX_SHARDS] }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct rlr_cand { pub key: u64, pub emb: f32, pub lex: f32 }

extern "C" {
    pub fn rlr_abi_version() -> c_int;
    pub fn rlr_last_error() -> *const c_char;
    pub fn rlr_device_count(out_count: *mut c_int) -> c_int;
    pub fn rlr_device_query(device: c_int, out: *mut rlr_device_info) -> c_int;
    pub fn rlr_normalize(v: *mut f32, n: usize) -> c_int;
    pub fn rlr_resolve_weights(overrides: *const rlr_query_weights, out: *mut rlr_resolved_weights) -> c_int;
    pub fn rlr_store_create(device: c_int, dim: u32, n_rows: u64, rows: *const f32, host_pitch: u64, row_base: u64, flags: u32, out: *mut *mut rlr_store) -> c_int;
    pub fn rlr_store_destroy(s: *mut rlr_store) -> c_int;
    pub fn rlr_store_info_get(s: *const rlr_store, out: *mut rlr_store_info) -> c_int;
    pub fn rlr_store_upload(s: *mut rlr_store, row0: u64, n: u64, rows: *const f32, host_pitch: u64) -> c_int;
    pub fn rlr_store_reserve(s: *mut rlr_store, capacity_rows: u64) -> c_int;
    pub fn rlr_store_append(s: *mut rlr_store, n: u64, rows: *const f32, host_pitch: u64, out_first_row: *mut u64) -> c_int;
    pub fn rlr_store_remove_rows(s: *mut rlr_store, rows: *const u32, n: u64, out_moved_from: *mut u32, out_moved_to: *mut u32, out_n_moved: *mut u64) -> c_int;
    pub fn rlr_store_read_rows(s: *const rlr_store, rows: *const u32, n: u64, out: *mut f32) -> c_int;
    pub fn rlr_store_fill_synthetic(s: *mut rlr_store, kind: c_int, seed: u64, centroid_seed: u64, n_clusters: u32, sigma: f32) -> c_int;
    pub fn rlr_search_topm(s: *mut rlr_store, query: *const f32, dim: u32, flags: u32, w: *const rlr_resolved_weights, lex_rows: *const u32, lex_scores: *const f32, n_lex: u32, m: u32, out_rows: *mut u32, out_combined: *mut f32, out_emb: *mut f32, out_lex: *mut f32, out_n: *mut u32) -> c_int;
    pub fn rlr_mmr(s: *mut rlr_store, cand_rows: *const u32, relevance: *const f32, p: u32, top_k: u32, lambda: f32, flags: u32, out_sel_pos: *mut u32, out_n: *mut u32) -> c_int;
    pub fn rlr_search_mmr(s: *mut rlr_store, query: *const f32, dim: u32, flags: u32, top_k: u32, diversity_factor: f32, w: *const rlr_resolved_weights, lex_rows: *const u32, lex_scores: *const f32, n_lex: u32, out_rows: *mut u32, out_score: *mut f32, out_emb: *mut f32, out_lex: *mut f32, out_n: *mut u32) -> c_int;
    pub fn rlr_search_mmr_multi(s: *mut rlr_store, queries: *const f32, nq: u32, dim: u32, flags: u32, top_k: u32, diversity_factor: f32, w: *const rlr_resolved_weights, lex_rows: *const *const u32, lex_scores: *const *const f32, n_lex: *const u32, out_rows: *mut u32, out_score: *mut f32, out_emb: *mut f32, out_lex: *mut f32, out_n: *mut u32) -> c_int;
    pub fn rlr_search_mmr_multi_async(ctxs: *const *mut rlr_ctx, nq: u32, d_queries: *const *const c_void, top_k: u32, diversity_factor: f32, w_embed: f32, w_lex: f32, d_results: *const *mut c_void, d_result_ns: *const *mut c_void, stream: *mut c_void) -> c_int;
    pub fn rlr_embedding_candidates(s: *mut rlr_store, query: *const f32, dim: u32, flags: u32, count: u32, out_rows: *mut u32, out_score: *mut f32, out_n: *mut u32) -> c_int;
    pub fn rlr_search_batch(s: *mut rlr_store, queries: *const f32, n_queries: u32, dim: u32, flags: u32, m: u32, out_rows: *mut u32, out_scores: *mut f32, out_n: *mut u32) -> c_int;
    pub fn rlr_search_batch_device(s: *mut rlr_store, queries: *const f32, n_queries: u32, dim: u32, flags: u32, m: u32, d_keys: *mut c_void, d_cnt: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn rlr_batch_merge_async(s: *mut rlr_store, d_lists: *const c_void, n_lists: u32, n_queries: u32, m: u32, d_out_keys: *mut c_void, d_out_cnt: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn rlr_bm25_create(s: *mut rlr_store, out: *mut *mut rlr_bm25) -> c_int;
    pub fn rlr_bm25_destroy(ix: *mut rlr_bm25) -> c_int;
    pub fn rlr_bm25_set_doc(ix: *mut rlr_bm25, row: u32, term_ids: *const u32, term_freqs: *const u32, n_terms: u32) -> c_int;
    pub fn rlr_bm25_set_docs(ix: *mut rlr_bm25, row0: u32, n_docs: u32, offsets: *const u64, term_ids: *const u32, term_freqs: *const u32) -> c_int;
    pub fn rlr_bm25_remove_doc(ix: *mut rlr_bm25, row: u32) -> c_int;
    pub fn rlr_bm25_move_doc(ix: *mut rlr_bm25, from_row: u32, to_row: u32) -> c_int;
    pub fn rlr_bm25_stats(ix: *const rlr_bm25, total_docs: *mut u64, total_length: *mut u64, n_terms: *mut u64) -> c_int;
    pub fn rlr_bm25_score(ix: *mut rlr_bm25, query_terms: *const u32, n_terms: u32, limit: u32, out_rows: *mut u32, out_scores: *mut f32, cap: u32, out_n: *mut u32) -> c_int;
    pub fn rlr_search_text_topm(s: *mut rlr_store, ix: *mut rlr_bm25, query: *const f32, dim: u32, flags: u32, w: *const rlr_resolved_weights, query_terms: *const u32, n_terms: u32, m: u32, out_rows: *mut u32, out_combined: *mut f32, out_emb: *mut f32, out_lex: *mut f32, out_n: *mut u32) -> c_int;
    pub fn rlr_search_text_mmr(s: *mut rlr_store, ix: *mut rlr_bm25, query: *const f32, dim: u32, flags: u32, top_k: u32, diversity_factor: f32, w: *const rlr_resolved_weights, query_terms: *const u32, n_terms: u32, out_rows: *mut u32, out_score: *mut f32, out_emb: *mut f32, out_lex: *mut f32, out_n: *mut u32) -> c_int;
    pub fn rlr_cluster_bm25_create(c: *mut rlr_cluster, out: *mut *mut rlr_cluster_bm25) -> c_int;
    pub fn rlr_cluster_bm25_destroy(ix: *mut rlr_cluster_bm25) -> c_int;
    pub fn rlr_cluster_bm25_set_doc(ix: *mut rlr_cluster_bm25, row: u32, term_ids: *const u32, term_freqs: *const u32, n_terms: u32) -> c_int;
    pub fn rlr_cluster_bm25_set_docs(ix: *mut rlr_cluster_bm25, row0: u32, n_docs: u32, offsets: *const u64, term_ids: *const u32, term_freqs: *const u32) -> c_int;
    pub fn rlr_cluster_bm25_remove_doc(ix: *mut rlr_cluster_bm25, row: u32) -> c_int;
    pub fn rlr_cluster_bm25_stats(ix: *const rlr_cluster_bm25, total_docs: *mut u64, total_length: *mut u64, n_terms: *mut u64) -> c_int;
    pub fn rlr_cluster_bm25_score(ix: *mut rlr_cluster_bm25, query_terms: *const u32, n_terms: u32, limit: u32, out_rows: *mut u32, out_scores: *mut f32, cap: u32, out_n: *mut u32) -> c_int;
    pub fn rlr_cluster_search_text_topm(c: *mut rlr_cluster, ix: *mut rlr_cluster_bm25, query: *const f32, dim: u32, flags: u32, w: *const rlr_resolved_weights, query_terms: *const u32, n_terms: u32, m: u32, out_rows: *mut u32, out_combined: *mut f32, out_emb: *mut f32, out_lex: *mut f32, out_n: *mut u32) -> c_int;
    pub fn rlr_cluster_search_text_mmr(c: *mut rlr_cluster, ix: *mut rlr_cluster_bm25, query: *const f32, dim: u32, flags: u32, top_k: u32, diversity_factor: f32, w: *const rlr_resolved_weights, query_terms: *const u32, n_terms: u32, out_rows: *mut u32, out_score: *mut f32, out_emb: *mut f32, out_lex: *mut f32, out_n: *mut u32) -> c_int;
    pub fn rlr_last_timings(out: *mut rlr_timings) -> c_int;
    pub fn rlr_cluster_create(devices: *const c_int, n_devices: u32, dim: u32, n_rows: u64, rows: *const f32, host_pitch: u64, flags: u32, shard_rows: *const u64, out: *mut *mut rlr_cluster) -> c_int;
    pub fn rlr_cluster_destroy(c: *mut rlr_cluster) -> c_int;
    pub fn rlr_cluster_info_get(c: *const rlr_cluster, out: *mut rlr_cluster_info) -> c_int;
    pub fn rlr_cluster_upload(c: *mut rlr_cluster, row0: u64, n: u64, rows: *const f32, host_pitch: u64) -> c_int;
    pub fn rlr_cluster_read_rows(c: *const rlr_cluster, rows: *const u32, n: u64, out: *mut f32) -> c_int;
    pub fn rlr_cluster_fill_synthetic(c: *mut rlr_cluster, kind: c_int, seed: u64, centroid_seed: u64, n_clusters: u32, sigma: f32) -> c_int;
    pub fn rlr_cluster_search_topm(c: *mut rlr_cluster, query: *const f32, dim: u32, flags: u32, w: *const rlr_resolved_weights, lex_rows: *const u32, lex_scores: *const f32, n_lex: u32, m: u32, out_rows: *mut u32, out_combined: *mut f32, out_emb: *mut f32, out_lex: *mut f32, out_n: *mut u32) -> c_int;
    pub fn rlr_cluster_mmr(c: *mut rlr_cluster, cand_rows: *const u32, relevance: *const f32, p: u32, top_k: u32, lambda: f32, flags: u32, out_sel_pos: *mut u32, out_n: *mut u32) -> c_int;
    pub fn rlr_cluster_search_mmr(c: *mut rlr_cluster, query: *const f32, dim: u32, flags: u32, top_k: u32, diversity_factor: f32, w: *const rlr_resolved_weights, lex_rows: *const u32, lex_scores: *const f32, n_lex: u32, out_rows: *mut u32, out_score: *mut f32, out_emb: *mut f32, out_lex: *mut f32, out_n: *mut u32) -> c_int;
    pub fn rlr_cluster_search_mmr_multi(c: *mut rlr_cluster, queries: *const f32, nq: u32, dim: u32, flags: u32, top_k: u32, diversity_factor: f32, w: *const rlr_resolved_weights, lex_rows: *const *const u32, lex_scores: *const *const f32, n_lex: *const u32, out_rows: *mut u32, out_score: *mut f32, out_emb: *mut f32, out_lex: *mut f32, out_n: *mut u32) -> c_int;
    pub fn rlr_cluster_search_batch(c: *mut rlr_cluster, queries: *const f32, n_queries: u32, dim: u32, flags: u32, m: u32, out_rows: *mut u32, out_scores: *mut f32, out_n: *mut u32) -> c_int;
    pub fn rlr_cluster_embedding_candidates(c: *mut rlr_cluster, query: *const f32, dim: u32, flags: u32, count: u32, out_rows: *mut u32, out_score: *mut f32, out_n: *mut u32) -> c_int;
    pub fn rlr_cluster_last_scan_ms(out_ms: *mut f32, cap: u32, out_n: *mut u32) -> c_int;
    pub fn rlr_cluster_launch_count(c: *const rlr_cluster, out: *mut u64) -> c_int;
    pub fn rlr_ctx_create(s: *mut rlr_store, out: *mut *mut rlr_ctx) -> c_int;
    pub fn rlr_ctx_destroy(c: *mut rlr_ctx) -> c_int;
    pub fn rlr_topm_async(c: *mut rlr_ctx, d_query: *const c_void, w_embed: f32, w_lex: f32, d_lex_rows: *const c_void, d_lex_norm: *const c_void, n_lex: u32, m: u32, d_out: *mut c_void, d_out_n: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn rlr_merge_async(c: *mut rlr_ctx, d_lists: *const c_void, n_lists: u32, m: u32, d_out: *mut c_void, d_out_n: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn rlr_gather_async(c: *mut rlr_ctx, d_cands: *const c_void, d_n: *const c_void, m: u32, d_out: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn rlr_mmr_async(c: *mut rlr_ctx, d_emb: *const c_void, pitch: u32, dim: u32, d_cands: *const c_void, d_n: *const c_void, p_cap: u32, top_k: u32, lambda: f32, d_sel_pos: *mut c_void, d_sel_n: *mut c_void, d_result: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn rlr_mmr_store_async(c: *mut rlr_ctx, d_cands: *const c_void, d_n: *const c_void, p_cap: u32, top_k: u32, lambda: f32, d_sel_pos: *mut c_void, d_sel_n: *mut c_void, d_result: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn rlr_store_ipc_export(s: *const rlr_store, search_flags: u32, handle_out: *mut c_void) -> c_int;
    pub fn rlr_peer_set_open(local: *mut rlr_store, my_index: u32, n_shards: u32, handles: *const c_void, row_base: *const u64, n_rows: *const u64, search_flags: u32, out: *mut *mut rlr_peer_set) -> c_int;
    pub fn rlr_peer_set_close(p: *mut rlr_peer_set) -> c_int;
    pub fn rlr_mmr_peers_async(c: *mut rlr_ctx, p: *mut rlr_peer_set, d_cands: *const c_void, d_n: *const c_void, p_cap: u32, top_k: u32, lambda: f32, d_sel_pos: *mut c_void, d_sel_n: *mut c_void, d_result: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn rlr_mailbox_create(device: c_int, n_ranks: u32, m_cap: u32, ring: u32, out: *mut *mut rlr_mailbox) -> c_int;
    pub fn rlr_mailbox_ipc_export(mb: *const rlr_mailbox, handle_out: *mut c_void) -> c_int;
    pub fn rlr_mailbox_open(device: c_int, handle: *const c_void, n_ranks: u32, m_cap: u32, ring: u32, out: *mut *mut rlr_mailbox) -> c_int;
    pub fn rlr_mailbox_close(mb: *mut rlr_mailbox) -> c_int;
    pub fn rlr_mailbox_status(mb: *mut rlr_mailbox, out: *mut u32) -> c_int;
    pub fn rlr_topm_post_async(c: *mut rlr_ctx, mb: *mut rlr_mailbox, my_rank: u32, seq: u64, d_query: *const c_void, w_embed: f32, w_lex: f32, d_lex_rows: *const c_void, d_lex_norm: *const c_void, n_lex: u32, m: u32, stream: *mut c_void) -> c_int;
    pub fn rlr_mailbox_merge_async(c: *mut rlr_ctx, mb: *mut rlr_mailbox, seq: u64, m: u32, d_out: *mut c_void, d_out_n: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn rlr_search_mmr_async(c: *mut rlr_ctx, d_query: *const c_void, top_k: u32, diversity_factor: f32, w_embed: f32, w_lex: f32, d_result: *mut c_void, d_result_n: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn rlr_ctx_set_flags(c: *mut rlr_ctx, search_flags: u32) -> c_int;
    pub fn rlr_ctx_launch_count(c: *const rlr_ctx, out: *mut u64) -> c_int;
    pub fn rlr_time_scan(c: *mut rlr_ctx, d_query: *const c_void, m: u32, iters: u32, stream: *mut c_void, out_ms_per_launch: *mut f32) -> c_int;
}
