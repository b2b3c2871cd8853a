// api.cu -- the C ABI of include/rlr_b200.h: store management, host<->device plumbing
// and launch sequencing around the sm_100a kernels.  No compute happens on the host
// except the O(dim) query normalisation and O(n_lex) lexical normalisation that the
// reference also does once per query (/root/reference/src/rag_engine.rs:494,511-530).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "api_internal.hpp"

namespace rlr_api {

thread_local std::string g_err;
thread_local rlr_timings g_timings = {0, 0, 0, 0, 0};

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

} // namespace rlr_api
using namespace rlr_api;

namespace {

struct DeviceState {
    bool checked = false;
    bool ok = false;
    int sm_count = 0;
    int smem_optin = 0;
    std::string why;
};
std::mutex g_dev_mu;
DeviceState g_dev[64];

} // namespace

// A device is usable iff it exists and is compute capability 10.x (the fatbin holds
// sm_100a SASS only).  There is no fallback.
int rlr_api::ensure_device(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(RLR_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= count || device >= 64)
        return fail(RLR_ERR_INVALID_ARG, "device %d out of range (have %d)", device, count);
    std::lock_guard<std::mutex> lk(g_dev_mu);
    DeviceState &d = g_dev[device];
    if (!d.checked) {
        d.checked = true;
        cudaDeviceProp prop;
        e = cudaGetDeviceProperties(&prop, device);
        if (e != cudaSuccess) {
            d.why = cudaGetErrorString(e);
        } else if (prop.major != 10) {
            char b[160];
            snprintf(b, sizeof b, "device %d (%s) is sm_%d%d; kernels are built for sm_100a only", device, prop.name,
                     prop.major, prop.minor);
            d.why = b;
        } else {
            d.sm_count = prop.multiProcessorCount;
            e = cudaSetDevice(device);
            if (e == cudaSuccess) e = cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
            if (e == cudaSuccess) e = rlr::scan_configure();
            if (e == cudaSuccess) e = rlr::mmr_configure();
            if (e != cudaSuccess) d.why = cudaGetErrorString(e);
            else d.ok = true;
        }
    }
    if (!d.ok) return fail(RLR_ERR_NO_DEVICE, "%s", d.why.c_str());
    CU_TRY(cudaSetDevice(device));
    return RLR_OK;
}

int rlr_api::device_sm_count(int device)
{
    std::lock_guard<std::mutex> lk(g_dev_mu);
    return (device >= 0 && device < 64) ? g_dev[device].sm_count : 0;
}

namespace {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
        else
            cudaGetLastError();
    });
    return fn;
}

} // namespace

void rlr_api::ctx_free(rlr_ctx *c)
{
    if (!c) return;
    cudaFree(c->d_query); cudaFree(c->d_lex_rows); cudaFree(c->d_lex_norm); cudaFree(c->d_lists);
    cudaFree(c->d_counts); cudaFree(c->d_ticket); cudaFree(c->d_pub); cudaFree(c->d_tmp); cudaFree(c->d_pool_blk); cudaFree(c->d_tri); cudaFree(c->d_gather);
    cudaFree(c->d_sel_pos); cudaFree(c->d_result_blk); cudaFree(c->d_rows_in);
    cudaFree(c->d_rel_in); cudaFree(c->d_p_in);
    cudaFreeHost(c->h_query); cudaFreeHost(c->h_lex_rows); cudaFreeHost(c->h_lex_norm);
    cudaFreeHost(c->h_result_blk); cudaFreeHost(c->h_u32); cudaFreeHost(c->h_rel);
    cudaFree(c->batch_mem); cudaFreeHost(c->h_batch_q); cudaFreeHost(c->h_batch_state);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
}

int rlr_api::ctx_new(rlr_store *s, rlr_ctx **out)
{
    rlr_ctx *c = new rlr_ctx();
    c->s = s;
    c->n_lists_cap = static_cast<uint32_t>(std::max(s->sm_count, 64));
#define CTX_TRY(expr)                                                                                    \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            cudaGetLastError();                                                                          \
            ctx_free(c);                                                                                 \
            return fail(e__ == cudaErrorMemoryAllocation ? RLR_ERR_OOM : RLR_ERR_CUDA, "%s failed: %s", #expr, \
                        cudaGetErrorString(e__));                                                        \
        }                                                                                                \
    } while (0)
    CTX_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto &e : c->ev) CTX_TRY(cudaEventCreate(&e));
    // room for one query per query group (throughput mode over a cluster stages them side by side)
    CTX_TRY(cudaMalloc(&c->d_query, rlr::kMaxQueryGroups * rlr::kQueryCap * sizeof(float)));
    CTX_TRY(cudaMemset(c->d_query, 0, rlr::kMaxQueryGroups * rlr::kQueryCap * sizeof(float)));
    CTX_TRY(cudaMalloc(&c->d_lex_rows, kLexCap * sizeof(uint32_t)));
    CTX_TRY(cudaMalloc(&c->d_lex_norm, kLexCap * sizeof(float)));
    // per-launch scan workspace, one slice per query group (kernels.cuh: ScanArgs)
    constexpr size_t kG = rlr::kMaxQueryGroups;
    CTX_TRY(cudaMalloc(&c->d_lists, kG * c->n_lists_cap * RLR_MAX_M * sizeof(rlr_cand)));
    CTX_TRY(cudaMalloc(&c->d_counts, kG * c->n_lists_cap * sizeof(uint32_t)));
    CTX_TRY(cudaMalloc(&c->d_ticket, 8 * sizeof(uint32_t)));
    CTX_TRY(cudaMemset(c->d_ticket, 0, 8 * sizeof(uint32_t)));
    CTX_TRY(cudaMalloc(&c->d_pub, kG * c->n_lists_cap * sizeof(uint32_t)));
    CTX_TRY(cudaMemset(c->d_pub, 0, kG * c->n_lists_cap * sizeof(uint32_t)));
    CTX_TRY(cudaMalloc(&c->d_tmp, (static_cast<size_t>(c->n_lists_cap) + 3) * RLR_MAX_M * sizeof(rlr_cand)));
    // (count, records) pairs live in one block each -- [u32 n | pad to 16 B | records] -- so that ONE D2H
    // copy brings back a result and its length
    memset(&c->lat, 0, sizeof c->lat);
    CTX_TRY(cudaMalloc(&c->d_pool_blk, 16 + RLR_MAX_M * sizeof(rlr_cand)));
    CTX_TRY(cudaMemset(c->d_pool_blk, 0, 16));
    c->d_pool_n = reinterpret_cast<uint32_t *>(c->d_pool_blk);
    c->d_pool = reinterpret_cast<rlr_cand *>(c->d_pool_blk + 16);
    CTX_TRY(cudaMalloc(&c->d_tri, static_cast<size_t>(RLR_MAX_M) * (RLR_MAX_M - 1) / 2 * sizeof(float)));
    CTX_TRY(cudaMalloc(&c->d_gather, static_cast<size_t>(RLR_MAX_M) * s->pitch * sizeof(float)));
    CTX_TRY(cudaMalloc(&c->d_sel_pos, RLR_MAX_M * sizeof(uint32_t)));
    CTX_TRY(cudaMalloc(&c->d_result_blk, 16 + RLR_MAX_M * sizeof(rlr_cand)));
    CTX_TRY(cudaMemset(c->d_result_blk, 0, 16));
    c->d_sel_n = reinterpret_cast<uint32_t *>(c->d_result_blk);
    c->d_result = reinterpret_cast<rlr_cand *>(c->d_result_blk + 16);
    CTX_TRY(cudaMalloc(&c->d_rows_in, RLR_MAX_M * sizeof(uint32_t)));
    CTX_TRY(cudaMalloc(&c->d_rel_in, RLR_MAX_M * sizeof(float)));
    CTX_TRY(cudaMalloc(&c->d_p_in, sizeof(uint32_t)));
    CTX_TRY(cudaMallocHost(&c->h_query, rlr::kMaxQueryGroups * rlr::kQueryCap * sizeof(float)));
    memset(c->h_query, 0, rlr::kMaxQueryGroups * rlr::kQueryCap * sizeof(float));
    CTX_TRY(cudaMallocHost(&c->h_lex_rows, kLexCap * sizeof(uint32_t)));
    CTX_TRY(cudaMallocHost(&c->h_lex_norm, kLexCap * sizeof(float)));
    CTX_TRY(cudaMallocHost(&c->h_result_blk, 16 + RLR_MAX_M * sizeof(rlr_cand)));
    memset(c->h_result_blk, 0, 16);
    c->h_result_n = reinterpret_cast<uint32_t *>(c->h_result_blk);
    c->h_result = reinterpret_cast<rlr_cand *>(c->h_result_blk + 16);
    CTX_TRY(cudaMallocHost(&c->h_u32, (RLR_MAX_M + 8) * sizeof(uint32_t)));
    CTX_TRY(cudaMallocHost(&c->h_rel, RLR_MAX_M * sizeof(float)));
#undef CTX_TRY
    *out = c;
    return RLR_OK;
}

void rlr_api::host_normalize(float *v, size_t n)
{
    // src/rag_engine.rs:1763-1771; volatile keeps gcc/nvcc-host from re-associating
    volatile float norm_sq = 0.0f;
    for (size_t i = 0; i < n; ++i) { volatile float p = v[i] * v[i]; norm_sq = norm_sq + p; }
    if (norm_sq > 1e-20f) {
        const float norm = sqrtf(norm_sq);
        for (size_t i = 0; i < n; ++i) v[i] = v[i] / norm;
    }
}

namespace {

// Stage the query in pinned memory, normalise (:494), upload.  Everything beyond dim
// stays zero so that the kernel's padded chunks contribute exact zeros.
int stage_query(rlr_ctx *c, const float *query, uint32_t dim, uint32_t flags, cudaStream_t st)
{
    rlr_store *s = c->s;
    if (!query) return fail(RLR_ERR_INVALID_ARG, "query is NULL");
    if (dim != s->dim)
        return fail(RLR_ERR_DIM_MISMATCH, "query has %u dims, store has %u (the reference would silently truncate, "
                    "src/rag_engine.rs:1778; this library refuses)", dim, s->dim);
    memcpy(c->h_query, query, dim * sizeof(float));
    for (uint32_t i = 0; i < dim; ++i)
        if (!std::isfinite(c->h_query[i])) return fail(RLR_ERR_NONFINITE, "query[%u] is not finite", i);
    if (!(flags & RLR_QUERY_PRENORMALIZED)) host_normalize(c->h_query, dim);
    const size_t n = ((s->dim + 63u) & ~63u) + 128;   // covers the last (zero) stage of either store copy
    CU_TRY(cudaMemcpyAsync(c->d_query, c->h_query, n * sizeof(float), cudaMemcpyHostToDevice, st));
    return RLR_OK;
}

// :505-530: lexical map -> (sorted local rows, score / max_lexical).
int stage_lex(rlr_ctx *c, const uint32_t *lex_rows, const float *lex_scores, uint32_t n_lex, uint32_t *out_n,
              cudaStream_t st)
{
    rlr_store *s = c->s;
    *out_n = 0;
    if (n_lex == 0) return RLR_OK;
    if (!lex_rows || !lex_scores) return fail(RLR_ERR_INVALID_ARG, "n_lex > 0 but lex_rows/lex_scores is NULL");
    if (n_lex > kLexCap) return fail(RLR_ERR_UNSUPPORTED, "n_lex %u exceeds %u", n_lex, kLexCap);
    float max_lexical = 0.0f; // fold(0.0_f32, f32::max).max(f32::EPSILON)
    for (uint32_t i = 0; i < n_lex; ++i) max_lexical = fmaxf(max_lexical, lex_scores[i]);
    max_lexical = fmaxf(max_lexical, 1.1920929e-07f);
    std::vector<std::pair<uint32_t, uint32_t>> order; // (row, original index); later duplicates win (HashMap collect)
    order.reserve(n_lex);
    for (uint32_t i = 0; i < n_lex; ++i) {
        const uint64_t g = lex_rows[i];
        if (g < s->row_base || g - s->row_base >= s->n_rows) continue; // `if let Some(chunk)`, :525
        order.emplace_back(static_cast<uint32_t>(g - s->row_base), i);
    }
    std::stable_sort(order.begin(), order.end(),
                     [](const std::pair<uint32_t, uint32_t> &a, const std::pair<uint32_t, uint32_t> &b) { return a.first < b.first; });
    uint32_t n = 0;
    for (size_t i = 0; i < order.size(); ++i) {
        if (i + 1 < order.size() && order[i + 1].first == order[i].first) continue;
        c->h_lex_rows[n] = order[i].first;
        c->h_lex_norm[n] = lex_scores[order[i].second] / max_lexical; // :527-530
        ++n;
    }
    if (n) {
        CU_TRY(cudaMemcpyAsync(c->d_lex_rows, c->h_lex_rows, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(c->d_lex_norm, c->h_lex_norm, n * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    *out_n = n;
    return RLR_OK;
}

// scan + merge on `st`: best m records of this store -> d_out / d_out_n
int enqueue_topm(rlr_ctx *c, const float *d_query, float w_e, float w_l, const uint32_t *d_lex_rows,
                 const float *d_lex_norm, uint32_t n_lex, uint32_t m, rlr_cand *d_out, uint32_t *d_out_n,
                 cudaStream_t st, cudaEvent_t ev_after_scan, bool half)
{
    rlr_store *s = c->s;
    rlr::ScanArgs a;
    memset(&a, 0, sizeof a);
    rlr::scan_plan(s->sm_count, s->smem_optin, static_cast<uint32_t>(s->n_rows), half ? s->pitch16 : s->pitch, half, &a);
    a.tmap = half ? &s->tmap16 : &s->tmap;
    a.d_query = d_query;
    a.n_rows = static_cast<uint32_t>(s->n_rows);
    a.row_base = static_cast<uint32_t>(s->row_base);
    a.pitch = half ? s->pitch16 : s->pitch;
    a.w_embed = w_e; a.w_lex = w_l;
    a.d_lex_rows = d_lex_rows; a.d_lex_norm = d_lex_norm; a.n_lex = n_lex;
    a.m = m;
    a.d_lists = c->d_lists; a.d_counts = c->d_counts;
    a.d_ticket = c->d_ticket; a.d_pub = c->d_pub; a.d_out = d_out; a.d_out_n = d_out_n; // cross-CTA merge happens in the scan's last CTA
    CU_TRY(rlr::scan_launch(a, st));
    ++c->launches;
    if (ev_after_scan) CU_TRY(cudaEventRecord(ev_after_scan, st));
    return RLR_OK;
}

} // namespace

void rlr_api::unpack(const rlr_cand *h, uint32_t n, uint32_t *rows, float *score, float *emb, float *lex)
{
    for (uint32_t i = 0; i < n; ++i) {
        if (rows) rows[i] = rlr::key_row(h[i].key);
        if (score) {
            const uint32_t b = rlr::bits_from_ord(static_cast<uint32_t>(h[i].key >> 32));
            memcpy(&score[i], &b, 4);
        }
        if (emb) emb[i] = h[i].emb;
        if (lex) lex[i] = h[i].lex;
    }
}

namespace {

int check_store(const rlr_store *s)
{
    if (!s) return fail(RLR_ERR_INVALID_ARG, "store is NULL");
    return RLR_OK;
}

float parse_env_weight(const char *name, float dflt)
{
    // parse_weight, src/rag_engine.rs:1813-1819: str::parse::<f32>() then finite && in [0,1]
    const char *v = getenv(name);
    if (!v || !*v) return dflt;
    if (isspace(static_cast<unsigned char>(v[0]))) return dflt;       // Rust does not trim
    if (strchr(v, 'x') || strchr(v, 'X')) return dflt;                // no hex floats in Rust
    char *end = nullptr;
    const float w = strtof(v, &end);
    if (end == v || *end != '\0') return dflt;
    if (!std::isfinite(w) || w < 0.0f || w > 1.0f) return dflt;
    return w;
}

float resolve_one(bool has, float w, float dflt)
{
    // resolve_weight, src/rag_engine.rs:1869-1873
    if (has && std::isfinite(w) && w >= 0.0f && w <= 1.0f) return w;
    return dflt;
}

} // namespace

// ---------------------------------------------------------------------------------
// library
// ---------------------------------------------------------------------------------
RLR_EXPORT int rlr_abi_version(void) { return RLR_ABI_VERSION; }
RLR_EXPORT const char *rlr_last_error(void) { return g_err.c_str(); }

RLR_EXPORT int rlr_device_count(int *out_count)
{
    if (!out_count) return fail(RLR_ERR_INVALID_ARG, "out_count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *out_count = 0;
        return fail(RLR_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *out_count = n;
    return RLR_OK;
}

RLR_EXPORT int rlr_device_query(int device, rlr_device_info *out)
{
    if (!out) return fail(RLR_ERR_INVALID_ARG, "out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(RLR_ERR_NO_DEVICE, "no CUDA device available");
    }
    if (device < 0 || device >= n) return fail(RLR_ERR_INVALID_ARG, "device %d out of range", device);
    cudaDeviceProp p;
    CU_TRY(cudaGetDeviceProperties(&p, device));
    memset(out, 0, sizeof *out);
    out->device = device;
    out->sm_count = p.multiProcessorCount;
    out->cc_major = p.major; out->cc_minor = p.minor;
    out->total_mem = p.totalGlobalMem;
    strncpy(out->name, p.name, sizeof(out->name) - 1);
    return RLR_OK;
}

RLR_EXPORT int rlr_normalize(float *v, size_t n)
{
    if (!v && n) return fail(RLR_ERR_INVALID_ARG, "v is NULL");
    host_normalize(v, n);
    return RLR_OK;
}

RLR_EXPORT int rlr_resolve_weights(const rlr_query_weights *o, rlr_resolved_weights *out)
{
    if (!out) return fail(RLR_ERR_INVALID_ARG, "out is NULL");
    // OnceLock caches, src/rag_engine.rs:1806-1841 (defaults :1801-1804)
    static float d_e, d_l, d_r, d_i;
    static std::once_flag once;
    std::call_once(once, [] {
        d_e = parse_env_weight("RAG_EMBEDDING_WEIGHT", 0.7f);
        d_l = parse_env_weight("RAG_LEXICAL_WEIGHT", 0.3f);
        d_r = parse_env_weight("RAG_RERANKER_WEIGHT", 0.7f);
        d_i = parse_env_weight("RAG_INITIAL_SCORE_WEIGHT", 0.3f);
    });
    out->embedding = resolve_one(o && (o->has & 1u), o ? o->embedding : 0.f, d_e);
    out->lexical = resolve_one(o && (o->has & 2u), o ? o->lexical : 0.f, d_l);
    out->reranker = resolve_one(o && (o->has & 4u), o ? o->reranker : 0.f, d_r);
    out->initial = resolve_one(o && (o->has & 8u), o ? o->initial : 0.f, d_i);
    return RLR_OK;
}

// ---------------------------------------------------------------------------------
// store
// ---------------------------------------------------------------------------------
namespace {

// dtype: 0 f32, 1 binary16, 2 bfloat16
int make_tmap(CUtensorMap *map, void *base, int dtype, uint32_t pitch_elems, uint64_t n_rows, uint32_t box_rows = rlr::kScanRows)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) return fail(RLR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    const uint32_t esz = dtype ? 2 : 4;
    const cuuint64_t gdim[2] = {pitch_elems, n_rows};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(pitch_elems) * esz};
    const cuuint32_t box[2] = {128u / esz, box_rows};          // 128-byte box rows: the SWIZZLE_128B span
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUresult r = enc(map, dt, 2, base, gdim, gstride,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RLR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return RLR_OK;
}

} // namespace

RLR_EXPORT int rlr_store_create(int device, uint32_t dim, uint64_t n_rows, const float *rows, uint64_t host_pitch,
                                uint64_t row_base, uint32_t flags, rlr_store **out)
{
    if (!out) return fail(RLR_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (dim == 0 || dim > RLR_MAX_DIM) return fail(RLR_ERR_INVALID_ARG, "dim %u not in 1..%d", dim, RLR_MAX_DIM);
    if (n_rows >= (1ull << 31)) return fail(RLR_ERR_UNSUPPORTED, "n_rows %llu >= 2^31 per store", (unsigned long long)n_rows);
    if (row_base + n_rows >= (1ull << 32)) return fail(RLR_ERR_UNSUPPORTED, "global rows must fit 32 bits");
    if ((flags & RLR_STORE_KEEP_F16) && (flags & RLR_STORE_F16_ONLY))
        return fail(RLR_ERR_INVALID_ARG, "RLR_STORE_KEEP_F16 and RLR_STORE_F16_ONLY are exclusive");
    if ((flags & RLR_STORE_KEEP_BF16) && (flags & RLR_STORE_F16_ONLY))
        return fail(RLR_ERR_INVALID_ARG, "RLR_STORE_KEEP_BF16 needs the f32 rows (exclusive with RLR_STORE_F16_ONLY)");
    if (host_pitch == 0) host_pitch = dim;
    if (host_pitch < dim) return fail(RLR_ERR_INVALID_ARG, "host_pitch %llu < dim %u", (unsigned long long)host_pitch, dim);
    int rc = ensure_device(device);
    if (rc) return rc;

    rlr_store *s = new rlr_store();
    s->device = device;
    s->dim = dim;
    s->pitch = (dim + 31u) & ~31u;
    s->pitch16 = (dim + 63u) & ~63u;
    s->n_rows = n_rows;
    s->capacity = n_rows;
    s->row_base = row_base;
    s->flags = flags;
    {
        std::lock_guard<std::mutex> lk(g_dev_mu);
        s->sm_count = g_dev[device].sm_count;
        s->smem_optin = g_dev[device].smem_optin;
    }
    memset(&s->tmap, 0, sizeof s->tmap);
    memset(&s->tmap16, 0, sizeof s->tmap16);
    memset(&s->tmap_bf16, 0, sizeof s->tmap_bf16);
    memset(&s->tmap_small, 0, sizeof s->tmap_small);
    const bool want32 = !(flags & RLR_STORE_F16_ONLY);
    const bool want16 = flags & (RLR_STORE_F16_ONLY | RLR_STORE_KEEP_F16);
    if (n_rows) {
        cudaError_t e = cudaSuccess;
        size_t bytes = 0;
        if (want32) { bytes = static_cast<size_t>(n_rows) * s->pitch * sizeof(float); e = cudaMalloc(&s->d_rows, bytes); }
        if (e == cudaSuccess && want16) { bytes = static_cast<size_t>(n_rows) * s->pitch16 * 2; e = cudaMalloc(&s->d_rows16, bytes); }
        if (e == cudaSuccess && (flags & RLR_STORE_KEEP_BF16)) { bytes = static_cast<size_t>(n_rows) * s->pitch16 * 2; e = cudaMalloc(&s->d_rows_bf16, bytes); }
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaFree(s->d_rows); cudaFree(s->d_rows16); cudaFree(s->d_rows_bf16);
            delete s;
            return fail(RLR_ERR_OOM, "cudaMalloc of %zu bytes for the store failed: %s", bytes, cudaGetErrorString(e));
        }
        rc = RLR_OK;
        if (want32) rc = make_tmap(&s->tmap, s->d_rows, 0, s->pitch, n_rows);
        if (rc == RLR_OK && want16) rc = make_tmap(&s->tmap16, s->d_rows16, 1, s->pitch16, n_rows);
        if (rc == RLR_OK && s->d_rows_bf16) rc = make_tmap(&s->tmap_bf16, s->d_rows_bf16, 2, s->pitch16, n_rows);
        s->rpt = rlr::scan_rows_per_tile(s->sm_count, n_rows);
        if (rc == RLR_OK && want32 && s->rpt != rlr::kScanRows) rc = make_tmap(&s->tmap_small, s->d_rows, 0, s->pitch, n_rows, s->rpt);
        if (rc != RLR_OK) { cudaFree(s->d_rows); cudaFree(s->d_rows16); cudaFree(s->d_rows_bf16); delete s; return rc; }
    }
    *out = s;
    if (rows && n_rows) {
        rc = rlr_store_upload(s, 0, n_rows, rows, host_pitch);
        if (rc) { std::string keep = g_err; rlr_store_destroy(s); *out = nullptr; g_err = keep; return rc; }
    }
    return RLR_OK;
}

RLR_EXPORT int rlr_store_destroy(rlr_store *s)
{
    if (!s) return RLR_OK;
    cudaSetDevice(s->device);
    for (rlr_ctx *c : s->free_ctx) ctx_free(c);
    cudaFree(s->d_rows);
    cudaFree(s->d_rows16);
    cudaFree(s->d_rows_bf16);
    cudaGetLastError();
    delete s;
    return RLR_OK;
}

RLR_EXPORT int rlr_store_info_get(const rlr_store *s, rlr_store_info *out)
{
    if (int rc = check_store(s)) return rc;
    if (!out) return fail(RLR_ERR_INVALID_ARG, "out is NULL");
    out->n_rows = s->n_rows; out->row_base = s->row_base; out->dim = s->dim; out->pitch = s->pitch;
    out->device = s->device; out->flags = s->flags;
    out->bytes_device = (s->d_rows ? s->n_rows * s->pitch * sizeof(float) : 0) + (s->d_rows16 ? s->n_rows * s->pitch16 * 2 : 0) +
                        (s->d_rows_bf16 ? s->n_rows * s->pitch16 * 2 : 0);
    return RLR_OK;
}

RLR_EXPORT int rlr_store_upload(rlr_store *s, uint64_t row0, uint64_t n, const float *rows, uint64_t host_pitch)
{
    if (int rc = check_store(s)) return rc;
    if (!rows && n) return fail(RLR_ERR_INVALID_ARG, "rows is NULL");
    if (row0 + n > s->n_rows) return fail(RLR_ERR_INVALID_ARG, "rows [%llu,%llu) outside the store", (unsigned long long)row0, (unsigned long long)(row0 + n));
    if (host_pitch == 0) host_pitch = s->dim;
    if (host_pitch < s->dim) return fail(RLR_ERR_INVALID_ARG, "host_pitch < dim");
    if (n == 0) return RLR_OK;
    CU_TRY(cudaSetDevice(s->device));
    // slabs: cudaMemcpy2D heights stay small, and an f16-only store needs only a slab of f32 staging
    const uint64_t slab = 1u << 18;
    float *d_stage = nullptr;
    if (!s->d_rows) CU_TRY(cudaMalloc(&d_stage, std::min(slab, n) * s->pitch * sizeof(float)));
    int rc = RLR_OK;
    for (uint64_t r = 0; r < n && rc == RLR_OK; r += slab) {
        const uint64_t cnt = std::min(slab, n - r);
        float *dst32 = s->d_rows ? s->d_rows + (row0 + r) * s->pitch : d_stage;
        cudaError_t e = cudaSuccess;
        if (s->pitch != s->dim) e = cudaMemset(dst32, 0, cnt * s->pitch * sizeof(float));
        if (e == cudaSuccess)
            e = cudaMemcpy2D(dst32, s->pitch * sizeof(float), rows + r * host_pitch, host_pitch * sizeof(float),
                             s->dim * sizeof(float), cnt, cudaMemcpyHostToDevice);
        if (e == cudaSuccess && (s->flags & RLR_STORE_NORMALIZE_ON_UPLOAD))      // :1678-1680 on the device, same bits
            e = rlr::normalize_rows_launch(dst32, s->pitch, s->dim, cnt, 0);
        if (e == cudaSuccess && (s->flags & RLR_STORE_CHECK_FINITE)) {
            uint32_t *d_flag = nullptr, h_flag = 0;
            e = cudaMalloc(&d_flag, sizeof(uint32_t));
            if (e == cudaSuccess) e = cudaMemset(d_flag, 0, sizeof(uint32_t));
            if (e == cudaSuccess) e = rlr::finite_check_launch(dst32, cnt * s->pitch, d_flag, 0);
            if (e == cudaSuccess) e = cudaMemcpy(&h_flag, d_flag, sizeof h_flag, cudaMemcpyDeviceToHost);
            cudaFree(d_flag);
            if (e == cudaSuccess && h_flag) rc = fail(RLR_ERR_NONFINITE, "uploaded rows contain NaN/Inf");
        }
        if (e == cudaSuccess && rc == RLR_OK && s->d_rows16) {
            e = rlr::to_half_launch(dst32, s->pitch, static_cast<__half *>(s->d_rows16) + (row0 + r) * s->pitch16, s->pitch16,
                                    s->dim, cnt, 0);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
        }
        if (e == cudaSuccess && rc == RLR_OK && s->d_rows_bf16)
            e = rlr::to_bf16_launch(dst32, s->pitch, static_cast<uint8_t *>(s->d_rows_bf16) + (row0 + r) * s->pitch16 * 2, s->pitch16, s->dim, cnt, 0);
        if (e != cudaSuccess) {
            cudaGetLastError();
            rc = fail(e == cudaErrorMemoryAllocation ? RLR_ERR_OOM : RLR_ERR_CUDA, "store upload failed: %s", cudaGetErrorString(e));
        }
    }
    // The upload ran on the legacy stream; search streams are cudaStreamNonBlocking and do not wait for it.
    // Return only when every row (memset, copy, normalize, f16 copy) has landed, and report async errors.
    {
        cudaError_t e = cudaStreamSynchronize(0);
        if (e != cudaSuccess && rc == RLR_OK) {
            cudaGetLastError();
            rc = fail(RLR_ERR_CUDA, "store upload failed: %s", cudaGetErrorString(e));
        }
    }
    cudaFree(d_stage);
    return rc;
}

RLR_EXPORT int rlr_store_read_rows(const rlr_store *s, const uint32_t *rows, uint64_t n, float *out)
{
    if (int rc = check_store(s)) return rc;
    if ((!rows || !out) && n) return fail(RLR_ERR_INVALID_ARG, "rows/out is NULL");
    CU_TRY(cudaSetDevice(s->device));
    std::vector<__half> tmp(s->d_rows ? 0 : s->dim);
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t g = rows[i];
        if (g < s->row_base || g - s->row_base >= s->n_rows) return fail(RLR_ERR_INVALID_ARG, "row %llu not in this store", (unsigned long long)g);
        if (s->d_rows) {
            CU_TRY(cudaMemcpy(out + i * s->dim, s->d_rows + (g - s->row_base) * s->pitch, s->dim * sizeof(float), cudaMemcpyDeviceToHost));
        } else {   // f16-only store: widen on the host (exact)
            CU_TRY(cudaMemcpy(tmp.data(), static_cast<const __half *>(s->d_rows16) + (g - s->row_base) * s->pitch16, s->dim * 2, cudaMemcpyDeviceToHost));
            for (uint32_t c = 0; c < s->dim; ++c) out[i * s->dim + c] = __half2float(tmp[c]);
        }
    }
    return RLR_OK;
}

// ---------------------------------------------------------------------------------
// store mutation (add_document, src/rag_engine.rs:347-386, under the write lock): drop a
// document's rows, append its new rows.  Mutators require exclusivity (header, "threading").
// ---------------------------------------------------------------------------------
namespace {

int store_remap(rlr_store *s)
{
    if (s->n_rows == 0) return RLR_OK;
    if (s->d_rows) if (int rc = make_tmap(&s->tmap, s->d_rows, 0, s->pitch, s->n_rows)) return rc;
    if (s->d_rows16) if (int rc = make_tmap(&s->tmap16, s->d_rows16, 1, s->pitch16, s->n_rows)) return rc;
    if (s->d_rows_bf16) if (int rc = make_tmap(&s->tmap_bf16, s->d_rows_bf16, 2, s->pitch16, s->n_rows)) return rc;
    s->rpt = rlr::scan_rows_per_tile(s->sm_count, s->n_rows);
    if (s->d_rows && s->rpt != rlr::kScanRows) if (int rc = make_tmap(&s->tmap_small, s->d_rows, 0, s->pitch, s->n_rows, s->rpt)) return rc;
    return RLR_OK;
}

int store_grow(rlr_store *s, uint64_t want)
{
    if (want <= s->capacity) return RLR_OK;
    uint64_t cap = std::max<uint64_t>(want, s->capacity + s->capacity / 2 + 1024);
    if (cap >= (1ull << 31)) cap = (1ull << 31) - 1;
    if (cap < want) return fail(RLR_ERR_UNSUPPORTED, "store would exceed 2^31 rows");
    const bool want32 = !(s->flags & RLR_STORE_F16_ONLY);
    const bool want16 = s->flags & (RLR_STORE_F16_ONLY | RLR_STORE_KEEP_F16);
    float *n32 = nullptr;
    void *n16 = nullptr;
    if (want32) CU_TRY(cudaMalloc(&n32, cap * s->pitch * sizeof(float)));
    if (want16) {
        cudaError_t e = cudaMalloc(&n16, cap * s->pitch16 * 2);
        if (e != cudaSuccess) { cudaGetLastError(); cudaFree(n32); return fail(RLR_ERR_OOM, "cudaMalloc failed growing the store: %s", cudaGetErrorString(e)); }
    }
    void *nbf = nullptr;
    if (s->flags & RLR_STORE_KEEP_BF16) {
        cudaError_t e = cudaMalloc(&nbf, cap * s->pitch16 * 2);
        if (e != cudaSuccess) { cudaGetLastError(); cudaFree(n32); cudaFree(n16); return fail(RLR_ERR_OOM, "cudaMalloc failed growing the store: %s", cudaGetErrorString(e)); }
    }
    if (s->n_rows) {
        if (want32) CU_TRY(cudaMemcpy(n32, s->d_rows, s->n_rows * s->pitch * sizeof(float), cudaMemcpyDeviceToDevice));
        if (want16) CU_TRY(cudaMemcpy(n16, s->d_rows16, s->n_rows * s->pitch16 * 2, cudaMemcpyDeviceToDevice));
        if (nbf && s->d_rows_bf16) CU_TRY(cudaMemcpy(nbf, s->d_rows_bf16, s->n_rows * s->pitch16 * 2, cudaMemcpyDeviceToDevice));
    }
    cudaFree(s->d_rows); cudaFree(s->d_rows16); cudaFree(s->d_rows_bf16);
    s->d_rows = n32; s->d_rows16 = n16; s->d_rows_bf16 = nbf; s->capacity = cap;
    return RLR_OK;
}

} // namespace

RLR_EXPORT int rlr_store_reserve(rlr_store *s, uint64_t capacity_rows)
{
    if (int rc = check_store(s)) return rc;
    if (int rc = ensure_device(s->device)) return rc;
    if (int rc = store_grow(s, capacity_rows)) return rc;
    return store_remap(s);
}

RLR_EXPORT int rlr_store_append(rlr_store *s, uint64_t n, const float *rows, uint64_t host_pitch, uint64_t *out_first_row)
{
    if (int rc = check_store(s)) return rc;
    if (!rows && n) return fail(RLR_ERR_INVALID_ARG, "rows is NULL");
    if (int rc = ensure_device(s->device)) return rc;
    if (s->row_base + s->n_rows + n >= (1ull << 32)) return fail(RLR_ERR_UNSUPPORTED, "global rows must fit 32 bits");
    if (out_first_row) *out_first_row = s->row_base + s->n_rows;
    if (n == 0) return RLR_OK;
    if (int rc = store_grow(s, s->n_rows + n)) return rc;
    const uint64_t first = s->n_rows;
    s->n_rows += n;
    if (int rc = rlr_store_upload(s, first, n, rows, host_pitch)) { s->n_rows = first; return rc; }
    return store_remap(s);
}

RLR_EXPORT int rlr_store_remove_rows(rlr_store *s, const uint32_t *rows, uint64_t n, uint32_t *out_moved_from,
                                     uint32_t *out_moved_to, uint64_t *out_n_moved)
{
    if (int rc = check_store(s)) return rc;
    if (out_n_moved) *out_n_moved = 0;
    if (n == 0) return RLR_OK;
    if (!rows) return fail(RLR_ERR_INVALID_ARG, "rows is NULL");
    if (int rc = ensure_device(s->device)) return rc;
    // local, sorted, unique
    std::vector<uint32_t> rm;
    rm.reserve(n);
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t g = rows[i];
        if (g < s->row_base || g - s->row_base >= s->n_rows) return fail(RLR_ERR_INVALID_ARG, "row %llu not in this store", (unsigned long long)g);
        rm.push_back(static_cast<uint32_t>(g - s->row_base));
    }
    std::sort(rm.begin(), rm.end());
    rm.erase(std::unique(rm.begin(), rm.end()), rm.end());
    const uint64_t new_n = s->n_rows - rm.size();
    // holes below new_n are filled, in order, by the surviving rows of the tail [new_n, n_rows)
    std::vector<uint32_t> from, to;
    size_t hi = rm.size();                       // rm[lo_end..) are removals inside the tail
    while (hi > 0 && rm[hi - 1] >= new_n) --hi;
    size_t tail_rm = hi;
    uint64_t src = new_n;
    for (size_t h = 0; h < hi; ++h) {
        while (tail_rm < rm.size() && rm[tail_rm] == src) { ++src; ++tail_rm; }
        from.push_back(static_cast<uint32_t>(src++));
        to.push_back(rm[h]);
    }
    if (!from.empty()) {
        uint32_t *d_from = nullptr, *d_to = nullptr;
        CU_TRY(cudaMalloc(&d_from, from.size() * 4));
        cudaError_t e = cudaMalloc(&d_to, to.size() * 4);
        if (e == cudaSuccess) e = cudaMemcpy(d_from, from.data(), from.size() * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_to, to.data(), to.size() * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess && s->d_rows) e = rlr::move_rows_launch(s->d_rows, s->pitch * 4, d_from, d_to, static_cast<uint32_t>(from.size()), 0);
        if (e == cudaSuccess && s->d_rows16) e = rlr::move_rows_launch(s->d_rows16, s->pitch16 * 2, d_from, d_to, static_cast<uint32_t>(from.size()), 0);
        if (e == cudaSuccess && s->d_rows_bf16) e = rlr::move_rows_launch(s->d_rows_bf16, s->pitch16 * 2, d_from, d_to, static_cast<uint32_t>(from.size()), 0);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        cudaFree(d_from); cudaFree(d_to);
        CU_TRY(e);
    }
    for (size_t i = 0; i < from.size(); ++i) {
        if (out_moved_from) out_moved_from[i] = static_cast<uint32_t>(s->row_base + from[i]);
        if (out_moved_to) out_moved_to[i] = static_cast<uint32_t>(s->row_base + to[i]);
    }
    if (out_n_moved) *out_n_moved = from.size();
    s->n_rows = new_n;
    return store_remap(s);
}

RLR_EXPORT int rlr_store_fill_synthetic(rlr_store *s, int kind, uint64_t seed, uint64_t centroid_seed,
                                        uint32_t n_clusters, float sigma)
{
    if (int rc = check_store(s)) return rc;
    if (kind != RLR_SYNTH_IID && kind != RLR_SYNTH_CLUSTERED) return fail(RLR_ERR_INVALID_ARG, "unknown synthetic kind %d", kind);
    if (kind == RLR_SYNTH_CLUSTERED && n_clusters == 0) return fail(RLR_ERR_INVALID_ARG, "n_clusters must be > 0");
    CU_TRY(cudaSetDevice(s->device));
    CU_TRY(rlr::synth_launch(s->d_rows, s->pitch, s->d_rows16, s->pitch16, s->dim, s->row_base,
                             static_cast<uint32_t>(s->n_rows), kind, seed, centroid_seed, n_clusters, sigma, 0));
    if (s->d_rows_bf16) CU_TRY(rlr::to_bf16_launch(s->d_rows, s->pitch, s->d_rows_bf16, s->pitch16, s->dim, s->n_rows, 0));
    CU_TRY(cudaDeviceSynchronize());
    return RLR_OK;
}

// ---------------------------------------------------------------------------------
// hot path, host-facing
// ---------------------------------------------------------------------------------
namespace {

// ---- latency path (small stores; LatParams in kernels.cuh) ----
constexpr uint64_t kLatMaxRows = 262144;     // beyond this the scan itself dominates and the regular path is as good
constexpr uint64_t kLatL2Bytes = 48ull << 20;  // stores up to this size are kept L2-resident between queries (126 MB L2, two 63 MB halves)

bool lat_eligible(const rlr_store *s, uint32_t flags)
{
    static const bool off = getenv("RLR_NO_LATENCY_PATH") != nullptr;
    return !off && !(s->flags & RLR_STORE_NO_LATENCY_PATH) && s->d_rows != nullptr && !s->use_half(flags) && s->pitch <= rlr::kLatQFloats && s->n_rows <= kLatMaxRows;
}

// query -> the parameter block: copy, NaN/Inf check, normalize (:494).  Floats beyond dim stay zero.
int lat_stage_query(rlr_ctx *c, const float *query, uint32_t dim, uint32_t flags)
{
    rlr_store *s = c->s;
    if (!query) return fail(RLR_ERR_INVALID_ARG, "query is NULL");
    if (dim != s->dim)
        return fail(RLR_ERR_DIM_MISMATCH, "query has %u dims, store has %u (the reference would silently truncate, "
                    "src/rag_engine.rs:1778; this library refuses)", dim, s->dim);
    memcpy(c->lat.q, query, dim * sizeof(float));
    for (uint32_t i = 0; i < dim; ++i)
        if (!std::isfinite(c->lat.q[i])) return fail(RLR_ERR_NONFINITE, "query[%u] is not finite", i);
    if (!(flags & RLR_QUERY_PRENORMALIZED)) host_normalize(c->lat.q, dim);
    return RLR_OK;
}

// :505-530 into the parameter block; *fits = false when there are more local pairs than the block holds
int lat_stage_lex(rlr_ctx *c, const uint32_t *lex_rows, const float *lex_scores, uint32_t n_lex, uint32_t *out_n, bool *fits)
{
    rlr_store *s = c->s;
    *out_n = 0; *fits = true;
    if (n_lex == 0) return RLR_OK;
    if (!lex_rows || !lex_scores) return fail(RLR_ERR_INVALID_ARG, "n_lex > 0 but lex_rows/lex_scores is NULL");
    if (n_lex > kLexCap) return fail(RLR_ERR_UNSUPPORTED, "n_lex %u exceeds %u", n_lex, kLexCap);
    float max_lexical = 0.0f;
    for (uint32_t i = 0; i < n_lex; ++i) max_lexical = fmaxf(max_lexical, lex_scores[i]);
    max_lexical = fmaxf(max_lexical, 1.1920929e-07f);
    std::vector<std::pair<uint32_t, uint32_t>> order;
    order.reserve(n_lex);
    for (uint32_t i = 0; i < n_lex; ++i) {
        const uint64_t g = lex_rows[i];
        if (g < s->row_base || g - s->row_base >= s->n_rows) continue;
        order.emplace_back(static_cast<uint32_t>(g - s->row_base), i);
    }
    std::stable_sort(order.begin(), order.end(),
                     [](const std::pair<uint32_t, uint32_t> &a, const std::pair<uint32_t, uint32_t> &b) { return a.first < b.first; });
    uint32_t n = 0;
    for (size_t i = 0; i < order.size(); ++i) {
        if (i + 1 < order.size() && order[i + 1].first == order[i].first) continue;
        if (n == rlr::kLatLex) { *fits = false; return RLR_OK; }
        c->lat.lex_rows[n] = order[i].first;
        c->lat.lex_norm[n] = lex_scores[order[i].second] / max_lexical;
        ++n;
    }
    *out_n = n;
    return RLR_OK;
}

// Wait for the kernel's completion word in mapped pinned memory.  Polls the stream now and then so that a kernel
// that died (or a launch that never ran) becomes an error instead of a hang.
int lat_wait(rlr_ctx *c, cudaStream_t st)
{
    volatile unsigned long long *flag = reinterpret_cast<volatile unsigned long long *>(c->h_result_blk + 8);
    const unsigned long long seq = c->lat_seq;
    for (uint32_t spins = 1;; ++spins) {
        if (*flag == seq) break;
        if ((spins & 0x1fffu) == 0) {
            const cudaError_t e = cudaStreamQuery(st);
            if (e == cudaSuccess) {
                if (*flag == seq) break;
                return fail(RLR_ERR_CUDA, "latency path: the kernels finished without delivering a result");
            }
            if (e != cudaErrorNotReady) { cudaGetLastError(); return fail(RLR_ERR_CUDA, "latency path: %s", cudaGetErrorString(e)); }
        }
        __builtin_ia32_pause();
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);     // the records were written before the flag (release, system scope)
    return RLR_OK;
}

// One request on the latency path.  pool: records the scan delivers; do_mmr: diversify them to top_k.
int lat_search(rlr_store *s, rlr_ctx *c, uint32_t flags, const rlr_resolved_weights *w, uint32_t n_lex, uint32_t pool, bool do_mmr,
               uint32_t top_k, float lambda, uint32_t *out_rows, float *out_score, float *out_emb, float *out_lex, uint32_t *out_n)
{
    cudaStream_t st = c->stream;
    const uint64_t launches0 = c->launches;
    const bool timed = flags & RLR_WANT_TIMINGS;
    rlr::LatParams &lp = c->lat;
    rlr::ScanArgs a;
    memset(&a, 0, sizeof a);
    rlr::scan_plan_small(s->sm_count, s->smem_optin, static_cast<uint32_t>(s->n_rows), s->pitch, s->rpt, &a);
    a.tmap = s->rpt == rlr::kScanRows ? &s->tmap : &s->tmap_small;
    a.n_rows = static_cast<uint32_t>(s->n_rows);
    a.row_base = static_cast<uint32_t>(s->row_base);
    a.pitch = s->pitch;
    a.w_embed = w->embedding; a.w_lex = w->lexical; a.n_lex = n_lex;
    a.m = pool;
    a.d_lists = c->d_lists; a.d_counts = c->d_counts; a.d_ticket = c->d_ticket; a.d_pub = c->d_pub;
    a.lat = &lp;
    const unsigned long long seq = ++c->lat_seq;
    *reinterpret_cast<volatile unsigned long long *>(c->h_result_blk + 8) = 0;
    lp.top_k = top_k; lp.lambda = lambda; lp.pitch = s->pitch; lp.g_rows = s->d_rows; lp.d_sel_pos = c->d_sel_pos;
    lp.keep_l2 = static_cast<uint64_t>(s->n_rows) * s->pitch * 4 <= kLatL2Bytes ? 1u : 0u;
    lp.result = c->h_result; lp.result_n = c->h_result_n;          // mapped pinned memory (UVA: same pointers on the device)
    lp.flag = reinterpret_cast<unsigned long long *>(c->h_result_blk + 8); lp.seq = seq;
    const size_t ring_bytes = static_cast<size_t>(a.n_stages) * rlr::kScanChunks * rlr::kScanRows * 128;
    const bool fuse = do_mmr && pool <= rlr::kLatFusePool && 4096 + static_cast<size_t>(pool) * (s->pitch * 4 + 128) <= ring_bytes;
    uint32_t cap = pool;
    if (timed) CU_TRY(cudaEventRecord(c->ev[0], st));
    if (!do_mmr) {              // search(top_k): the merged list goes straight to the host
        lp.mode = 1;
        a.d_out = c->h_result; a.d_out_n = c->h_result_n;
        CU_TRY(rlr::scan_launch(a, st));
        ++c->launches;
        if (timed) { CU_TRY(cudaEventRecord(c->ev[1], st)); CU_TRY(cudaEventRecord(c->ev[2], st)); }
    } else if (fuse) {          // scan + merge + pairwise + greedy + delivery: ONE launch
        lp.mode = 2;
        a.d_out = c->d_pool; a.d_out_n = c->d_pool_n;
        static const bool trace = getenv("RLR_DEBUG_LAT_TRACE") != nullptr;      // dev-only: phase timestamps of one request
        if (trace) {
            const int g = a.grid;
            unsigned long long *d_tr = nullptr;
            std::vector<unsigned long long> h(5 * g + 16, 0);
            CU_TRY(cudaMalloc(&d_tr, h.size() * 8));
            CU_TRY(cudaMemset(d_tr, 0, h.size() * 8));
            a.d_trace = d_tr;
            unsigned long long host_t0 = 0;
            { timespec ts; clock_gettime(CLOCK_REALTIME, &ts); host_t0 = ts.tv_sec * 1000000000ull + ts.tv_nsec; }
            CU_TRY(rlr::scan_launch(a, st));
            CU_TRY(cudaStreamSynchronize(st));
            CU_TRY(cudaMemcpy(h.data(), d_tr, h.size() * 8, cudaMemcpyDeviceToHost));
            cudaFree(d_tr);
            a.d_trace = nullptr;
            unsigned long long t0 = ~0ull, t0max = 0, loop_max = 0, loop_min = ~0ull, list_max = 0;
            for (int i = 0; i < g; ++i) { t0 = std::min(t0, h[i]); t0max = std::max(t0max, h[i]); loop_max = std::max(loop_max, h[g + i]); loop_min = std::min(loop_min, h[g + i]); list_max = std::max(list_max, h[2 * g + i]); }
            const unsigned long long *tr = h.data() + 4 * g;
            unsigned long long entry_min = ~0ull, entry_max = 0;
            for (int i = 0; i < g; ++i) { entry_min = std::min(entry_min, tr[16 + i]); entry_max = std::max(entry_max, tr[16 + i]); }
            (void)host_t0;
            fprintf(stderr, "[lat trace grid=%d rpt=%u] times from the first CTA's kernel entry: entry last %.1f; prologue done first/last %.1f/%.1f; "
                            "scan loop end first/last %.1f/%.1f; lists written %.1f; merge start %.1f, heads sorted %.1f, merged %.1f, rows staged %.1f, "
                            "pairwise %.1f, greedy %.1f, delivered %.1f us\n",
                    g, s->rpt, (entry_max - entry_min) / 1e3, (t0 - entry_min) / 1e3, (t0max - entry_min) / 1e3, (loop_min - entry_min) / 1e3,
                    (loop_max - entry_min) / 1e3, (list_max - entry_min) / 1e3, (tr[0] - entry_min) / 1e3, (tr[2] - entry_min) / 1e3,
                    (tr[4] - entry_min) / 1e3, (tr[5] - entry_min) / 1e3, (tr[6] - entry_min) / 1e3, (tr[10] - entry_min) / 1e3, (tr[7] - entry_min) / 1e3);
            *reinterpret_cast<volatile unsigned long long *>(c->h_result_blk + 8) = 0;
        }
        CU_TRY(rlr::scan_launch(a, st));
        ++c->launches;
        if (timed) { CU_TRY(cudaEventRecord(c->ev[1], st)); CU_TRY(cudaEventRecord(c->ev[2], st)); CU_TRY(cudaEventRecord(c->ev[3], st)); }
        cap = std::min<uint32_t>(pool, std::max<uint32_t>(top_k, 1));
    } else {                    // larger pools: the two MMR kernels follow; the greedy kernel delivers
        lp.mode = 0;
        a.d_out = c->d_pool; a.d_out_n = c->d_pool_n;
        CU_TRY(rlr::scan_launch(a, st));
        ++c->launches;
        if (timed) { CU_TRY(cudaEventRecord(c->ev[1], st)); CU_TRY(cudaEventRecord(c->ev[2], st)); }
        rlr::MmrArgs ma;
        memset(&ma, 0, sizeof ma);
        ma.half = 0;
        ma.d_emb = s->d_rows; ma.pitch = s->pitch; ma.dim = s->dim;
        ma.d_cands = c->d_pool; ma.d_n = c->d_pool_n;
        ma.row_base = static_cast<uint32_t>(s->row_base); ma.use_rows = 1;
        ma.p_cap = pool; ma.top_k = top_k; ma.lambda = lambda;
        ma.d_tri = c->d_tri; ma.d_sel_pos = c->d_sel_pos; ma.d_sel_n = c->h_result_n; ma.d_result = c->h_result;
        ma.max_smem_optin = s->smem_optin;
        ma.done_flag = lp.flag; ma.done_seq = seq;
        uint32_t l = 0;
        CU_TRY(rlr::mmr_launch(ma, st, &l));
        c->launches += l;
        if (timed) CU_TRY(cudaEventRecord(c->ev[3], st));
        cap = std::min<uint32_t>(pool, std::max<uint32_t>(top_k, 1));
    }
    if (int rc = lat_wait(c, st)) return rc;
    const uint32_t n = std::min(c->h_result_n[0], cap);
    unpack(c->h_result, n, out_rows, out_score, out_emb, out_lex);
    *out_n = n;
    if (timed) {
        CU_TRY(cudaEventSynchronize(c->ev[do_mmr ? 3 : 2]));
        rlr_timings t = {0, 0, 0, 0, 0};
        cudaEventElapsedTime(&t.scan_ms, c->ev[0], c->ev[1]);
        if (do_mmr) cudaEventElapsedTime(&t.mmr_ms, c->ev[2], c->ev[3]);
        cudaEventElapsedTime(&t.total_ms, c->ev[0], c->ev[do_mmr ? 3 : 2]);
        cudaGetLastError();
        t.launches = static_cast<uint32_t>(c->launches - launches0);
        g_timings = t;
    }
    return RLR_OK;
}

void timings_from_events(rlr_ctx *c, bool has_mmr)
{
    rlr_timings t = {0, 0, 0, 0, 0};
    cudaEventElapsedTime(&t.scan_ms, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&t.merge_ms, c->ev[1], c->ev[2]);
    if (has_mmr) cudaEventElapsedTime(&t.mmr_ms, c->ev[2], c->ev[3]);
    cudaEventElapsedTime(&t.total_ms, c->ev[0], has_mmr ? c->ev[3] : c->ev[2]);
    cudaGetLastError();
    g_timings = t;
}

} // namespace

namespace {

// A text query's lexical stage on the device (bm25.cu): LexicalIndex::score(query, 5 * top_k) (:505) enqueued on the
// ctx's stream, its (sorted local rows, score / max_lexical) form written into the ctx's lexical arrays.
struct TextQuery {
    rlr_bm25 *ix = nullptr;           // null: the caller supplied (row, score) pairs instead
    const uint32_t *terms = nullptr;
    uint32_t n_terms = 0;
};
struct Bm25WsGuard {
    rlr_bm25 *ix = nullptr;
    void *ws = nullptr;
    ~Bm25WsGuard() { if (ws) rlr_api_bm25_ws_release(ix, ws); }
};
int stage_text(rlr_ctx *c, const TextQuery &tq, uint32_t limit, Bm25WsGuard &g, cudaStream_t st, uint32_t *out_nl)
{
    *out_nl = 0;
    if (rlr_api_bm25_store(tq.ix) != c->s) return fail(RLR_ERR_INVALID_ARG, "the BM25 index belongs to another store");
    if (limit > kLexCap) return fail(RLR_ERR_UNSUPPORTED, "5 * m = %u lexical candidates exceed %u", limit, kLexCap);
    g.ix = tq.ix;
    if (int rc = rlr_api_bm25_ws_acquire(tq.ix, &g.ws)) return rc;
    bool active = false;
    uint32_t launches = 0;
    if (int rc = rlr_api_bm25_enqueue(tq.ix, g.ws, tq.terms, tq.n_terms, limit, c->d_lex_rows, c->d_lex_norm, limit, nullptr, nullptr,
                                      nullptr, st, &active, nullptr, &launches))
        return rc;
    if (active) { *out_nl = limit; c->launches += launches; }
    return RLR_OK;
}

int search_topm_impl(rlr_store *s, const float *query, uint32_t dim, uint32_t flags,
                               const rlr_resolved_weights *w, const uint32_t *lex_rows, const float *lex_scores,
                               uint32_t n_lex, const TextQuery &tq, uint32_t m, uint32_t *out_rows, float *out_combined, float *out_emb,
                               float *out_lex, uint32_t *out_n)
{
    if (int rc = check_store(s)) return rc;
    if (!out_rows || !out_n) return fail(RLR_ERR_INVALID_ARG, "out_rows/out_n is NULL");
    if (!w) return fail(RLR_ERR_INVALID_ARG, "weights is NULL");
    if (m == 0 || m > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "m %u not in 1..%d", m, RLR_MAX_M);
    *out_n = 0;
    if ((flags & RLR_SEARCH_F16) && s->d_rows16 == nullptr && s->n_rows)
        return fail(RLR_ERR_INVALID_ARG, "RLR_SEARCH_F16 but the store holds no f16 copy");
    if (int rc = ensure_device(s->device)) return rc;
    if (s->n_rows == 0) return RLR_OK; // :476-478
    CtxLease lease(s);
    if (int rc = lease.acquire()) return rc;
    rlr_ctx *c = lease.c;
    cudaStream_t st = c->stream;
    const uint64_t launches0 = c->launches;
    const uint32_t m_eff = static_cast<uint32_t>(std::min<uint64_t>(m, s->n_rows));
    Bm25WsGuard bm_guard;
    const uint32_t lex_limit = 5u * m;       // lexical_index.score(query, top_k * 5), :505 (m is `top_k` here)
    if (lat_eligible(s, flags)) {         // small store: one launch, no copies, no stream synchronisation
        uint32_t nl_lat = 0;
        bool fits = true;
        c->lat.d_lex_rows = nullptr; c->lat.d_lex_norm = nullptr;
        if (tq.ix) {
            if (int rc = lat_stage_query(c, query, dim, flags)) return rc;
            if (int rc = stage_text(c, tq, lex_limit, bm_guard, st, &nl_lat)) return rc;
            c->lat.d_lex_rows = c->d_lex_rows; c->lat.d_lex_norm = c->d_lex_norm;
            return lat_search(s, c, flags, w, nl_lat, m_eff, false, 0, 0.0f, out_rows, out_combined, out_emb, out_lex, out_n);
        }
        if (int rc = lat_stage_lex(c, lex_rows, lex_scores, n_lex, &nl_lat, &fits)) return rc;
        if (fits) {
            if (int rc = lat_stage_query(c, query, dim, flags)) return rc;
            return lat_search(s, c, flags, w, nl_lat, m_eff, false, 0, 0.0f, out_rows, out_combined, out_emb, out_lex, out_n);
        }
    }
    if (int rc = stage_query(c, query, dim, flags, st)) return rc;
    uint32_t nl = 0;
    if (tq.ix) { if (int rc = stage_text(c, tq, lex_limit, bm_guard, st, &nl)) return rc; }
    else if (int rc = stage_lex(c, lex_rows, lex_scores, n_lex, &nl, st)) return rc;
    const bool timed = flags & RLR_WANT_TIMINGS;
    if (timed) CU_TRY(cudaEventRecord(c->ev[0], st));
    if (int rc = enqueue_topm(c, c->d_query, w->embedding, w->lexical, c->d_lex_rows, c->d_lex_norm, nl, m_eff,
                              c->d_pool, c->d_pool_n, st, timed ? c->ev[1] : nullptr, s->use_half(flags)))
        return rc;
    if (timed) CU_TRY(cudaEventRecord(c->ev[2], st));
    CU_TRY(cudaMemcpyAsync(c->h_result_blk, c->d_pool_blk, 16 + m_eff * sizeof(rlr_cand), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    const uint32_t n = std::min(c->h_result_n[0], m_eff);
    unpack(c->h_result, n, out_rows, out_combined, out_emb, out_lex);
    *out_n = n;
    if (timed) { timings_from_events(c, false); g_timings.launches = static_cast<uint32_t>(c->launches - launches0); }
    return RLR_OK;
}

} // namespace

RLR_EXPORT int rlr_search_topm(rlr_store *s, const float *query, uint32_t dim, uint32_t flags,
                               const rlr_resolved_weights *w, const uint32_t *lex_rows, const float *lex_scores,
                               uint32_t n_lex, uint32_t m, uint32_t *out_rows, float *out_combined, float *out_emb,
                               float *out_lex, uint32_t *out_n)
{
    return search_topm_impl(s, query, dim, flags, w, lex_rows, lex_scores, n_lex, TextQuery(), m, out_rows, out_combined, out_emb, out_lex, out_n);
}

RLR_EXPORT int rlr_search_text_topm(rlr_store *s, rlr_bm25 *ix, const float *query, uint32_t dim, uint32_t flags,
                                    const rlr_resolved_weights *w, const uint32_t *query_terms, uint32_t n_terms, uint32_t m,
                                    uint32_t *out_rows, float *out_combined, float *out_emb, float *out_lex, uint32_t *out_n)
{
    if (!ix) return fail(RLR_ERR_INVALID_ARG, "the BM25 index is NULL");
    TextQuery tq;
    tq.ix = ix; tq.terms = query_terms; tq.n_terms = n_terms;
    return search_topm_impl(s, query, dim, flags, w, nullptr, nullptr, 0, tq, m, out_rows, out_combined, out_emb, out_lex, out_n);
}

RLR_EXPORT int rlr_embedding_candidates(rlr_store *s, const float *query, uint32_t dim, uint32_t flags, uint32_t count,
                                        uint32_t *out_rows, float *out_score, uint32_t *out_n)
{
    // :438-447 raw dot, sort desc, take(count): w_embed = 1 makes combined == emb exactly
    if (!out_n) return fail(RLR_ERR_INVALID_ARG, "out_n is NULL");
    *out_n = 0;
    if (count == 0) return RLR_OK;
    const rlr_resolved_weights w = {1.0f, 0.0f, 0.0f, 0.0f};
    return rlr_search_topm(s, query, dim, flags, &w, nullptr, nullptr, 0, count, out_rows, nullptr, out_score, nullptr, out_n);
}

RLR_EXPORT int rlr_mmr(rlr_store *s, const uint32_t *cand_rows, const float *relevance, uint32_t p, uint32_t top_k,
                       float lambda, uint32_t flags, uint32_t *out_sel_pos, uint32_t *out_n)
{
    if (int rc = check_store(s)) return rc;
    if (!out_sel_pos || !out_n) return fail(RLR_ERR_INVALID_ARG, "out_sel_pos/out_n is NULL");
    *out_n = 0;
    if (p == 0) return RLR_OK; // :773-775
    if (!cand_rows || !relevance) return fail(RLR_ERR_INVALID_ARG, "cand_rows/relevance is NULL");
    if (p > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "p %u exceeds %d", p, RLR_MAX_M);
    if (int rc = ensure_device(s->device)) return rc;
    for (uint32_t i = 0; i < p; ++i) {
        const uint64_t g = cand_rows[i];
        if (g < s->row_base || g - s->row_base >= s->n_rows) return fail(RLR_ERR_INVALID_ARG, "cand_rows[%u]=%llu not in this store", i, (unsigned long long)g);
    }
    CtxLease lease(s);
    if (int rc = lease.acquire()) return rc;
    rlr_ctx *c = lease.c;
    cudaStream_t st = c->stream;
    const uint64_t launches0 = c->launches;
    memcpy(c->h_u32, cand_rows, p * sizeof(uint32_t));
    c->h_u32[RLR_MAX_M] = p;
    memcpy(c->h_rel, relevance, p * sizeof(float));
    CU_TRY(cudaMemcpyAsync(c->d_rows_in, c->h_u32, p * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(c->d_p_in, c->h_u32 + RLR_MAX_M, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(c->d_rel_in, c->h_rel, p * sizeof(float), cudaMemcpyHostToDevice, st));
    const bool timed = flags & RLR_WANT_TIMINGS;
    if (timed) CU_TRY(cudaEventRecord(c->ev[2], st));
    rlr::MmrArgs a;
    memset(&a, 0, sizeof a);
    a.half = s->use_half(flags);
    a.d_emb = a.half ? s->d_rows16 : static_cast<const void *>(s->d_rows); a.pitch = a.half ? s->pitch16 : s->pitch; a.dim = s->dim;
    a.d_cands = nullptr; a.d_n = c->d_p_in; a.d_rows = c->d_rows_in; a.d_rel = c->d_rel_in;
    a.row_base = static_cast<uint32_t>(s->row_base); a.use_rows = 1;
    a.p_cap = p; a.top_k = top_k; a.lambda = lambda;
    a.d_tri = c->d_tri; a.d_sel_pos = c->d_sel_pos; a.d_sel_n = c->d_sel_n; a.d_result = nullptr;
    a.max_smem_optin = s->smem_optin;
    uint32_t l = 0;
    CU_TRY(rlr::mmr_launch(a, st, &l));
    c->launches += l;
    if (timed) CU_TRY(cudaEventRecord(c->ev[3], st));
    const uint32_t cap = std::min<uint32_t>(p, std::max<uint32_t>(top_k, 1));
    CU_TRY(cudaMemcpyAsync(c->h_u32, c->d_sel_pos, cap * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(c->h_u32 + RLR_MAX_M, c->d_sel_n, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    const uint32_t n = std::min(c->h_u32[RLR_MAX_M], cap);
    memcpy(out_sel_pos, c->h_u32, n * sizeof(uint32_t));
    *out_n = n;
    if (timed) {
        rlr_timings t = {0, 0, 0, 0, 0};
        cudaEventElapsedTime(&t.mmr_ms, c->ev[2], c->ev[3]);
        t.total_ms = t.mmr_ms;
        t.launches = static_cast<uint32_t>(c->launches - launches0);
        g_timings = t;
    }
    return RLR_OK;
}

namespace {
int search_mmr_impl(rlr_store *s, const float *query, uint32_t dim, uint32_t flags, uint32_t top_k,
                              float diversity_factor, const rlr_resolved_weights *w, const uint32_t *lex_rows,
                              const float *lex_scores, uint32_t n_lex, const TextQuery &tq, uint32_t *out_rows, float *out_score,
                              float *out_emb, float *out_lex, uint32_t *out_n)
{
    if (int rc = check_store(s)) return rc;
    if (!out_rows || !out_n) return fail(RLR_ERR_INVALID_ARG, "out_rows/out_n is NULL");
    if (!w) return fail(RLR_ERR_INVALID_ARG, "weights is NULL");
    // :725 f32::clamp (NaN stays NaN)
    float lambda = diversity_factor;
    if (lambda < 0.0f) lambda = 0.0f;
    if (lambda > 1.0f) lambda = 1.0f;
    if (lambda == 0.0f) { // :728-730 -> search(top_k), top_k.max(1) at :490
        const uint32_t m = std::max<uint32_t>(top_k, 1);
        return search_topm_impl(s, query, dim, flags, w, lex_rows, lex_scores, n_lex, tq, m, out_rows, out_score, out_emb,
                                out_lex, out_n);
    }
    *out_n = 0;
    const uint64_t pool = std::max<uint64_t>(3ull * top_k, static_cast<uint64_t>(top_k) + 10); // :734
    if (pool > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "candidate pool %llu exceeds %d (top_k %u)", (unsigned long long)pool, RLR_MAX_M, top_k);
    if (int rc = ensure_device(s->device)) return rc;
    if (s->n_rows == 0) return RLR_OK;
    CtxLease lease(s);
    if (int rc = lease.acquire()) return rc;
    rlr_ctx *c = lease.c;
    cudaStream_t st = c->stream;
    const uint64_t launches0 = c->launches;
    const uint32_t p = static_cast<uint32_t>(std::min<uint64_t>(pool, s->n_rows));
    Bm25WsGuard bm_guard;
    const uint32_t lex_limit = static_cast<uint32_t>(5u * pool);      // search(pool): lexical_index.score(query, pool * 5), :505
    if (lat_eligible(s, flags)) {         // small store: one launch (pool <= 32) or three, no copies, no stream synchronisation
        uint32_t nl_lat = 0;
        bool fits = true;
        c->lat.d_lex_rows = nullptr; c->lat.d_lex_norm = nullptr;
        if (tq.ix) {
            if (int rc = lat_stage_query(c, query, dim, flags)) return rc;
            if (int rc = stage_text(c, tq, lex_limit, bm_guard, st, &nl_lat)) return rc;
            c->lat.d_lex_rows = c->d_lex_rows; c->lat.d_lex_norm = c->d_lex_norm;
            return lat_search(s, c, flags, w, nl_lat, p, true, top_k, lambda, out_rows, out_score, out_emb, out_lex, out_n);
        }
        if (int rc = lat_stage_lex(c, lex_rows, lex_scores, n_lex, &nl_lat, &fits)) return rc;
        if (fits) {
            if (int rc = lat_stage_query(c, query, dim, flags)) return rc;
            return lat_search(s, c, flags, w, nl_lat, p, true, top_k, lambda, out_rows, out_score, out_emb, out_lex, out_n);
        }
    }
    if (int rc = stage_query(c, query, dim, flags, st)) return rc;
    uint32_t nl = 0;
    if (tq.ix) { if (int rc = stage_text(c, tq, lex_limit, bm_guard, st, &nl)) return rc; }
    else if (int rc = stage_lex(c, lex_rows, lex_scores, n_lex, &nl, st)) return rc;
    const bool timed = flags & RLR_WANT_TIMINGS;
    if (timed) CU_TRY(cudaEventRecord(c->ev[0], st));
    const bool half = s->use_half(flags);
    if (int rc = enqueue_topm(c, c->d_query, w->embedding, w->lexical, c->d_lex_rows, c->d_lex_norm, nl, p, c->d_pool,
                              c->d_pool_n, st, timed ? c->ev[1] : nullptr, half))
        return rc;
    if (timed) CU_TRY(cudaEventRecord(c->ev[2], st));
    rlr::MmrArgs a;
    memset(&a, 0, sizeof a);
    a.half = half;
    a.d_emb = half ? s->d_rows16 : static_cast<const void *>(s->d_rows); a.pitch = half ? s->pitch16 : s->pitch; a.dim = s->dim;
    a.d_cands = c->d_pool; a.d_n = c->d_pool_n; a.d_rows = nullptr; a.d_rel = nullptr;
    a.row_base = static_cast<uint32_t>(s->row_base); a.use_rows = 1;
    a.p_cap = p; a.top_k = top_k; a.lambda = lambda;
    a.d_tri = c->d_tri; a.d_sel_pos = c->d_sel_pos; a.d_sel_n = c->d_sel_n; a.d_result = c->d_result;
    a.max_smem_optin = s->smem_optin;
    uint32_t l = 0;
    CU_TRY(rlr::mmr_launch(a, st, &l));
    c->launches += l;
    if (timed) CU_TRY(cudaEventRecord(c->ev[3], st));
    const uint32_t cap = std::min<uint32_t>(p, std::max<uint32_t>(top_k, 1));
    CU_TRY(cudaMemcpyAsync(c->h_result_blk, c->d_result_blk, 16 + cap * sizeof(rlr_cand), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    const uint32_t n = std::min(c->h_result_n[0], cap);
    unpack(c->h_result, n, out_rows, out_score, out_emb, out_lex);
    *out_n = n;
    if (timed) { timings_from_events(c, true); g_timings.launches = static_cast<uint32_t>(c->launches - launches0); }
    return RLR_OK;
}
} // namespace

RLR_EXPORT int rlr_search_mmr(rlr_store *s, const float *query, uint32_t dim, uint32_t flags, uint32_t top_k,
                              float diversity_factor, const rlr_resolved_weights *w, const uint32_t *lex_rows,
                              const float *lex_scores, uint32_t n_lex, uint32_t *out_rows, float *out_score,
                              float *out_emb, float *out_lex, uint32_t *out_n)
{
    return search_mmr_impl(s, query, dim, flags, top_k, diversity_factor, w, lex_rows, lex_scores, n_lex, TextQuery(), out_rows, out_score,
                           out_emb, out_lex, out_n);
}

RLR_EXPORT int rlr_search_text_mmr(rlr_store *s, rlr_bm25 *ix, const float *query, uint32_t dim, uint32_t flags, uint32_t top_k,
                                   float diversity_factor, const rlr_resolved_weights *w, const uint32_t *query_terms,
                                   uint32_t n_terms, uint32_t *out_rows, float *out_score, float *out_emb, float *out_lex,
                                   uint32_t *out_n)
{
    if (!ix) return fail(RLR_ERR_INVALID_ARG, "the BM25 index is NULL");
    TextQuery tq;
    tq.ix = ix; tq.terms = query_terms; tq.n_terms = n_terms;
    return search_mmr_impl(s, query, dim, flags, top_k, diversity_factor, w, nullptr, nullptr, 0, tq, out_rows, out_score, out_emb, out_lex, out_n);
}

// ---------------------------------------------------------------------------------
// throughput mode: several queries per pass over the rows (query groups, scan_topm.cu)
// ---------------------------------------------------------------------------------
namespace {

// scan for nq queries (one launch) + per-query MMR, all on `st`.  Group q's query / lexical pairs / pool / MMR buffers
// are ctx q's; the launch workspace is ctx 0's.  d_result[q] / d_result_n[q]: where query q's result goes.
int enqueue_multi(rlr_store *s, rlr_ctx *const *cs, uint32_t nq, const float *const *d_query, const uint32_t *nl,
                  float w_e, float w_l, uint32_t top_k, float lambda, uint32_t flags, rlr_cand *const *d_result,
                  uint32_t *const *d_result_n, cudaStream_t st, cudaEvent_t ev_after_scan)
{
    const bool half = s->use_half(flags);
    const bool do_mmr = lambda != 0.0f;
    const uint64_t pool = do_mmr ? std::max<uint64_t>(3ull * top_k, static_cast<uint64_t>(top_k) + 10) : std::max<uint32_t>(top_k, 1);
    if (pool > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "candidate pool %llu exceeds %d (top_k %u)", (unsigned long long)pool, RLR_MAX_M, top_k);
    const uint32_t p = static_cast<uint32_t>(std::min<uint64_t>(pool, s->n_rows));
    rlr_ctx *c0 = cs[0];
    rlr::ScanArgs a;
    memset(&a, 0, sizeof a);
    rlr::scan_plan(s->sm_count, s->smem_optin, static_cast<uint32_t>(s->n_rows), half ? s->pitch16 : s->pitch, half, &a, nq);
    if (a.grid <= 0) return fail(RLR_ERR_UNSUPPORTED, "%u query groups do not fit this store's row size in shared memory", nq);
    a.tmap = half ? &s->tmap16 : &s->tmap;
    a.n_rows = static_cast<uint32_t>(s->n_rows);
    a.row_base = static_cast<uint32_t>(s->row_base);
    a.pitch = half ? s->pitch16 : s->pitch;
    a.w_embed = w_e; a.w_lex = w_l;
    a.m = p;
    a.d_lists = c0->d_lists; a.d_counts = c0->d_counts; a.d_ticket = c0->d_ticket; a.d_pub = c0->d_pub;
    for (uint32_t q = 0; q < nq; ++q) {
        rlr::ScanGroupIO &g = a.groups.g[q];
        g.query = d_query[q];
        g.lex_rows = cs[q]->d_lex_rows; g.lex_norm = cs[q]->d_lex_norm; g.n_lex = nl ? nl[q] : 0;
        g.out = do_mmr ? cs[q]->d_pool : d_result[q];
        g.out_n = do_mmr ? cs[q]->d_pool_n : d_result_n[q];
    }
    if (nq == 1) {          // scan_launch reads the single-query fields
        a.d_query = a.groups.g[0].query; a.d_lex_rows = a.groups.g[0].lex_rows; a.d_lex_norm = a.groups.g[0].lex_norm;
        a.n_lex = a.groups.g[0].n_lex; a.d_out = a.groups.g[0].out; a.d_out_n = a.groups.g[0].out_n;
    }
    CU_TRY(rlr::scan_launch(a, st));
    ++c0->launches;
    if (ev_after_scan) CU_TRY(cudaEventRecord(ev_after_scan, st));
    if (!do_mmr) return RLR_OK;
    for (uint32_t q = 0; q < nq; ++q) {
        rlr_ctx *c = cs[q];
        rlr::MmrArgs ma;
        memset(&ma, 0, sizeof ma);
        ma.half = half;
        ma.d_emb = half ? s->d_rows16 : static_cast<const void *>(s->d_rows); ma.pitch = half ? s->pitch16 : s->pitch; ma.dim = s->dim;
        ma.d_cands = c->d_pool; ma.d_n = c->d_pool_n;
        ma.row_base = static_cast<uint32_t>(s->row_base); ma.use_rows = 1;
        ma.p_cap = p; ma.top_k = top_k; ma.lambda = lambda;
        ma.d_tri = c->d_tri; ma.d_sel_pos = c->d_sel_pos; ma.d_sel_n = d_result_n[q]; ma.d_result = d_result[q];
        ma.max_smem_optin = s->smem_optin;
        uint32_t l = 0;
        CU_TRY(rlr::mmr_launch(ma, st, &l));
        c0->launches += l;
    }
    return RLR_OK;
}

} // namespace

RLR_EXPORT int rlr_search_mmr_multi(rlr_store *s, const float *queries, uint32_t nq, uint32_t dim, uint32_t flags, uint32_t top_k,
                                    float diversity_factor, const rlr_resolved_weights *w, const uint32_t *const *lex_rows,
                                    const float *const *lex_scores, const uint32_t *n_lex, uint32_t *out_rows, float *out_score,
                                    float *out_emb, float *out_lex, uint32_t *out_n)
{
    if (int rc = check_store(s)) return rc;
    if (!out_rows || !out_n) return fail(RLR_ERR_INVALID_ARG, "out_rows/out_n is NULL");
    if (!w) return fail(RLR_ERR_INVALID_ARG, "weights is NULL");
    if (nq == 0) return RLR_OK;
    if (nq > RLR_MAX_MULTI) return fail(RLR_ERR_UNSUPPORTED, "nq %u exceeds RLR_MAX_MULTI (%d)", nq, RLR_MAX_MULTI);
    if (!queries) return fail(RLR_ERR_INVALID_ARG, "queries is NULL");
    const uint32_t cap = std::max<uint32_t>(top_k, 1);
    if (nq == 1)
        return rlr_search_mmr(s, queries, dim, flags, top_k, diversity_factor, w, lex_rows ? lex_rows[0] : nullptr,
                              lex_scores ? lex_scores[0] : nullptr, n_lex ? n_lex[0] : 0, out_rows, out_score, out_emb, out_lex, out_n);
    for (uint32_t q = 0; q < nq; ++q) out_n[q] = 0;
    float lambda = diversity_factor;                    // :725 f32::clamp (NaN stays NaN)
    if (lambda < 0.0f) lambda = 0.0f;
    if (lambda > 1.0f) lambda = 1.0f;
    if ((flags & RLR_SEARCH_F16) && s->d_rows16 == nullptr && s->n_rows)
        return fail(RLR_ERR_INVALID_ARG, "RLR_SEARCH_F16 but the store holds no f16 copy");
    if (int rc = ensure_device(s->device)) return rc;
    if (s->n_rows == 0) return RLR_OK;
    // one pooled ctx per query: its query, lexical pairs, pool, MMR buffers and pinned staging
    std::vector<std::unique_ptr<CtxLease>> leases;
    rlr_ctx *cs[RLR_MAX_MULTI];
    for (uint32_t q = 0; q < nq; ++q) {
        leases.emplace_back(new CtxLease(s));
        if (int rc = leases.back()->acquire()) return rc;
        cs[q] = leases.back()->c;
    }
    cudaStream_t st = cs[0]->stream;
    const float *dq[RLR_MAX_MULTI];
    uint32_t nl[RLR_MAX_MULTI];
    rlr_cand *dres[RLR_MAX_MULTI];
    uint32_t *dres_n[RLR_MAX_MULTI];
    for (uint32_t q = 0; q < nq; ++q) {
        if (int rc = stage_query(cs[q], queries + static_cast<size_t>(q) * dim, dim, flags, st)) return rc;
        nl[q] = 0;
        if (int rc = stage_lex(cs[q], lex_rows ? lex_rows[q] : nullptr, lex_scores ? lex_scores[q] : nullptr, n_lex ? n_lex[q] : 0, &nl[q], st)) return rc;
        dq[q] = cs[q]->d_query;
        dres[q] = cs[q]->d_result; dres_n[q] = cs[q]->d_sel_n;
    }
    const bool timed = flags & RLR_WANT_TIMINGS;
    const uint64_t launches0 = cs[0]->launches;
    if (timed) CU_TRY(cudaEventRecord(cs[0]->ev[0], st));
    if (int rc = enqueue_multi(s, cs, nq, dq, nl, w->embedding, w->lexical, top_k, lambda, flags, dres, dres_n, st, timed ? cs[0]->ev[1] : nullptr)) return rc;
    if (timed) CU_TRY(cudaEventRecord(cs[0]->ev[3], st));
    const uint64_t pool = lambda != 0.0f ? std::max<uint64_t>(3ull * top_k, static_cast<uint64_t>(top_k) + 10) : cap;
    const uint32_t p = static_cast<uint32_t>(std::min<uint64_t>(pool, s->n_rows));
    const uint32_t n_cap = std::min<uint32_t>(p, cap);
    for (uint32_t q = 0; q < nq; ++q)
        CU_TRY(cudaMemcpyAsync(cs[q]->h_result_blk, cs[q]->d_result_blk, 16 + n_cap * sizeof(rlr_cand), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    for (uint32_t q = 0; q < nq; ++q) {
        const uint32_t n = std::min(cs[q]->h_result_n[0], n_cap);
        unpack(cs[q]->h_result, n, out_rows + static_cast<size_t>(q) * cap, out_score ? out_score + static_cast<size_t>(q) * cap : nullptr,
               out_emb ? out_emb + static_cast<size_t>(q) * cap : nullptr, out_lex ? out_lex + static_cast<size_t>(q) * cap : nullptr);
        out_n[q] = n;
    }
    if (timed) {
        rlr_timings t = {0, 0, 0, 0, 0};
        cudaEventElapsedTime(&t.scan_ms, cs[0]->ev[0], cs[0]->ev[1]);
        cudaEventElapsedTime(&t.mmr_ms, cs[0]->ev[1], cs[0]->ev[3]);
        cudaEventElapsedTime(&t.total_ms, cs[0]->ev[0], cs[0]->ev[3]);
        cudaGetLastError();
        t.launches = static_cast<uint32_t>(cs[0]->launches - launches0);
        g_timings = t;
    }
    return RLR_OK;
}

RLR_EXPORT int rlr_search_mmr_multi_async(rlr_ctx *const *ctxs, uint32_t nq, const void *const *d_queries, uint32_t top_k,
                                          float diversity_factor, float w_embed, float w_lex, void *const *d_results,
                                          void *const *d_result_ns, void *stream)
{
    if (!ctxs || !d_queries || !d_results || !d_result_ns) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    if (nq == 0 || nq > RLR_MAX_MULTI) return fail(RLR_ERR_UNSUPPORTED, "nq %u not in 1..%d", nq, RLR_MAX_MULTI);
    for (uint32_t q = 0; q < nq; ++q) {
        if (!ctxs[q] || !d_queries[q] || !d_results[q] || !d_result_ns[q]) return fail(RLR_ERR_INVALID_ARG, "NULL argument for query %u", q);
        if (ctxs[q]->s != ctxs[0]->s) return fail(RLR_ERR_INVALID_ARG, "the ctxs must belong to one store");
    }
    rlr_store *s = ctxs[0]->s;
    CU_TRY(cudaSetDevice(s->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float lambda = diversity_factor;
    if (lambda < 0.0f) lambda = 0.0f;
    if (lambda > 1.0f) lambda = 1.0f;
    if (s->n_rows == 0) {
        for (uint32_t q = 0; q < nq; ++q) CU_TRY(cudaMemsetAsync(d_result_ns[q], 0, sizeof(uint32_t), st));
        return RLR_OK;
    }
    const float *dq[RLR_MAX_MULTI];
    rlr_cand *dres[RLR_MAX_MULTI];
    uint32_t *dres_n[RLR_MAX_MULTI];
    for (uint32_t q = 0; q < nq; ++q) {
        dq[q] = static_cast<const float *>(d_queries[q]);
        dres[q] = static_cast<rlr_cand *>(d_results[q]);
        dres_n[q] = static_cast<uint32_t *>(d_result_ns[q]);
    }
    return enqueue_multi(s, ctxs, nq, dq, nullptr, w_embed, w_lex, top_k, lambda, ctxs[0]->search_flags, dres, dres_n, st, nullptr);
}

// ---------------------------------------------------------------------------------
// batched queries on the tensor cores
// ---------------------------------------------------------------------------------
namespace {

struct BatchBufs {                 // views into the ctx's batch workspace
    int prec = 0;                  // rlr::kPrecF16 / kPrecBF16 / kPrecTF32: what the operand tiles hold
    const CUtensorMap *tmapA = nullptr;
    uint32_t pitch_op = 0;         // operand elements per row (store pitch of that precision)
    float *d_q32 = nullptr;
    void *d_q16 = nullptr;
    float *d_tau = nullptr;
    unsigned long long *d_state = nullptr, *d_app = nullptr;
    uint32_t *d_state_cnt = nullptr, *d_app_cnt = nullptr, *d_overflow = nullptr;
};

constexpr uint32_t kBatchCap = 1024;      // appended candidates per query per phase
constexpr uint32_t kBatchQTile = 256;     // queries per MMA tile (UMMA N)
constexpr uint32_t kBatchRTile = 128;     // store rows per MMA tile (UMMA M)

// GEMM + prune over row tiles [t0, t1); on candidate-list overflow the phase is split and retried
// (the running top-m is only modified by the prune, so a failed GEMM pass leaves it intact).
int batch_phase(rlr_store *s, const CUtensorMap *tmapQ, BatchBufs &b, uint32_t nq, uint32_t nq_pad, uint32_t m,
                uint32_t t0, uint32_t t1, cudaStream_t st, uint32_t *launches, uint32_t *h_flag /* pinned */,
                bool checked)
{
    // tmapQ[0]: box {64, 256} for the 1-CTA kernel; tmapQ[1]: box {64, 128} for the cta_group::2 kernel,
    // which needs an even first tile (tiles are handed out in 256-row pairs)
    static const bool two_cta = getenv("RLR_BATCH_1CTA") == nullptr;
    const uint32_t n_tiles_all = static_cast<uint32_t>((s->n_rows + kBatchRTile - 1) / kBatchRTile);
    // first phase of a batch: tau is -inf everywhere and the phase holds <= cap rows, so every row is kept:
    // the kernel writes row r to slot r of every list (no filter, no atomics) and the counts are set here
    const bool dense = two_cta && t0 == 0 && (t1 - t0) * kBatchRTile <= kBatchCap && (t1 % 2 == 0 || t1 == n_tiles_all);
    if (dense) {
        CU_TRY(rlr::batch_gemm2_launch(b.tmapA, tmapQ + 1, s->sm_count, static_cast<uint32_t>(s->n_rows),
                                       static_cast<uint32_t>(s->row_base), t0, t1, nq_pad, b.pitch_op, b.prec, b.d_tau, b.d_app,
                                       b.d_app_cnt, kBatchCap, b.d_overflow, 1, st));
        const uint32_t rows_in_phase = static_cast<uint32_t>(std::min<uint64_t>(s->n_rows, static_cast<uint64_t>(t1) * kBatchRTile));
        CU_TRY(rlr::batch_set_cnt_launch(b.d_app_cnt, nq, rows_in_phase, st));
        *launches += 2;
        CU_TRY(rlr::batch_prune_launch(b.d_state, b.d_state_cnt, m, b.d_app, b.d_app_cnt, kBatchCap, b.d_tau, nq, st));
        ++*launches;
        return RLR_OK;
    }
    if (two_cta && (t0 % 2 == 0) && (t1 % 2 == 0 || t1 == n_tiles_all) && (t1 - t0) >= 2)
        CU_TRY(rlr::batch_gemm2_launch(b.tmapA, tmapQ + 1, s->sm_count, static_cast<uint32_t>(s->n_rows),
                                       static_cast<uint32_t>(s->row_base), t0, t1, nq_pad, b.pitch_op, b.prec, b.d_tau, b.d_app,
                                       b.d_app_cnt, kBatchCap, b.d_overflow, 0, st));
    else
    CU_TRY(rlr::batch_gemm_launch(b.tmapA, tmapQ, s->sm_count, static_cast<uint32_t>(s->n_rows),
                                  static_cast<uint32_t>(s->row_base), t0, t1, nq_pad, b.pitch_op, b.prec, b.d_tau, b.d_app,
                                  b.d_app_cnt, kBatchCap, b.d_overflow, st));
    ++*launches;
    if (checked) {      // safe mode: look at the overflow flag after every phase (a host round trip each)
        CU_TRY(cudaMemcpyAsync(h_flag, b.d_overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
    }
    if (checked && *h_flag) {
        if (t1 - t0 <= 1) return fail(RLR_ERR_CUDA, "batch candidate list overflow on a single tile (internal error)");
        CU_TRY(cudaMemsetAsync(b.d_app_cnt, 0, static_cast<size_t>(nq_pad) * 32 * sizeof(uint32_t), st));
        CU_TRY(cudaMemsetAsync(b.d_overflow, 0, sizeof(uint32_t), st));
        const uint32_t mid = t0 + (t1 - t0) / 2;
        if (int rc = batch_phase(s, tmapQ, b, nq, nq_pad, m, t0, mid, st, launches, h_flag, true)) return rc;
        return batch_phase(s, tmapQ, b, nq, nq_pad, m, mid, t1, st, launches, h_flag, true);
    }
    CU_TRY(rlr::batch_prune_launch(b.d_state, b.d_state_cnt, m, b.d_app, b.d_app_cnt, kBatchCap, b.d_tau, nq, st));
    ++*launches;
    return RLR_OK;
}

} // namespace

namespace {
// Shared body of rlr_search_batch (host outputs) and rlr_search_batch_device (d_keys/d_cnt: the
// per-query rank-ordered key lists stay in HBM, padded with key 0, for the multi-GPU merge).
int search_batch_core(rlr_store *s, const float *queries, uint32_t n_queries, uint32_t dim, uint32_t flags,
                      uint32_t m, uint32_t *out_rows, float *out_scores, uint32_t *out_n,
                      unsigned long long *d_keys, uint32_t *d_cnt, cudaStream_t user_stream)
{
    const bool to_device = d_keys != nullptr;
    if (int rc = check_store(s)) return rc;
    if (!to_device && (!out_rows || !out_n)) return fail(RLR_ERR_INVALID_ARG, "out_rows/out_n is NULL");
    if (n_queries == 0) return RLR_OK;
    if (!queries) return fail(RLR_ERR_INVALID_ARG, "queries is NULL");
    if (m == 0 || m > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "m %u not in 1..%d", m, RLR_MAX_M);
    if (dim != s->dim) return fail(RLR_ERR_DIM_MISMATCH, "queries have %u dims, store has %u", dim, s->dim);
    if (n_queries > 4096) return fail(RLR_ERR_UNSUPPORTED, "n_queries %u exceeds 4096 per call", n_queries);
    if (out_n) for (uint32_t q = 0; q < n_queries; ++q) out_n[q] = 0;
    if (int rc = ensure_device(s->device)) return rc;
    if (s->n_rows == 0) {
        if (to_device) {
            CU_TRY(cudaMemsetAsync(d_keys, 0, static_cast<size_t>(n_queries) * m * 8, user_stream));
            if (d_cnt) CU_TRY(cudaMemsetAsync(d_cnt, 0, n_queries * sizeof(uint32_t), user_stream));
        }
        return RLR_OK;
    }
    // operand precision: what the caller asked for, else binary16 when the store keeps that copy, else tf32 straight
    // over the f32 rows (no second copy of the store needed)
    int prec;
    if (flags & RLR_BATCH_TF32) prec = rlr::kPrecTF32;
    else if (flags & RLR_BATCH_BF16) prec = rlr::kPrecBF16;
    else if (flags & RLR_BATCH_F16) prec = rlr::kPrecF16;
    else prec = s->d_rows16 ? rlr::kPrecF16 : rlr::kPrecTF32;
    if (prec == rlr::kPrecF16 && s->d_rows16 == nullptr)
        return fail(RLR_ERR_INVALID_ARG, "RLR_BATCH_F16 needs the binary16 store copy (RLR_STORE_KEEP_F16 / RLR_STORE_F16_ONLY)");
    if (prec == rlr::kPrecBF16 && s->d_rows_bf16 == nullptr)
        return fail(RLR_ERR_INVALID_ARG, "RLR_BATCH_BF16 needs the bfloat16 store copy (RLR_STORE_KEEP_BF16)");
    if (prec == rlr::kPrecTF32 && s->d_rows == nullptr)
        return fail(RLR_ERR_INVALID_ARG, "RLR_BATCH_TF32 needs the f32 rows (the store is RLR_STORE_F16_ONLY)");
    {
        // per device, remembered only when it succeeded: a failed configure is retried (and reported) by the next call
        static std::mutex mu;
        static bool configured[64] = {false};
        std::lock_guard<std::mutex> lk(mu);
        if (!configured[s->device]) {
            CU_TRY(rlr::batch_configure(s->smem_optin));
            configured[s->device] = true;
        }
    }
    CtxLease lease(s);
    if (int rc = lease.acquire()) return rc;
    rlr_ctx *c = lease.c;
    cudaStream_t st = c->stream;
    const uint32_t nq = n_queries, nq_pad = (nq + kBatchQTile - 1) / kBatchQTile * kBatchQTile;
    const uint32_t m_eff = static_cast<uint32_t>(std::min<uint64_t>(m, s->n_rows));

    // The queries go straight from the caller's buffer to pinned memory and to the device; normalize (:494) and the
    // NaN/Inf check run there (same sequential arithmetic, same bits as the host normalize; a 1024 x 1024 batch
    // cost ~3 ms of single-threaded host work before the first kernel could start).
    const size_t q_floats = static_cast<size_t>(nq) * dim;
    BatchBufs b;
    b.prec = prec;
    b.tmapA = prec == rlr::kPrecTF32 ? &s->tmap : prec == rlr::kPrecBF16 ? &s->tmap_bf16 : &s->tmap16;
    b.pitch_op = prec == rlr::kPrecTF32 ? s->pitch : s->pitch16;
    const uint32_t op_esz = prec == rlr::kPrecTF32 ? 4u : 2u;
    {
        // one workspace allocation per ctx, grown on demand (a batch call is a few ms: no per-call cudaMalloc)
        auto up = [](size_t x) { return (x + 255) & ~static_cast<size_t>(255); };
        const size_t sz_q32 = up(q_floats * sizeof(float)), sz_q16 = up(static_cast<size_t>(nq_pad) * b.pitch_op * op_esz);
        const size_t sz_tau = up(nq_pad * sizeof(float)), sz_state = up(static_cast<size_t>(nq_pad) * m_eff * 8);
        const size_t sz_app = up(static_cast<size_t>(nq_pad) * kBatchCap * 8), sz_cnt = up(nq_pad * sizeof(uint32_t));
        const size_t sz_acnt = up(static_cast<size_t>(nq_pad) * 32 * sizeof(uint32_t));   // one 128-byte line per counter
        const size_t total = sz_q32 + sz_q16 + sz_tau + sz_state + sz_app + sz_cnt + sz_acnt + 256;
        if (total > c->batch_bytes) {
            cudaFree(c->batch_mem); c->batch_mem = nullptr; c->batch_bytes = 0;
            CU_TRY(cudaMalloc(&c->batch_mem, total));
            c->batch_bytes = total;
        }
        uint8_t *p = static_cast<uint8_t *>(c->batch_mem);
        b.d_q32 = reinterpret_cast<float *>(p); p += sz_q32;
        b.d_q16 = p; p += sz_q16;
        b.d_tau = reinterpret_cast<float *>(p); p += sz_tau;
        b.d_state = reinterpret_cast<unsigned long long *>(p); p += sz_state;
        b.d_app = reinterpret_cast<unsigned long long *>(p); p += sz_app;
        b.d_state_cnt = reinterpret_cast<uint32_t *>(p); p += sz_cnt;
        b.d_app_cnt = reinterpret_cast<uint32_t *>(p); p += sz_acnt;
        b.d_overflow = reinterpret_cast<uint32_t *>(p);
        if (q_floats * sizeof(float) > c->h_batch_q_bytes) {
            cudaFreeHost(c->h_batch_q); c->h_batch_q = nullptr; c->h_batch_q_bytes = 0;
            CU_TRY(cudaMallocHost(&c->h_batch_q, q_floats * sizeof(float)));
            c->h_batch_q_bytes = q_floats * sizeof(float);
        }
        const size_t st_bytes = static_cast<size_t>(nq) * m_eff * 8 + nq * sizeof(uint32_t);
        if (st_bytes > c->h_batch_state_bytes) {
            cudaFreeHost(c->h_batch_state); c->h_batch_state = nullptr; c->h_batch_state_bytes = 0;
            CU_TRY(cudaMallocHost(&c->h_batch_state, st_bytes));
            c->h_batch_state_bytes = st_bytes;
        }
        memcpy(c->h_batch_q, queries, q_floats * sizeof(float));
    }
    CUtensorMap tmapQ[2];   // [0] box {64, 256} (1-CTA kernel), [1] box {64, 128} (cta_group::2 kernel)
    {
        PFN_encodeTiled enc = get_encode();
        if (!enc) return fail(RLR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
        const cuuint64_t gdim[2] = {b.pitch_op, nq_pad};
        const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(b.pitch_op) * op_esz};
        const cuuint32_t estr[2] = {1, 1};
        const CUtensorMapDataType qdt = prec == rlr::kPrecTF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                      : prec == rlr::kPrecBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
        for (int v = 0; v < 2; ++v) {
            const cuuint32_t box[2] = {128u / op_esz, v == 0 ? kBatchQTile : kBatchQTile / 2};
            CUresult r = enc(&tmapQ[v], qdt, 2, b.d_q16, gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(RLR_ERR_CUDA, "cuTensorMapEncodeTiled (queries) failed with CUresult %d", static_cast<int>(r));
        }
    }
    const bool timed = flags & RLR_WANT_TIMINGS;
    uint32_t launches = 0;
    CU_TRY(cudaMemcpyAsync(b.d_q32, c->h_batch_q, q_floats * sizeof(float), cudaMemcpyHostToDevice, st));
    if (timed) CU_TRY(cudaEventRecord(c->ev[0], st));
    CU_TRY(rlr::batch_init_launch(b.d_tau, b.d_state_cnt, b.d_app_cnt, nq, nq_pad, b.d_overflow, st));
    if (!(flags & RLR_QUERY_PRENORMALIZED)) { CU_TRY(rlr::normalize_rows_launch(b.d_q32, dim, dim, nq, st)); ++launches; }
    CU_TRY(rlr::batch_queries_to_operand_launch(b.d_q32, dim, b.d_q16, b.pitch_op, prec, nq, nq_pad, b.d_overflow + 1, st));
    launches += 2;
    bool nonfinite = false;
    // Geometrically growing phases.  tau is frozen during a phase, so a phase over rows [a, g*a)
    // lets ~m*(g-1) rows per query through; g is chosen to keep that near 70 % of the list capacity.
    // Attempt 0 enqueues all phases back to back and looks at the overflow flag once at the end;
    // if any list overflowed (skewed data), attempt 1 redoes the batch checking (and splitting) per phase.
    const uint32_t n_tiles = static_cast<uint32_t>((s->n_rows + kBatchRTile - 1) / kBatchRTile);
    const uint32_t growth = std::min<uint32_t>(16, std::max<uint32_t>(2, 1 + (7 * kBatchCap) / (10 * m_eff)));
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (attempt == 1) {
            CU_TRY(rlr::batch_init_launch(b.d_tau, b.d_state_cnt, b.d_app_cnt, nq, nq_pad, b.d_overflow, st));
            ++launches;
        }
        uint32_t t0 = 0, span = kBatchCap / kBatchRTile;          // first phase: cap rows, every row is kept
        while (t0 < n_tiles) {
            const uint32_t t1 = std::min(n_tiles, t0 + span);
            if (int rc = batch_phase(s, tmapQ, b, nq, nq_pad, m_eff, t0, t1, st, &launches, c->h_u32, attempt == 1)) return rc;
            t0 = t1;
            span = ((attempt == 0 ? growth - 1 : 1) * t0 + 1) & ~1u;   // rows [t0, g*t0); even tile counts
        }
        if (attempt == 0) {
            CU_TRY(cudaMemcpyAsync(c->h_u32, b.d_overflow, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
            nonfinite = c->h_u32[1] != 0;
            if (c->h_u32[0] == 0 || nonfinite) break;
        }
    }
    if (nonfinite) return fail(RLR_ERR_NONFINITE, "the query batch contains NaN/Inf");
    if (timed) CU_TRY(cudaEventRecord(c->ev[1], st));
    if (flags & RLR_BATCH_EXACT_RESCORE) {
        const bool half = s->d_rows == nullptr;
        CU_TRY(rlr::batch_rescore_launch(half ? s->d_rows16 : static_cast<const void *>(s->d_rows), half,
                                         half ? s->pitch16 : s->pitch, s->dim, static_cast<uint32_t>(s->row_base), b.d_q32,
                                         b.d_state, b.d_state_cnt, m_eff, nq, st));
        CU_TRY(rlr::batch_prune_launch(b.d_state, b.d_state_cnt, m_eff, b.d_app, b.d_app_cnt, kBatchCap, b.d_tau, nq, st));
        launches += 2;
    }
    if (timed) CU_TRY(cudaEventRecord(c->ev[2], st));
    if (to_device) {
        // rank-ordered keys stay on the device: [n_queries][m], rows beyond min(m, n_rows) are key 0.
        // The caller may still have reads of d_keys / d_cnt queued on its stream: order our writes behind them.
        CU_TRY(cudaEventRecord(c->ev[4], user_stream));
        CU_TRY(cudaStreamWaitEvent(st, c->ev[4], 0));
        if (m_eff != m) CU_TRY(cudaMemsetAsync(d_keys, 0, static_cast<size_t>(nq) * m * 8, st));
        CU_TRY(cudaMemcpy2DAsync(d_keys, static_cast<size_t>(m) * 8, b.d_state, static_cast<size_t>(m_eff) * 8,
                                 static_cast<size_t>(m_eff) * 8, nq, cudaMemcpyDeviceToDevice, st));
        if (d_cnt) CU_TRY(cudaMemcpyAsync(d_cnt, b.d_state_cnt, nq * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        CU_TRY(cudaEventRecord(c->ev[3], st));
        CU_TRY(cudaStreamWaitEvent(user_stream, c->ev[3], 0));   // the caller's stream sees the lists
        CU_TRY(cudaStreamSynchronize(st));                        // the ctx goes back to the pool idle
        c->launches += launches;
        if (timed) {
            rlr_timings t = {0, 0, 0, 0, 0};
            cudaEventElapsedTime(&t.scan_ms, c->ev[0], c->ev[1]);
            cudaEventElapsedTime(&t.merge_ms, c->ev[1], c->ev[2]);
            cudaEventElapsedTime(&t.total_ms, c->ev[0], c->ev[2]);
            cudaGetLastError();
            t.launches = launches;
            g_timings = t;
        }
        return RLR_OK;
    }
    unsigned long long *h_state = c->h_batch_state;
    uint32_t *h_cnt = reinterpret_cast<uint32_t *>(h_state + static_cast<size_t>(nq) * m_eff);
    CU_TRY(cudaMemcpyAsync(h_state, b.d_state, static_cast<size_t>(nq) * m_eff * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(h_cnt, b.d_state_cnt, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    for (uint32_t q = 0; q < nq; ++q) {
        const uint32_t n = std::min(h_cnt[q], m_eff);
        out_n[q] = n;
        for (uint32_t i = 0; i < n; ++i) {
            const unsigned long long k = h_state[static_cast<size_t>(q) * m_eff + i];
            out_rows[static_cast<size_t>(q) * m + i] = rlr::key_row(k);
            if (out_scores) {
                const uint32_t bits = rlr::bits_from_ord(static_cast<uint32_t>(k >> 32));
                memcpy(&out_scores[static_cast<size_t>(q) * m + i], &bits, 4);
            }
        }
    }
    c->launches += launches;
    if (timed) {
        rlr_timings t = {0, 0, 0, 0, 0};
        cudaEventElapsedTime(&t.scan_ms, c->ev[0], c->ev[1]);      // contraction + per-phase prunes
        cudaEventElapsedTime(&t.merge_ms, c->ev[1], c->ev[2]);     // exact re-score (if requested)
        cudaEventElapsedTime(&t.total_ms, c->ev[0], c->ev[2]);
        cudaGetLastError();
        t.launches = launches;
        g_timings = t;
    }
    return RLR_OK;
}
} // namespace

RLR_EXPORT int rlr_search_batch(rlr_store *s, const float *queries, uint32_t n_queries, uint32_t dim, uint32_t flags,
                                uint32_t m, uint32_t *out_rows, float *out_scores, uint32_t *out_n)
{
    return search_batch_core(s, queries, n_queries, dim, flags, m, out_rows, out_scores, out_n, nullptr, nullptr, nullptr);
}

RLR_EXPORT int rlr_search_batch_device(rlr_store *s, const float *queries, uint32_t n_queries, uint32_t dim,
                                       uint32_t flags, uint32_t m, void *d_keys, void *d_cnt, void *stream)
{
    if (!d_keys) return fail(RLR_ERR_INVALID_ARG, "d_keys is NULL");
    return search_batch_core(s, queries, n_queries, dim, flags, m, nullptr, nullptr, nullptr,
                             static_cast<unsigned long long *>(d_keys), static_cast<uint32_t *>(d_cnt),
                             static_cast<cudaStream_t>(stream));
}

RLR_EXPORT int rlr_batch_merge_async(rlr_store *s, const void *d_lists, uint32_t n_lists, uint32_t n_queries, uint32_t m,
                                     void *d_out_keys, void *d_out_cnt, void *stream)
{
    if (int rc = check_store(s)) return rc;
    if (!d_lists || !d_out_keys) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    if (m == 0 || m > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "m %u not in 1..%d", m, RLR_MAX_M);
    if (n_lists == 0 || n_lists > rlr::kMaxPeers) return fail(RLR_ERR_UNSUPPORTED, "n_lists %u not in 1..%d", n_lists, rlr::kMaxPeers);
    if (n_queries == 0) return RLR_OK;
    if (int rc = ensure_device(s->device)) return rc;
    CU_TRY(rlr::batch_merge_launch(static_cast<const unsigned long long *>(d_lists), n_lists, n_queries, m,
                                   static_cast<unsigned long long *>(d_out_keys), static_cast<uint32_t *>(d_out_cnt),
                                   static_cast<cudaStream_t>(stream)));
    return RLR_OK;
}

RLR_EXPORT int rlr_last_timings(rlr_timings *out)
{
    if (!out) return fail(RLR_ERR_INVALID_ARG, "out is NULL");
    *out = g_timings;
    return RLR_OK;
}

// ---------------------------------------------------------------------------------
// device-level building blocks
// ---------------------------------------------------------------------------------
RLR_EXPORT int rlr_ctx_create(rlr_store *s, rlr_ctx **out)
{
    if (int rc = check_store(s)) return rc;
    if (!out) return fail(RLR_ERR_INVALID_ARG, "out is NULL");
    if (int rc = ensure_device(s->device)) return rc;
    return ctx_new(s, out);
}

RLR_EXPORT int rlr_ctx_destroy(rlr_ctx *c)
{
    if (!c) return RLR_OK;
    cudaSetDevice(c->s->device);
    ctx_free(c);
    return RLR_OK;
}

RLR_EXPORT int rlr_ctx_set_flags(rlr_ctx *c, uint32_t search_flags)
{
    if (!c) return fail(RLR_ERR_INVALID_ARG, "ctx is NULL");
    if ((search_flags & RLR_SEARCH_F16) && c->s->d_rows16 == nullptr)
        return fail(RLR_ERR_INVALID_ARG, "RLR_SEARCH_F16 but the store holds no f16 copy");
    c->search_flags = search_flags;
    return RLR_OK;
}

RLR_EXPORT int rlr_ctx_launch_count(const rlr_ctx *c, uint64_t *out)
{
    if (!c || !out) return fail(RLR_ERR_INVALID_ARG, "ctx/out is NULL");
    *out = c->launches;
    return RLR_OK;
}

RLR_EXPORT int rlr_topm_async(rlr_ctx *c, const void *d_query, float w_embed, float w_lex, const void *d_lex_rows,
                              const void *d_lex_norm, uint32_t n_lex, uint32_t m, void *d_out, void *d_out_n,
                              void *stream)
{
    if (!c || !d_query || !d_out || !d_out_n) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    if (m == 0 || m > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "m %u not in 1..%d", m, RLR_MAX_M);
    rlr_store *s = c->s;
    CU_TRY(cudaSetDevice(s->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (s->n_rows == 0) {
        CU_TRY(cudaMemsetAsync(d_out, 0, m * sizeof(rlr_cand), st));
        CU_TRY(cudaMemsetAsync(d_out_n, 0, sizeof(uint32_t), st));
        return RLR_OK;
    }
    // lists shorter than m are zero-key padded by the kernels, so m need not be clamped
    const uint32_t m_eff = m;
    return enqueue_topm(c, static_cast<const float *>(d_query), w_embed, w_lex, static_cast<const uint32_t *>(d_lex_rows),
                        static_cast<const float *>(d_lex_norm), n_lex, m_eff, static_cast<rlr_cand *>(d_out),
                        static_cast<uint32_t *>(d_out_n), st, nullptr, s->use_half(c->search_flags));
}

RLR_EXPORT int rlr_merge_async(rlr_ctx *c, const void *d_lists, uint32_t n_lists, uint32_t m, void *d_out,
                               void *d_out_n, void *stream)
{
    if (!c || !d_lists || !d_out || !d_out_n) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    if (m == 0 || m > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "m %u not in 1..%d", m, RLR_MAX_M);
    if (n_lists == 0 || n_lists > c->n_lists_cap) return fail(RLR_ERR_UNSUPPORTED, "n_lists %u not in 1..%u", n_lists, c->n_lists_cap);
    CU_TRY(cudaSetDevice(c->s->device));
    uint32_t l = 0;
    CU_TRY(rlr::merge_launch(static_cast<const rlr_cand *>(d_lists), n_lists, m, c->d_tmp, static_cast<rlr_cand *>(d_out),
                             static_cast<uint32_t *>(d_out_n), static_cast<cudaStream_t>(stream), &l));
    c->launches += l;
    return RLR_OK;
}

RLR_EXPORT int rlr_gather_async(rlr_ctx *c, const void *d_cands, const void *d_n, uint32_t m, void *d_out, void *stream)
{
    if (!c || !d_cands || !d_n || !d_out) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    rlr_store *s = c->s;
    CU_TRY(cudaSetDevice(s->device));
    const bool half = s->use_half(c->search_flags);
    CU_TRY(rlr::gather_launch(half ? s->d_rows16 : static_cast<const void *>(s->d_rows), half, half ? s->pitch16 : s->pitch,
                              static_cast<uint32_t>(s->n_rows), static_cast<uint32_t>(s->row_base),
                              static_cast<const rlr_cand *>(d_cands), static_cast<const uint32_t *>(d_n), m,
                              static_cast<float *>(d_out), s->pitch, static_cast<cudaStream_t>(stream)));
    ++c->launches;
    return RLR_OK;
}

RLR_EXPORT int rlr_mmr_async(rlr_ctx *c, const void *d_emb, uint32_t pitch, uint32_t dim, const void *d_cands,
                             const void *d_n, uint32_t p_cap, uint32_t top_k, float lambda, void *d_sel_pos,
                             void *d_sel_n, void *d_result, void *stream)
{
    if (!c || !d_emb || !d_cands || !d_n || !d_sel_pos || !d_sel_n) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    if (p_cap == 0 || p_cap > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "p_cap %u not in 1..%d", p_cap, RLR_MAX_M);
    if (pitch % 4) return fail(RLR_ERR_INVALID_ARG, "pitch must be a multiple of 4 floats");
    CU_TRY(cudaSetDevice(c->s->device));
    rlr::MmrArgs a;
    memset(&a, 0, sizeof a);
    a.d_emb = d_emb; a.half = 0; a.pitch = pitch; a.dim = dim;   // a gathered matrix is always f32
    a.d_cands = static_cast<const rlr_cand *>(d_cands); a.d_n = static_cast<const uint32_t *>(d_n);
    a.use_rows = 0; a.p_cap = p_cap; a.top_k = top_k; a.lambda = lambda;
    a.d_tri = c->d_tri; a.d_sel_pos = static_cast<uint32_t *>(d_sel_pos); a.d_sel_n = static_cast<uint32_t *>(d_sel_n);
    a.d_result = static_cast<rlr_cand *>(d_result);
    a.max_smem_optin = c->s->smem_optin;
    uint32_t l = 0;
    CU_TRY(rlr::mmr_launch(a, static_cast<cudaStream_t>(stream), &l));
    c->launches += l;
    return RLR_OK;
}

RLR_EXPORT int rlr_mmr_store_async(rlr_ctx *c, const void *d_cands, const void *d_n, uint32_t p_cap, uint32_t top_k,
                                   float lambda, void *d_sel_pos, void *d_sel_n, void *d_result, void *stream)
{
    if (!c || !d_cands || !d_n || !d_sel_pos || !d_sel_n) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    if (p_cap == 0 || p_cap > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "p_cap %u not in 1..%d", p_cap, RLR_MAX_M);
    rlr_store *s = c->s;
    CU_TRY(cudaSetDevice(s->device));
    rlr::MmrArgs a;
    memset(&a, 0, sizeof a);
    a.half = s->use_half(c->search_flags);
    a.d_emb = a.half ? s->d_rows16 : static_cast<const void *>(s->d_rows); a.pitch = a.half ? s->pitch16 : s->pitch; a.dim = s->dim;
    a.d_cands = static_cast<const rlr_cand *>(d_cands); a.d_n = static_cast<const uint32_t *>(d_n);
    a.row_base = static_cast<uint32_t>(s->row_base); a.use_rows = 1;
    a.p_cap = p_cap; a.top_k = top_k; a.lambda = lambda;
    a.d_tri = c->d_tri; a.d_sel_pos = static_cast<uint32_t *>(d_sel_pos); a.d_sel_n = static_cast<uint32_t *>(d_sel_n);
    a.d_result = static_cast<rlr_cand *>(d_result);
    a.max_smem_optin = s->smem_optin;
    uint32_t l = 0;
    CU_TRY(rlr::mmr_launch(a, static_cast<cudaStream_t>(stream), &l));
    c->launches += l;
    return RLR_OK;
}

struct rlr_peer_set {
    rlr_store *local = nullptr;
    rlr::PeerTable table;
    void *opened[rlr::kMaxPeers];
    bool half = false;
};

RLR_EXPORT int rlr_store_ipc_export(const rlr_store *s, uint32_t search_flags, void *handle_out)
{
    if (int rc = check_store(s)) return rc;
    if (!handle_out) return fail(RLR_ERR_INVALID_ARG, "handle_out is NULL");
    static_assert(sizeof(cudaIpcMemHandle_t) == RLR_IPC_HANDLE_BYTES, "IPC handle size");
    memset(handle_out, 0, RLR_IPC_HANDLE_BYTES);
    if (s->n_rows == 0) return RLR_OK;
    CU_TRY(cudaSetDevice(s->device));
    const bool half = s->use_half(search_flags);
    if (half && !s->d_rows16) return fail(RLR_ERR_INVALID_ARG, "store holds no f16 copy");
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, half ? s->d_rows16 : static_cast<void *>(s->d_rows)));
    memcpy(handle_out, &h, sizeof h);
    return RLR_OK;
}

RLR_EXPORT int rlr_peer_set_open(rlr_store *local, uint32_t my_index, uint32_t n_shards, const void *handles,
                                 const uint64_t *row_base, const uint64_t *n_rows, uint32_t search_flags,
                                 rlr_peer_set **out)
{
    if (int rc = check_store(local)) return rc;
    if (!handles || !row_base || !n_rows || !out) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    if (n_shards == 0 || n_shards > rlr::kMaxPeers || my_index >= n_shards)
        return fail(RLR_ERR_INVALID_ARG, "n_shards %u / my_index %u out of range (max %d)", n_shards, my_index, rlr::kMaxPeers);
    CU_TRY(cudaSetDevice(local->device));
    rlr_peer_set *p = new rlr_peer_set();
    p->local = local;
    p->half = local->use_half(search_flags);
    memset(&p->table, 0, sizeof p->table);
    memset(p->opened, 0, sizeof p->opened);
    for (uint32_t i = 0; i < n_shards; ++i) {
        p->table.row_base[i] = static_cast<uint32_t>(row_base[i]);
        p->table.n_rows[i] = static_cast<uint32_t>(n_rows[i]);
        if (i == my_index) {
            p->table.base[i] = p->half ? local->d_rows16 : static_cast<void *>(local->d_rows);
        } else if (n_rows[i] != 0) {
            cudaIpcMemHandle_t h;
            memcpy(&h, static_cast<const uint8_t *>(handles) + static_cast<size_t>(i) * RLR_IPC_HANDLE_BYTES, sizeof h);
            cudaError_t e = cudaIpcOpenMemHandle(&p->opened[i], h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                for (uint32_t j = 0; j < i; ++j) if (p->opened[j]) cudaIpcCloseMemHandle(p->opened[j]);
                delete p;
                return fail(RLR_ERR_CUDA, "cudaIpcOpenMemHandle for shard %u failed: %s", i, cudaGetErrorString(e));
            }
            p->table.base[i] = p->opened[i];
        }
    }
    p->table.n = n_shards;
    *out = p;
    return RLR_OK;
}

RLR_EXPORT int rlr_peer_set_close(rlr_peer_set *p)
{
    if (!p) return RLR_OK;
    cudaSetDevice(p->local->device);
    for (int i = 0; i < rlr::kMaxPeers; ++i) if (p->opened[i]) cudaIpcCloseMemHandle(p->opened[i]);
    cudaGetLastError();
    delete p;
    return RLR_OK;
}

RLR_EXPORT int rlr_mmr_peers_async(rlr_ctx *c, rlr_peer_set *p, const void *d_cands, const void *d_n, uint32_t p_cap,
                                   uint32_t top_k, float lambda, void *d_sel_pos, void *d_sel_n, void *d_result,
                                   void *stream)
{
    if (!c || !p || !d_cands || !d_n || !d_sel_pos || !d_sel_n) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    if (p_cap == 0 || p_cap > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "p_cap %u not in 1..%d", p_cap, RLR_MAX_M);
    rlr_store *s = c->s;
    CU_TRY(cudaSetDevice(s->device));
    rlr::MmrArgs a;
    memset(&a, 0, sizeof a);
    a.half = p->half;
    a.d_emb = nullptr; a.pitch = p->half ? s->pitch16 : s->pitch; a.dim = s->dim;
    a.d_cands = static_cast<const rlr_cand *>(d_cands); a.d_n = static_cast<const uint32_t *>(d_n);
    a.use_rows = 1; a.p_cap = p_cap; a.top_k = top_k; a.lambda = lambda;
    a.d_tri = c->d_tri; a.d_sel_pos = static_cast<uint32_t *>(d_sel_pos); a.d_sel_n = static_cast<uint32_t *>(d_sel_n);
    a.d_result = static_cast<rlr_cand *>(d_result);
    a.max_smem_optin = s->smem_optin;
    a.peers = &p->table;
    a.d_gather = c->d_gather;
    uint32_t l = 0;
    CU_TRY(rlr::mmr_launch(a, static_cast<cudaStream_t>(stream), &l));
    c->launches += l;
    return RLR_OK;
}

// ---------------------------------------------------------------------------------
// fused exchange: scan kernels post their lists straight into the root GPU's mailbox
// ---------------------------------------------------------------------------------
namespace {
int mailbox_check_shape(uint32_t n_ranks, uint32_t m_cap, uint32_t ring)
{
    if (n_ranks == 0 || n_ranks > rlr::kMaxPeers) return fail(RLR_ERR_INVALID_ARG, "n_ranks %u not in 1..%d", n_ranks, rlr::kMaxPeers);
    if (m_cap == 0 || m_cap > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "m_cap %u not in 1..%d", m_cap, RLR_MAX_M);
    if (ring < 2 || ring > 64) return fail(RLR_ERR_INVALID_ARG, "ring %u not in 2..64", ring);
    return RLR_OK;
}
} // namespace

RLR_EXPORT int rlr_mailbox_create(int device, uint32_t n_ranks, uint32_t m_cap, uint32_t ring, rlr_mailbox **out)
{
    if (!out) return fail(RLR_ERR_INVALID_ARG, "out is NULL");
    if (int rc = mailbox_check_shape(n_ranks, m_cap, ring)) return rc;
    if (int rc = ensure_device(device)) return rc;
    rlr_mailbox *mb = new rlr_mailbox();
    mb->device = device; mb->owner = true; mb->n_ranks = n_ranks; mb->m_cap = m_cap; mb->ring = ring;
    mb->bytes = mb->total();
    cudaError_t e = cudaMalloc(&mb->base, mb->bytes);
    if (e == cudaSuccess) e = cudaMemset(mb->base, 0, mb->bytes);
    if (e == cudaSuccess) e = cudaMalloc(&mb->d_status, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(mb->d_status, 0, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(mb->base); cudaFree(mb->d_status);
        delete mb;
        return fail(e == cudaErrorMemoryAllocation ? RLR_ERR_OOM : RLR_ERR_CUDA, "mailbox allocation failed: %s", cudaGetErrorString(e));
    }
    *out = mb;
    return RLR_OK;
}

RLR_EXPORT int rlr_mailbox_ipc_export(const rlr_mailbox *mb, void *handle_out)
{
    if (!mb || !handle_out) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    if (!mb->owner) return fail(RLR_ERR_INVALID_ARG, "only the creating (root) rank can export a mailbox");
    CU_TRY(cudaSetDevice(mb->device));
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, mb->base));
    memcpy(handle_out, &h, sizeof h);
    return RLR_OK;
}

RLR_EXPORT int rlr_mailbox_open(int device, const void *handle, uint32_t n_ranks, uint32_t m_cap, uint32_t ring,
                                rlr_mailbox **out)
{
    if (!handle || !out) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    if (int rc = mailbox_check_shape(n_ranks, m_cap, ring)) return rc;
    if (int rc = ensure_device(device)) return rc;
    rlr_mailbox *mb = new rlr_mailbox();
    mb->device = device; mb->owner = false; mb->n_ranks = n_ranks; mb->m_cap = m_cap; mb->ring = ring;
    mb->bytes = mb->total();
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e == cudaSuccess) { mb->base = static_cast<uint8_t *>(p); e = cudaMalloc(&mb->d_status, sizeof(uint32_t)); }
    if (e == cudaSuccess) e = cudaMemset(mb->d_status, 0, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (mb->base) cudaIpcCloseMemHandle(mb->base);
        cudaFree(mb->d_status);
        delete mb;
        return fail(RLR_ERR_CUDA, "opening the root's mailbox failed: %s", cudaGetErrorString(e));
    }
    *out = mb;
    return RLR_OK;
}

RLR_EXPORT int rlr_mailbox_close(rlr_mailbox *mb)
{
    if (!mb) return RLR_OK;
    cudaSetDevice(mb->device);
    if (mb->owner) cudaFree(mb->base);
    else if (mb->base) cudaIpcCloseMemHandle(mb->base);
    cudaFree(mb->d_status);
    cudaGetLastError();
    delete mb;
    return RLR_OK;
}

RLR_EXPORT int rlr_mailbox_status(rlr_mailbox *mb, uint32_t *out)
{
    if (!mb || !out) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    CU_TRY(cudaSetDevice(mb->device));
    CU_TRY(cudaMemcpy(out, mb->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return RLR_OK;
}

RLR_EXPORT int rlr_topm_post_async(rlr_ctx *c, rlr_mailbox *mb, uint32_t my_rank, uint64_t seq, const void *d_query,
                                   float w_embed, float w_lex, const void *d_lex_rows, const void *d_lex_norm,
                                   uint32_t n_lex, uint32_t m, void *stream)
{
    if (!c || !mb || !d_query) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    if (m == 0 || m > mb->m_cap) return fail(RLR_ERR_UNSUPPORTED, "m %u not in 1..%u (mailbox capacity)", m, mb->m_cap);
    if (my_rank >= mb->n_ranks) return fail(RLR_ERR_INVALID_ARG, "my_rank %u >= n_ranks %u", my_rank, mb->n_ranks);
    if (seq == 0) return fail(RLR_ERR_INVALID_ARG, "sequence numbers start at 1");
    rlr_store *s = c->s;
    if (s->n_rows == 0) return fail(RLR_ERR_UNSUPPORTED, "a rank with an empty shard cannot post (give every rank rows)");
    if (s->device != mb->device) return fail(RLR_ERR_INVALID_ARG, "mailbox opened on device %d, store on %d", mb->device, s->device);
    CU_TRY(cudaSetDevice(s->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool half = s->use_half(c->search_flags);
    const uint32_t slot = static_cast<uint32_t>(seq % mb->ring);
    rlr::ScanArgs a;
    memset(&a, 0, sizeof a);
    rlr::scan_plan(s->sm_count, s->smem_optin, static_cast<uint32_t>(s->n_rows), half ? s->pitch16 : s->pitch, half, &a);
    a.tmap = half ? &s->tmap16 : &s->tmap;
    a.d_query = static_cast<const float *>(d_query);
    a.n_rows = static_cast<uint32_t>(s->n_rows);
    a.row_base = static_cast<uint32_t>(s->row_base);
    a.pitch = half ? s->pitch16 : s->pitch;
    a.w_embed = w_embed; a.w_lex = w_lex;
    a.d_lex_rows = static_cast<const uint32_t *>(d_lex_rows); a.d_lex_norm = static_cast<const float *>(d_lex_norm); a.n_lex = n_lex;
    a.m = m;
    a.d_lists = c->d_lists; a.d_counts = c->d_counts; a.d_ticket = c->d_ticket; a.d_pub = c->d_pub;
    a.d_out = mb->list(slot, my_rank); a.d_out_n = mb->count(slot, my_rank);
    a.post.flag = mb->flag(slot, my_rank);
    a.post.consumed = mb->consumed(slot);
    a.post.seq = seq; a.post.ring = mb->ring; a.post.status = mb->d_status;
    CU_TRY(rlr::scan_launch(a, st));
    ++c->launches;
    return RLR_OK;
}

RLR_EXPORT int rlr_mailbox_merge_async(rlr_ctx *c, rlr_mailbox *mb, uint64_t seq, uint32_t m, void *d_out, void *d_out_n,
                                       void *stream)
{
    if (!c || !mb || !d_out || !d_out_n) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    if (!mb->owner) return fail(RLR_ERR_INVALID_ARG, "only the root rank merges its mailbox");
    if (m == 0 || m > mb->m_cap) return fail(RLR_ERR_UNSUPPORTED, "m %u not in 1..%u (mailbox capacity)", m, mb->m_cap);
    if (seq == 0) return fail(RLR_ERR_INVALID_ARG, "sequence numbers start at 1");
    CU_TRY(cudaSetDevice(mb->device));
    const uint32_t slot = static_cast<uint32_t>(seq % mb->ring);
    CU_TRY(rlr::mailbox_merge_launch(mb->list(slot, 0), mb->m_cap, mb->flag(slot, 0), seq, mb->consumed(slot), mb->n_ranks, m,
                                     static_cast<rlr_cand *>(d_out), static_cast<uint32_t *>(d_out_n), mb->d_status,
                                     static_cast<cudaStream_t>(stream)));
    ++c->launches;
    return RLR_OK;
}

RLR_EXPORT int rlr_search_mmr_async(rlr_ctx *c, const void *d_query, uint32_t top_k, float diversity_factor,
                                    float w_embed, float w_lex, void *d_result, void *d_result_n, void *stream)
{
    if (!c || !d_query || !d_result || !d_result_n) return fail(RLR_ERR_INVALID_ARG, "NULL argument");
    rlr_store *s = c->s;
    CU_TRY(cudaSetDevice(s->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float lambda = diversity_factor;
    if (lambda < 0.0f) lambda = 0.0f;
    if (lambda > 1.0f) lambda = 1.0f;
    if (s->n_rows == 0) { CU_TRY(cudaMemsetAsync(d_result_n, 0, sizeof(uint32_t), st)); return RLR_OK; }
    if (lambda == 0.0f) {
        const uint32_t m = static_cast<uint32_t>(std::min<uint64_t>(std::max<uint32_t>(top_k, 1), s->n_rows));
        if (m > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "top_k too large");
        return enqueue_topm(c, static_cast<const float *>(d_query), w_embed, w_lex, nullptr, nullptr, 0, m,
                            static_cast<rlr_cand *>(d_result), static_cast<uint32_t *>(d_result_n), st, nullptr,
                            s->use_half(c->search_flags));
    }
    const uint64_t pool = std::max<uint64_t>(3ull * top_k, static_cast<uint64_t>(top_k) + 10);
    if (pool > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "candidate pool %llu exceeds %d", (unsigned long long)pool, RLR_MAX_M);
    const uint32_t p = static_cast<uint32_t>(std::min<uint64_t>(pool, s->n_rows));
    const bool half = s->use_half(c->search_flags);
    if (int rc = enqueue_topm(c, static_cast<const float *>(d_query), w_embed, w_lex, nullptr, nullptr, 0, p, c->d_pool,
                              c->d_pool_n, st, nullptr, half))
        return rc;
    rlr::MmrArgs a;
    memset(&a, 0, sizeof a);
    a.half = half;
    a.d_emb = half ? s->d_rows16 : static_cast<const void *>(s->d_rows); a.pitch = half ? s->pitch16 : s->pitch; a.dim = s->dim;
    a.d_cands = c->d_pool; a.d_n = c->d_pool_n;
    a.row_base = static_cast<uint32_t>(s->row_base); a.use_rows = 1;
    a.p_cap = p; a.top_k = top_k; a.lambda = lambda;
    a.d_tri = c->d_tri; a.d_sel_pos = c->d_sel_pos; a.d_sel_n = static_cast<uint32_t *>(d_result_n);
    a.d_result = static_cast<rlr_cand *>(d_result);
    a.max_smem_optin = s->smem_optin;
    uint32_t l = 0;
    CU_TRY(rlr::mmr_launch(a, st, &l));
    c->launches += l;
    return RLR_OK;
}

RLR_EXPORT int rlr_time_scan(rlr_ctx *c, const void *d_query, uint32_t m, uint32_t iters, void *stream,
                             float *out_ms_per_launch)
{
    if (!c || !d_query || !out_ms_per_launch || iters == 0) return fail(RLR_ERR_INVALID_ARG, "bad argument");
    if (m == 0 || m > RLR_MAX_M) return fail(RLR_ERR_UNSUPPORTED, "m %u not in 1..%d", m, RLR_MAX_M);
    rlr_store *s = c->s;
    if (s->n_rows == 0) return fail(RLR_ERR_INVALID_ARG, "empty store");
    CU_TRY(cudaSetDevice(s->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rlr::ScanArgs a;
    memset(&a, 0, sizeof a);
    const bool half = s->use_half(c->search_flags);
    rlr::scan_plan(s->sm_count, s->smem_optin, static_cast<uint32_t>(s->n_rows), half ? s->pitch16 : s->pitch, half, &a);
    a.tmap = half ? &s->tmap16 : &s->tmap; a.d_query = static_cast<const float *>(d_query);
    a.n_rows = static_cast<uint32_t>(s->n_rows); a.row_base = static_cast<uint32_t>(s->row_base);
    a.pitch = half ? s->pitch16 : s->pitch;
    a.w_embed = 0.7f; a.w_lex = 0.3f; a.m = m; a.d_lists = c->d_lists; a.d_counts = c->d_counts;
    a.d_ticket = c->d_ticket; a.d_pub = c->d_pub; a.d_out = c->d_pool; a.d_out_n = c->d_pool_n;
    CU_TRY(rlr::scan_launch(a, st)); // warm
    CU_TRY(cudaEventRecord(c->ev[0], st));
    for (uint32_t i = 0; i < iters; ++i) CU_TRY(rlr::scan_launch(a, st));
    CU_TRY(cudaEventRecord(c->ev[1], st));
    CU_TRY(cudaEventSynchronize(c->ev[1]));
    c->launches += iters + 1;
    float ms = 0;
    CU_TRY(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
    *out_ms_per_launch = ms / iters;
    if (getenv("RLR_DEBUG_TRACE")) {
        // dev-only: one more launch with phase timestamps, printed to stderr
        const int g = a.grid;
        unsigned long long *d_tr = nullptr;
        std::vector<unsigned long long> h(5 * g + 16, 0);
        CU_TRY(cudaMalloc(&d_tr, h.size() * 8));
        CU_TRY(cudaMemset(d_tr, 0, h.size() * 8));
        a.d_trace = d_tr;
        CU_TRY(rlr::scan_launch(a, st));
        CU_TRY(cudaStreamSynchronize(st));
        CU_TRY(cudaMemcpy(h.data(), d_tr, h.size() * 8, cudaMemcpyDeviceToHost));
        cudaFree(d_tr);
        unsigned long long t0 = ~0ull, loop_max = 0, list_max = 0, loop_min = ~0ull;
        for (int i = 0; i < g; ++i) { t0 = std::min(t0, h[i]); loop_max = std::max(loop_max, h[g + i]); loop_min = std::min(loop_min, h[g + i]); list_max = std::max(list_max, h[2 * g + i]); }
        unsigned long long cnt_sum = 0, cnt2_sum = 0, cnt_max = 0, tiles_min = ~0ull, tiles_max = 0;
        for (int i = 0; i < g; ++i) {
            const unsigned long long w = h[3 * g + i], cn = (w >> 24) & 0xffffffu, tl = w >> 48;
            cnt_sum += cn; cnt2_sum += w & 0xffffffu; cnt_max = std::max(cnt_max, cn);
            tiles_min = std::min(tiles_min, tl); tiles_max = std::max(tiles_max, tl);
        }
        fprintf(stderr, "[trace m=%u] tiles per CTA min/max %llu/%llu\n", m, tiles_min, tiles_max);
        const unsigned long long *tr = h.data() + 4 * g;
        fprintf(stderr, "[trace m=%u] loop_end first/last %.1f/%.1f us, lists written by %.1f us; buffer entries avg %.0f max %llu, after global-bound compaction avg %.1f\n",
                m, (loop_min - t0) / 1e3, (loop_max - t0) / 1e3, (list_max - t0) / 1e3, double(cnt_sum) / g, cnt_max, double(cnt2_sum) / g);
        fprintf(stderr, "[trace m=%u] merge: start %.1f, sample loaded +%.1f, sorted +%.1f, extras counted +%.1f, ready +%.1f, done +%.1f us (total valid %llu, extras %llu, c %llu, n2 %llu)\n",
                m, (tr[0] - t0) / 1e3, (tr[1] - tr[0]) / 1e3, (tr[2] - tr[1]) / 1e3, (tr[3] - tr[2]) / 1e3, (tr[4] - tr[3]) / 1e3, (tr[7] - tr[4]) / 1e3,
                tr[8] >> 32, tr[8] & 0xffffffffu, tr[9] >> 32, tr[9] & 0xffffffffu);
    }
    return RLR_OK;
}
