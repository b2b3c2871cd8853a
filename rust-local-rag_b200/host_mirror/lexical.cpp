// lexical.cpp -- HOST-MIRROR SUPPORT, not part of the product library: a C++ twin of the reference's
// LexicalIndex (BM25) and tokenizer, /root/reference/src/rag_engine.rs:2083-2247, built into its own
// librlr_hostmirror.so (plain g++, no CUDA).
//
// SURVEY.md section 2 row 7 marks LexicalIndex OUT OF SCOPE (host string/hash work) and 8(f) N4 (scoring the
// postings on the device) is DECLINED -- DESIGN.md section 8 has the measurement behind that decision.  A Rust
// maintainer keeps using the reference's own LexicalIndex and hands its <= 5*top_k (row, score) pairs to
// rlr_search_topm / rlr_search_mmr / rlr_cluster_*, which blend them on the device.  This twin exists only so
// that the NON-Rust host mirrors shipped here (include/rlr_engine.hpp, rust-local-rag_b200/engine.py) can answer
// text queries in their tests; nothing in librlr_b200.so depends on it.
//
// Fidelity notes:
//  * arithmetic: f32 throughout, the reference's operation order (:2193-2219); ln is the C library's logf,
//    which is what Rust's f32::ln calls on Linux.
//  * the reference sums a document's per-term scores in HashSet iteration order (random per process, :2194)
//    and sorts ties in HashMap order (:2222-2223): its output is only defined up to f32 summation order and
//    tie order.  This twin is deterministic: terms in bytewise order, ties by ascending chunk key -- one of
//    the reference's valid outcomes.
//  * tokenize (:2242-2247): split on !char::is_alphanumeric, keep tokens of >= 3 BYTES, str::to_lowercase.
//    Classification (Alphabetic | Nd | Nl | No) and the full lowercase mapping, U+0130's two-code-point
//    mapping and the Final_Sigma rule included, come from GENERATED Unicode tables (unicode_tables.inc,
//    tools/gen_unicode_tables.py; tests/test_lexical_cpu.py checks every code point up to U+10FFFF against
//    the generating databases).  The Unicode versions are recorded in the generated file; Rust 1.88 ships 16.0.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/rlr_hostmirror.h"

#define RLR_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

#include "unicode_tables.inc"

template <size_t N>
bool in_ranges(const uint32_t (&t)[N][2], uint32_t c)
{
    size_t lo = 0, hi = N;                       // first range whose end is >= c
    while (lo < hi) { const size_t mid = (lo + hi) / 2; if (t[mid][1] < c) lo = mid + 1; else hi = mid; }
    return lo < N && t[lo][0] <= c;
}

bool is_alnum_cp(uint32_t c)
{
    if (c < 0x80) return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z');
    return in_ranges(kAlnum, c);
}
bool is_cased_cp(uint32_t c) { return in_ranges(kCased, c); }
bool is_case_ignorable_cp(uint32_t c) { return in_ranges(kCaseIgnorable, c); }

// full lowercase mapping of one code point: writes 1..3 code points, returns how many
uint32_t lower_cp(uint32_t c, uint32_t out[3])
{
    if (c < 0x80) { out[0] = (c >= 'A' && c <= 'Z') ? c + 32 : c; return 1; }
    const size_t n = sizeof(kLower) / sizeof(kLower[0]);
    size_t lo = 0, hi = n;
    while (lo < hi) { const size_t mid = (lo + hi) / 2; if (kLower[mid].cp < c) lo = mid + 1; else hi = mid; }
    if (lo < n && kLower[lo].cp == c) { for (uint32_t i = 0; i < kLower[lo].n; ++i) out[i] = kLower[lo].to[i]; return kLower[lo].n; }
    out[0] = c;
    return 1;
}

void append_utf8(std::string &s, uint32_t c)
{
    if (c < 0x80) s.push_back(static_cast<char>(c));
    else if (c < 0x800) { s.push_back(static_cast<char>(0xC0 | (c >> 6))); s.push_back(static_cast<char>(0x80 | (c & 0x3F))); }
    else if (c < 0x10000) { s.push_back(static_cast<char>(0xE0 | (c >> 12))); s.push_back(static_cast<char>(0x80 | ((c >> 6) & 0x3F))); s.push_back(static_cast<char>(0x80 | (c & 0x3F))); }
    else { s.push_back(static_cast<char>(0xF0 | (c >> 18))); s.push_back(static_cast<char>(0x80 | ((c >> 12) & 0x3F))); s.push_back(static_cast<char>(0x80 | ((c >> 6) & 0x3F))); s.push_back(static_cast<char>(0x80 | (c & 0x3F))); }
}

// str::to_lowercase of one token (a run of alphanumeric code points): full mapping per code point, and
// U+03A3 -> U+03C2 where Final_Sigma holds: preceded by a cased letter (skipping case-ignorable ones) and not
// followed by one (library/alloc/src/str.rs: map_uppercase_sigma / case_ignorable_then_cased).
std::string lowercase_token(const std::vector<uint32_t> &cps)
{
    std::string out;
    for (size_t i = 0; i < cps.size(); ++i) {
        const uint32_t c = cps[i];
        if (c == 0x03A3) {
            bool before = false, after = false;
            for (size_t j = i; j-- > 0;) { if (is_case_ignorable_cp(cps[j])) continue; before = is_cased_cp(cps[j]); break; }
            for (size_t j = i + 1; j < cps.size(); ++j) { if (is_case_ignorable_cp(cps[j])) continue; after = is_cased_cp(cps[j]); break; }
            append_utf8(out, before && !after ? 0x03C2 : 0x03C3);
            continue;
        }
        uint32_t lo[3];
        const uint32_t n = lower_cp(c, lo);
        for (uint32_t k = 0; k < n; ++k) append_utf8(out, lo[k]);
    }
    return out;
}

// fn tokenize, :2242-2247
std::vector<std::string> tokenize(const char *text, size_t len)
{
    std::vector<std::string> out;
    std::vector<uint32_t> cur;
    size_t cur_bytes = 0;                       // byte length of the ORIGINAL token (the filter runs before to_lowercase)
    auto flush = [&] {
        if (cur_bytes >= 3) out.push_back(lowercase_token(cur));
        cur.clear();
        cur_bytes = 0;
    };
    size_t i = 0;
    while (i < len) {
        const unsigned char b = static_cast<unsigned char>(text[i]);
        uint32_t c;
        size_t n;
        if (b < 0x80) { c = b; n = 1; }
        else if ((b >> 5) == 6 && i + 1 < len) { c = ((b & 0x1Fu) << 6) | (static_cast<unsigned char>(text[i + 1]) & 0x3Fu); n = 2; }
        else if ((b >> 4) == 14 && i + 2 < len) { c = ((b & 0x0Fu) << 12) | ((static_cast<unsigned char>(text[i + 1]) & 0x3Fu) << 6) | (static_cast<unsigned char>(text[i + 2]) & 0x3Fu); n = 3; }
        else if ((b >> 3) == 30 && i + 3 < len) { c = ((b & 0x07u) << 18) | ((static_cast<unsigned char>(text[i + 1]) & 0x3Fu) << 12) | ((static_cast<unsigned char>(text[i + 2]) & 0x3Fu) << 6) | (static_cast<unsigned char>(text[i + 3]) & 0x3Fu); n = 4; }
        else { c = 0xFFFD; n = 1; }             // invalid byte: a separator (Rust strings cannot hold it)
        if (c != 0xFFFD && is_alnum_cp(c)) { cur.push_back(c); cur_bytes += n; }
        else flush();
        i += n;
    }
    flush();
    return out;
}

} // namespace

struct rlr_lexical {
    // struct LexicalIndex, :2084-2090 (chunk ids are opaque u64 keys chosen by the caller)
    std::map<std::string, std::unordered_map<uint64_t, uint64_t>> term_postings;   // ordered: deterministic term order
    std::unordered_map<uint64_t, uint64_t> doc_lengths;
    std::unordered_map<uint64_t, std::unordered_map<std::string, uint64_t>> doc_terms;
    uint64_t total_docs = 0, total_length = 0;

    void remove_chunk(uint64_t id)              // :2140-2167
    {
        auto it = doc_terms.find(id);
        if (it != doc_terms.end()) {
            for (auto &tc : it->second) {
                auto p = term_postings.find(tc.first);
                if (p != term_postings.end()) {
                    p->second.erase(id);
                    if (p->second.empty()) term_postings.erase(p);
                }
            }
            doc_terms.erase(it);
            auto l = doc_lengths.find(id);
            if (l != doc_lengths.end()) {
                total_length = total_length >= l->second ? total_length - l->second : 0;
                doc_lengths.erase(l);
            }
            if (total_docs > 0) --total_docs;
        } else {
            doc_lengths.erase(id);
        }
        if (total_docs == 0) total_length = 0;
    }

    void add_chunk(uint64_t id, const char *text, size_t len)   // :2106-2138
    {
        if (doc_terms.count(id)) remove_chunk(id);
        const std::vector<std::string> tokens = tokenize(text, len);
        if (tokens.empty()) return;
        std::unordered_map<std::string, uint64_t> counts;
        for (auto &t : tokens) ++counts[t];
        uint64_t doc_length = 0;
        for (auto &c : counts) doc_length += c.second;
        if (doc_length == 0) return;
        for (auto &c : counts) term_postings[c.first][id] = c.second;
        doc_lengths[id] = doc_length;
        doc_terms[id] = std::move(counts);
        ++total_docs;
        total_length += doc_length;
    }

    std::vector<std::pair<uint64_t, float>> score(const char *query, size_t len, uint32_t limit) const   // :2169-2227
    {
        std::vector<std::pair<uint64_t, float>> results;
        if (total_docs == 0) return results;
        std::vector<std::string> terms = tokenize(query, len);
        if (terms.empty()) return results;
        std::sort(terms.begin(), terms.end());
        terms.erase(std::unique(terms.begin(), terms.end()), terms.end());
        const float avg_doc_len = static_cast<float>(total_length) / static_cast<float>(total_docs);
        const float k1 = 1.5f, b = 0.75f;
        std::unordered_map<uint64_t, float> scores;
        for (auto &term : terms) {
            auto p = term_postings.find(term);
            if (p == term_postings.end()) continue;
            const float df = static_cast<float>(p->second.size());
            volatile float num = static_cast<float>(total_docs) - df;     // volatile: no re-association by the compiler
            num = num + 0.5f;
            volatile float den = df + 0.5f;
            volatile float ratio = num / den;
            float idf = logf(ratio);
            idf = fmaxf(idf, 0.0f);
            for (auto &post : p->second) {
                auto dl = doc_lengths.find(post.first);
                const float doc_length = dl == doc_lengths.end() ? 0.0f : static_cast<float>(dl->second);
                if (doc_length == 0.0f) continue;
                const float tf = static_cast<float>(post.second);
                volatile float x = doc_length / avg_doc_len;
                x = b * x;
                volatile float one_minus_b = 1.0f - b;
                x = one_minus_b + x;
                x = k1 * x;
                volatile float denom = tf + x;
                if (denom == 0.0f) continue;
                volatile float k1p1 = k1 + 1.0f;
                volatile float t2 = tf * k1p1;
                t2 = idf * t2;
                const float sc = t2 / denom;
                auto ins = scores.emplace(post.first, 0.0f);
                volatile float acc = ins.first->second;
                acc = acc + sc;
                ins.first->second = acc;
            }
        }
        results.assign(scores.begin(), scores.end());
        std::sort(results.begin(), results.end(), [](const std::pair<uint64_t, float> &a, const std::pair<uint64_t, float> &c) {
            if (a.second != c.second) return a.second > c.second;      // score desc (:2223)
            return a.first < c.first;                                  // ties: ascending key
        });
        if (limit > 0 && results.size() > limit) results.resize(limit);
        return results;
    }
};

namespace {
thread_local std::string g_hm_err;
int lex_fail(int code, const char *msg)
{
    g_hm_err = msg;
    return code;
}
} // namespace

RLR_EXPORT const char *rlr_hostmirror_last_error(void) { return g_hm_err.c_str(); }

// test hook: bit 0 of out[cp] = is_alphanumeric(cp); lower[3*cp ..] = its full lowercase mapping (0-padded)
RLR_EXPORT int rlr_hostmirror_unicode_dump(uint8_t *out_alnum, uint32_t *out_lower, uint32_t n_cp)
{
    if (!out_alnum || !out_lower) return lex_fail(RLR_HM_ERR_INVALID_ARG, "NULL argument");
    for (uint32_t c = 0; c < n_cp; ++c) {
        out_alnum[c] = is_alnum_cp(c) ? 1 : 0;
        uint32_t lo[3] = {0, 0, 0};
        lower_cp(c, lo);
        out_lower[3 * c] = lo[0]; out_lower[3 * c + 1] = lo[1]; out_lower[3 * c + 2] = lo[2];
    }
    return RLR_HM_OK;
}

RLR_EXPORT int rlr_lexical_create(rlr_lexical **out)
{
    if (!out) return lex_fail(RLR_HM_ERR_INVALID_ARG, "out is NULL");
    *out = new rlr_lexical();
    return RLR_HM_OK;
}

RLR_EXPORT int rlr_lexical_destroy(rlr_lexical *lx)
{
    delete lx;
    return RLR_HM_OK;
}

RLR_EXPORT int rlr_lexical_add_chunk(rlr_lexical *lx, uint64_t chunk_key, const char *text, size_t len)
{
    if (!lx || (!text && len)) return lex_fail(RLR_HM_ERR_INVALID_ARG, "NULL argument");
    lx->add_chunk(chunk_key, text ? text : "", len);
    return RLR_HM_OK;
}

RLR_EXPORT int rlr_lexical_remove_chunk(rlr_lexical *lx, uint64_t chunk_key)
{
    if (!lx) return lex_fail(RLR_HM_ERR_INVALID_ARG, "NULL argument");
    lx->remove_chunk(chunk_key);
    return RLR_HM_OK;
}

RLR_EXPORT int rlr_lexical_contains(const rlr_lexical *lx, uint64_t chunk_key, int *out)
{
    if (!lx || !out) return lex_fail(RLR_HM_ERR_INVALID_ARG, "NULL argument");
    *out = lx->doc_terms.count(chunk_key) ? 1 : 0;      // :2229-2231
    return RLR_HM_OK;
}

RLR_EXPORT int rlr_lexical_stats(const rlr_lexical *lx, uint64_t *total_docs, uint64_t *total_length, uint64_t *n_terms)
{
    if (!lx) return lex_fail(RLR_HM_ERR_INVALID_ARG, "NULL argument");
    if (total_docs) *total_docs = lx->total_docs;
    if (total_length) *total_length = lx->total_length;
    if (n_terms) *n_terms = lx->term_postings.size();
    return RLR_HM_OK;
}

RLR_EXPORT int rlr_lexical_score(const rlr_lexical *lx, const char *query, size_t len, uint32_t limit, uint64_t *out_keys,
                                 float *out_scores, uint32_t cap, uint32_t *out_n)
{
    if (!lx || (!query && len) || !out_n) return lex_fail(RLR_HM_ERR_INVALID_ARG, "NULL argument");
    const auto r = lx->score(query ? query : "", len, limit);
    const uint32_t n = static_cast<uint32_t>(std::min<size_t>(r.size(), cap));
    if (n && (!out_keys || !out_scores)) return lex_fail(RLR_HM_ERR_INVALID_ARG, "output buffers are NULL");
    for (uint32_t i = 0; i < n; ++i) { out_keys[i] = r[i].first; out_scores[i] = r[i].second; }
    *out_n = n;
    if (r.size() > cap) return lex_fail(RLR_HM_ERR_UNSUPPORTED, "output capacity too small for the result (pass cap >= limit)");
    return RLR_HM_OK;
}

RLR_EXPORT int rlr_tokenize(const char *text, size_t len, char *out, size_t out_cap, size_t *out_len, uint32_t *out_tokens)
{
    // tokens joined by '\n' (a separator no token can contain)
    if ((!text && len) || !out_len) return lex_fail(RLR_HM_ERR_INVALID_ARG, "NULL argument");
    const auto toks = tokenize(text ? text : "", len);
    std::string joined;
    for (size_t i = 0; i < toks.size(); ++i) { if (i) joined.push_back('\n'); joined += toks[i]; }
    *out_len = joined.size();
    if (out_tokens) *out_tokens = static_cast<uint32_t>(toks.size());
    if (joined.size() > out_cap) return lex_fail(RLR_HM_ERR_UNSUPPORTED, "output capacity too small");
    if (!joined.empty()) {
        if (!out) return lex_fail(RLR_HM_ERR_INVALID_ARG, "out is NULL");
        memcpy(out, joined.data(), joined.size());
    }
    return RLR_HM_OK;
}
