// lexical.cpp -- host-side BM25 index: the LexicalIndex the hot path blends with
// (/root/reference/src/rag_engine.rs:2083-2247; consumed by RagEngine::search :505-532).
//
// SURVEY.md 8(f) N4.  The reference keeps this index on the host (string/hash work) and so does
// this build: the scan kernel takes its OUTPUT (<= 5*top_k (row, score) pairs) and blends it
// in-kernel.  A maintainer's Rust glue keeps using the reference's own LexicalIndex; this C++
// twin exists so that the host mirrors shipped here (include/rlr_engine.hpp,
// rust-local-rag_b200/engine.py) reproduce `search` on real chunk text, not only on embeddings.
//
// Fidelity notes (all stated in DESIGN.md):
//  * arithmetic: f32 throughout, the reference's operation order (:2193-2219); ln is the C
//    library's logf, which is what Rust's f32::ln calls on Linux.
//  * the reference sums a document's per-term scores in HashSet iteration order (random per
//    process, :2194) and sorts ties in HashMap order (:2222-2223): its output is only defined up
//    to f32 summation order and tie order.  This twin is deterministic: terms in bytewise order,
//    ties by ascending chunk key -- one of the reference's valid outcomes.
//  * tokenize (:2242-2247): split on !char::is_alphanumeric, keep tokens of >= 3 BYTES, lowercase.
//    ASCII, Latin-1, Latin Extended-A, Greek and Cyrillic are classified and case-folded like
//    Rust does; other non-ASCII code points are treated as letters without case mapping, except
//    the punctuation/symbol blocks listed in is_alnum_cp (an approximation of the Unicode tables).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/rlr_b200.h"

#define RLR_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

bool is_alnum_cp(uint32_t c)
{
    if (c < 0x80) return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z');
    if (c < 0xC0) return c == 0xAA || c == 0xB2 || c == 0xB3 || c == 0xB5 || c == 0xB9 || c == 0xBA || (c >= 0xBC && c <= 0xBE);
    if (c == 0xD7 || c == 0xF7) return false;
    if (c >= 0x02C2 && c <= 0x02C5) return false;
    if (c >= 0x02D2 && c <= 0x02DF) return false;
    if (c >= 0x0300 && c <= 0x036F) return c == 0x0345;             // combining marks (only ypogegrammeni is Alphabetic)
    if (c == 0x037E || c == 0x0387 || c == 0x0482) return false;
    if (c >= 0x2000 && c <= 0x206F) return false;                   // general punctuation
    if (c >= 0x20A0 && c <= 0x20FF) return false;                   // currency, combining marks for symbols
    if (c >= 0x2100 && c <= 0x214F)                                  // letterlike symbols: the Alphabetic ones only
        return c == 0x2102 || c == 0x2107 || (c >= 0x210A && c <= 0x2113) || c == 0x2115 || (c >= 0x2119 && c <= 0x211D) ||
               c == 0x2124 || c == 0x2126 || c == 0x2128 || (c >= 0x212A && c <= 0x212D) || (c >= 0x212F && c <= 0x2139) ||
               (c >= 0x213C && c <= 0x213F) || (c >= 0x2145 && c <= 0x2149) || c == 0x214E;
    if (c >= 0x2190 && c <= 0x245F) return false;                   // arrows, math, technical, control pictures, OCR
    if (c >= 0x2500 && c <= 0x2BFF) return (c >= 0x2776 && c <= 0x2793);   // box drawing .. misc symbols (dingbat digits are No)
    if (c >= 0x2E00 && c <= 0x2E7F) return c == 0x2E2F;             // supplemental punctuation
    if (c >= 0x3000 && c <= 0x303F) return (c >= 0x3005 && c <= 0x3007) || (c >= 0x3021 && c <= 0x3029) || (c >= 0x3031 && c <= 0x3035) || (c >= 0x3038 && c <= 0x303C);
    if (c >= 0xE000 && c <= 0xF8FF) return false;                   // private use
    if (c >= 0xFE10 && c <= 0xFE6F) return false;                   // vertical forms, small form variants
    if (c >= 0xFF01 && c <= 0xFF0F) return false;
    if (c >= 0xFF1A && c <= 0xFF20) return false;
    if (c >= 0xFF3B && c <= 0xFF40) return false;
    if (c >= 0xFF5B && c <= 0xFF65) return false;
    if (c >= 0xFFE0) return false;
    return true;
}

uint32_t lower_cp(uint32_t c)
{
    if (c < 0x80) return (c >= 'A' && c <= 'Z') ? c + 32 : c;
    if (c >= 0xC0 && c <= 0xDE && c != 0xD7) return c + 32;
    if (c >= 0x0100 && c <= 0x017F) {
        if (c == 0x0130) return c;                                  // I with dot: multi-char mapping in Rust; left as is
        if (c == 0x0178) return 0xFF;
        if ((c >= 0x0139 && c <= 0x0148) || (c >= 0x0179 && c <= 0x017E)) return (c & 1) ? c + 1 : c;
        if (c == 0x0138 || c == 0x0149 || c == 0x017F) return c;
        return (c & 1) ? c : c + 1;
    }
    if (c >= 0x0391 && c <= 0x03A9 && c != 0x03A2) return c + 32;   // (Rust maps a word-final sigma contextually; not reproduced)
    if (c >= 0x0386 && c <= 0x038F) {
        if (c == 0x0386) return 0x03AC;
        if (c >= 0x0388 && c <= 0x038A) return c + 37;
        if (c == 0x038C) return 0x03CC;
        if (c == 0x038E || c == 0x038F) return c + 63;
        return c;
    }
    if (c >= 0x0400 && c <= 0x040F) return c + 80;
    if (c >= 0x0410 && c <= 0x042F) return c + 32;
    if (c >= 0x0460 && c <= 0x0481) return (c & 1) ? c : c + 1;
    if (c >= 0x048A && c <= 0x04BF) return (c & 1) ? c : c + 1;
    if (c >= 0x1E00 && c <= 0x1E95) return (c & 1) ? c : c + 1;     // Latin Extended Additional
    if (c >= 0x1EA0 && c <= 0x1EFF) return (c & 1) ? c : c + 1;
    if (c >= 0xFF21 && c <= 0xFF3A) return c + 32;                  // fullwidth Latin
    return c;
}

void append_utf8(std::string &s, uint32_t c)
{
    if (c < 0x80) s.push_back(static_cast<char>(c));
    else if (c < 0x800) { s.push_back(static_cast<char>(0xC0 | (c >> 6))); s.push_back(static_cast<char>(0x80 | (c & 0x3F))); }
    else if (c < 0x10000) { s.push_back(static_cast<char>(0xE0 | (c >> 12))); s.push_back(static_cast<char>(0x80 | ((c >> 6) & 0x3F))); s.push_back(static_cast<char>(0x80 | (c & 0x3F))); }
    else { s.push_back(static_cast<char>(0xF0 | (c >> 18))); s.push_back(static_cast<char>(0x80 | ((c >> 12) & 0x3F))); s.push_back(static_cast<char>(0x80 | ((c >> 6) & 0x3F))); s.push_back(static_cast<char>(0x80 | (c & 0x3F))); }
}

// fn tokenize, :2242-2247
std::vector<std::string> tokenize(const char *text, size_t len)
{
    std::vector<std::string> out;
    std::string cur;
    size_t cur_bytes = 0;                       // byte length of the ORIGINAL token (the filter runs before to_lowercase)
    auto flush = [&] {
        if (cur_bytes >= 3) out.push_back(cur);
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
        if (c != 0xFFFD && is_alnum_cp(c)) { append_utf8(cur, lower_cp(c)); cur_bytes += n; }
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

extern "C" void rlr_internal_set_error(const char *msg);   // api.cu: the thread-local message behind rlr_last_error()

namespace {
int lex_fail(int code, const char *msg)
{
    rlr_internal_set_error(msg);
    return code;
}
} // namespace

RLR_EXPORT int rlr_lexical_create(rlr_lexical **out)
{
    if (!out) return lex_fail(RLR_ERR_INVALID_ARG, "out is NULL");
    *out = new rlr_lexical();
    return RLR_OK;
}

RLR_EXPORT int rlr_lexical_destroy(rlr_lexical *lx)
{
    delete lx;
    return RLR_OK;
}

RLR_EXPORT int rlr_lexical_add_chunk(rlr_lexical *lx, uint64_t chunk_key, const char *text, size_t len)
{
    if (!lx || (!text && len)) return lex_fail(RLR_ERR_INVALID_ARG, "NULL argument");
    lx->add_chunk(chunk_key, text ? text : "", len);
    return RLR_OK;
}

RLR_EXPORT int rlr_lexical_remove_chunk(rlr_lexical *lx, uint64_t chunk_key)
{
    if (!lx) return lex_fail(RLR_ERR_INVALID_ARG, "NULL argument");
    lx->remove_chunk(chunk_key);
    return RLR_OK;
}

RLR_EXPORT int rlr_lexical_contains(const rlr_lexical *lx, uint64_t chunk_key, int *out)
{
    if (!lx || !out) return lex_fail(RLR_ERR_INVALID_ARG, "NULL argument");
    *out = lx->doc_terms.count(chunk_key) ? 1 : 0;      // :2229-2231
    return RLR_OK;
}

RLR_EXPORT int rlr_lexical_stats(const rlr_lexical *lx, uint64_t *total_docs, uint64_t *total_length, uint64_t *n_terms)
{
    if (!lx) return lex_fail(RLR_ERR_INVALID_ARG, "NULL argument");
    if (total_docs) *total_docs = lx->total_docs;
    if (total_length) *total_length = lx->total_length;
    if (n_terms) *n_terms = lx->term_postings.size();
    return RLR_OK;
}

RLR_EXPORT int rlr_lexical_score(const rlr_lexical *lx, const char *query, size_t len, uint32_t limit, uint64_t *out_keys,
                                 float *out_scores, uint32_t cap, uint32_t *out_n)
{
    if (!lx || (!query && len) || !out_n) return lex_fail(RLR_ERR_INVALID_ARG, "NULL argument");
    const auto r = lx->score(query ? query : "", len, limit);
    const uint32_t n = static_cast<uint32_t>(std::min<size_t>(r.size(), cap));
    if (n && (!out_keys || !out_scores)) return lex_fail(RLR_ERR_INVALID_ARG, "output buffers are NULL");
    for (uint32_t i = 0; i < n; ++i) { out_keys[i] = r[i].first; out_scores[i] = r[i].second; }
    *out_n = n;
    if (r.size() > cap) return lex_fail(RLR_ERR_UNSUPPORTED, "output capacity too small for the result (pass cap >= limit)");
    return RLR_OK;
}

RLR_EXPORT int rlr_tokenize(const char *text, size_t len, char *out, size_t out_cap, size_t *out_len, uint32_t *out_tokens)
{
    // tokens joined by '\n' (a separator no token can contain)
    if ((!text && len) || !out_len) return lex_fail(RLR_ERR_INVALID_ARG, "NULL argument");
    const auto toks = tokenize(text ? text : "", len);
    std::string joined;
    for (size_t i = 0; i < toks.size(); ++i) { if (i) joined.push_back('\n'); joined += toks[i]; }
    *out_len = joined.size();
    if (out_tokens) *out_tokens = static_cast<uint32_t>(toks.size());
    if (joined.size() > out_cap) return lex_fail(RLR_ERR_UNSUPPORTED, "output capacity too small");
    if (!joined.empty()) {
        if (!out) return lex_fail(RLR_ERR_INVALID_ARG, "out is NULL");
        memcpy(out, joined.data(), joined.size());
    }
    return RLR_OK;
}
