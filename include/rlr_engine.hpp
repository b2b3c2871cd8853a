// rlr_engine.hpp -- C++ host-side mirror of the reference's `RagEngine` for the retrieval path,
// header-only, above the C ABI of rlr_b200.h.
//
// The reference is Rust (src/rag_engine.rs) and there is no Rust toolchain in this image, so this
// is the compiled-language host layer: same method names, argument meaning and error behaviour as
//   RagEngine::search                    src/rag_engine.rs:470-701 (reranker absent)
//   RagEngine::search_with_diversity     :717-759
//   RagEngine::get_embedding_candidates  :415-461
//   load_from_disk / apply_loaded_state  :1520-1696  (chunks_{model}.json, version gate, re-normalise)
//   add_document's store update          :347-386    (replace_document)
//   search_documents clamps              src/mcp_server.rs:81-110
// Errors surface as rlr::Error (what `anyhow::Error` is on the Rust side).  Everything numeric
// happens in librlr_b200.so on the GPU; there is no CPU fallback.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "rlr_b200.h"
#include "rlr_hostmirror.h"   // BM25 twin for text queries: librlr_hostmirror.so (host-mirror support, not the boundary)

namespace rlr {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check_hm(int rc)
{
    if (rc != RLR_HM_OK) throw Error(rc, std::string("rlr_hostmirror error ") + std::to_string(rc) + ": " + rlr_hostmirror_last_error());
}
inline void check(int rc)
{
    if (rc != RLR_OK) throw Error(rc, std::string("rlr_b200 error ") + std::to_string(rc) + ": " + rlr_last_error());
}

// ---- minimal JSON reader (objects keep insertion order: rows follow file order) ----
struct Json {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;                      // Str value, or the raw token of a Num
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;
    const Json *get(const std::string &k) const
    {
        for (auto &kv : obj) if (kv.first == k) return &kv.second;
        return nullptr;
    }
};
class JsonParser {
    const std::string &s;
    size_t i = 0;
    [[noreturn]] void bad(const char *m) const { throw Error(RLR_ERR_INVALID_ARG, std::string("JSON: ") + m + " at byte " + std::to_string(i)); }
    void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i; }
    std::string string_()
    {
        if (s[i] != '"') bad("expected string");
        ++i;
        std::string out;
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\') {
                if (++i >= s.size()) bad("bad escape");
                switch (s[i]) {
                case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                case 'u': {
                    if (i + 4 >= s.size()) bad("bad \\u");
                    unsigned cp = std::stoul(s.substr(i + 1, 4), nullptr, 16);
                    i += 4;
                    if (cp < 0x80) out += char(cp);
                    else if (cp < 0x800) { out += char(0xC0 | (cp >> 6)); out += char(0x80 | (cp & 0x3F)); }
                    else { out += char(0xE0 | (cp >> 12)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
                    break;
                }
                default: out += s[i];
                }
                ++i;
            } else out += s[i++];
        }
        if (i >= s.size()) bad("unterminated string");
        ++i;
        return out;
    }
public:
    explicit JsonParser(const std::string &text) : s(text) {}
    Json value()
    {
        ws();
        if (i >= s.size()) bad("unexpected end");
        Json j;
        const char c = s[i];
        if (c == '{') {
            j.kind = Json::Obj; ++i; ws();
            if (s[i] == '}') { ++i; return j; }
            for (;;) {
                ws(); std::string k = string_(); ws();
                if (s[i] != ':') bad("expected ':'");
                ++i;
                j.obj.emplace_back(std::move(k), value());
                ws();
                if (s[i] == ',') { ++i; continue; }
                if (s[i] == '}') { ++i; break; }
                bad("expected ',' or '}'");
            }
        } else if (c == '[') {
            j.kind = Json::Arr; ++i; ws();
            if (s[i] == ']') { ++i; return j; }
            for (;;) {
                j.arr.push_back(value());
                ws();
                if (s[i] == ',') { ++i; continue; }
                if (s[i] == ']') { ++i; break; }
                bad("expected ',' or ']'");
            }
        } else if (c == '"') { j.kind = Json::Str; j.str = string_(); }
        else if (!s.compare(i, 4, "true")) { j.kind = Json::Bool; j.b = true; i += 4; }
        else if (!s.compare(i, 5, "false")) { j.kind = Json::Bool; j.b = false; i += 5; }
        else if (!s.compare(i, 4, "null")) { j.kind = Json::Null; i += 4; }
        else {
            const size_t st = i;
            while (i < s.size() && (std::isdigit((unsigned char)s[i]) || s[i] == '-' || s[i] == '+' || s[i] == '.' || s[i] == 'e' || s[i] == 'E')) ++i;
            if (i == st) bad("unexpected character");
            j.kind = Json::Num; j.str = s.substr(st, i - st); j.num = std::strtod(j.str.c_str(), nullptr);
        }
        return j;
    }
};

// ---- the reference's data carriers ----
struct QueryWeights {                        // src/rag_engine.rs:1846-1863
    std::optional<float> embedding, lexical, reranker, initial;
};
struct DocumentChunk {                       // :46-59 minus the embedding (device-resident)
    std::string id, document_name, text;
    size_t chunk_index = 0, page_number = 0;
    std::optional<std::string> section;
};
struct SearchResult {                        // :72-100
    std::string text;
    float score = 0;
    std::string document, chunk_id;
    size_t chunk_index = 0, page_number = 0;
    std::optional<std::string> section;
    std::optional<float> embedding_score, lexical_score, initial_score, reranker_score;
    std::optional<double> yes_logprob, no_logprob;
    uint32_t row = 0;
};
struct RerankerCandidate {                   // returned by get_embedding_candidates, :448-457
    std::string chunk_id, document, text;
    size_t page_number = 0;
    std::optional<std::string> section;
    float initial_score = 0;
};

inline rlr_resolved_weights resolve_weights(const QueryWeights *w)      // :1888-1896
{
    rlr_query_weights q{};
    if (w) {
        if (w->embedding) { q.embedding = *w->embedding; q.has |= 1; }
        if (w->lexical) { q.lexical = *w->lexical; q.has |= 2; }
        if (w->reranker) { q.reranker = *w->reranker; q.has |= 4; }
        if (w->initial) { q.initial = *w->initial; q.has |= 8; }
    }
    rlr_resolved_weights out{};
    check(rlr_resolve_weights(w ? &q : nullptr, &out));
    return out;
}

inline std::string sanitize_model_name(const std::string &model)        // :1435-1462
{
    size_t a = 0, b = model.size();
    while (a < b && std::isspace((unsigned char)model[a])) ++a;
    while (b > a && std::isspace((unsigned char)model[b - 1])) --b;
    if (a == b) return "default";
    std::string s;
    bool all_sep = true;
    for (size_t i = a; i < b; ++i) {
        const unsigned char c = model[i];
        const bool ok = (c < 128 && std::isalnum(c)) || c == '-' || c == '_' || c == '.';
        s += ok ? char(c) : '_';
        if (s.back() != '_' && s.back() != '.') all_sep = false;
    }
    return all_sep ? "default" : s;
}
inline std::string get_index_path(const std::string &data_dir, const std::string &model)   // :1465-1468
{
    return data_dir + "/chunks_" + sanitize_model_name(model) + ".json";
}

// LexicalIndex (src/rag_engine.rs:2083-2231) over the library's host-side BM25 index: chunk ids get
// u64 keys in insertion order (also the tie order of equal scores).
class LexicalIndex {
    rlr_lexical *lx_ = nullptr;
    std::unordered_map<std::string, uint64_t> key_of_;
    std::unordered_map<uint64_t, std::string> id_of_;
    uint64_t next_ = 0;

public:
    LexicalIndex() { check_hm(rlr_lexical_create(&lx_)); }
    LexicalIndex(const LexicalIndex &) = delete;
    LexicalIndex &operator=(const LexicalIndex &) = delete;
    ~LexicalIndex() { rlr_lexical_destroy(lx_); }
    void add_chunk(const std::string &id, const std::string &text)
    {
        auto it = key_of_.find(id);
        uint64_t key;
        if (it == key_of_.end()) { key = next_++; key_of_[id] = key; id_of_[key] = id; } else key = it->second;
        check_hm(rlr_lexical_add_chunk(lx_, key, text.data(), text.size()));
    }
    void remove_chunk(const std::string &id)
    {
        auto it = key_of_.find(id);
        if (it == key_of_.end()) return;
        check_hm(rlr_lexical_remove_chunk(lx_, it->second));
        id_of_.erase(it->second);
        key_of_.erase(it);
    }
    bool contains(const std::string &id) const
    {
        auto it = key_of_.find(id);
        if (it == key_of_.end()) return false;
        int out = 0;
        check_hm(rlr_lexical_contains(lx_, it->second, &out));
        return out != 0;
    }
    std::vector<std::pair<std::string, float>> score(const std::string &query, size_t limit) const
    {
        const uint32_t cap = static_cast<uint32_t>(limit > 0 ? limit : std::max<size_t>(key_of_.size(), 1));
        std::vector<uint64_t> keys(cap);
        std::vector<float> sc(cap);
        uint32_t n = 0;
        check_hm(rlr_lexical_score(lx_, query.data(), query.size(), static_cast<uint32_t>(limit), keys.data(), sc.data(), cap, &n));
        std::vector<std::pair<std::string, float>> out;
        for (uint32_t i = 0; i < n; ++i) out.emplace_back(id_of_.at(keys[i]), sc[i]);
        return out;
    }
};

// LexicalIndex with the postings scored ON THE DEVICE (rlr_bm25_*): the host keeps the tokenizer (host-mirror support
// library) and the term -> id dictionary; chunks are identified by their row in the store.
class DeviceLexicalIndex {
    rlr_bm25 *ix_ = nullptr;                 // over one store ...
    rlr_cluster_bm25 *cix_ = nullptr;        // ... or over a cluster (one device index per shard, global statistics)
    std::unordered_map<std::string, uint32_t> vocab_;

    static std::vector<std::string> tokenize(const std::string &text)      // fn tokenize, :2242-2247
    {
        std::string buf(3 * text.size() + 16, '\0');
        size_t n = 0;
        uint32_t nt = 0;
        check_hm(rlr_tokenize(text.data(), text.size(), &buf[0], buf.size(), &n, &nt));
        std::vector<std::string> out;
        size_t a = 0;
        for (size_t i = 0; i <= n && n; ++i)
            if (i == n || buf[i] == '\n') { out.emplace_back(buf.substr(a, i - a)); a = i + 1; }
        return out;
    }

public:
    explicit DeviceLexicalIndex(rlr_store *s) { check(rlr_bm25_create(s, &ix_)); }
    explicit DeviceLexicalIndex(rlr_cluster *c) { check(rlr_cluster_bm25_create(c, &cix_)); }
    DeviceLexicalIndex(const DeviceLexicalIndex &) = delete;
    DeviceLexicalIndex &operator=(const DeviceLexicalIndex &) = delete;
    ~DeviceLexicalIndex() { rlr_bm25_destroy(ix_); rlr_cluster_bm25_destroy(cix_); }
    rlr_bm25 *handle() const { return ix_; }
    rlr_cluster_bm25 *cluster_handle() const { return cix_; }
    void add_chunk(uint32_t row, const std::string &text)
    {
        std::map<std::string, uint32_t> counts;
        for (auto &t : tokenize(text)) ++counts[t];
        std::vector<uint32_t> ids, tfs;
        for (auto &kv : counts) {
            auto it = vocab_.find(kv.first);
            if (it == vocab_.end()) it = vocab_.emplace(kv.first, static_cast<uint32_t>(vocab_.size())).first;
            ids.push_back(it->second); tfs.push_back(kv.second);
        }
        check(cix_ ? rlr_cluster_bm25_set_doc(cix_, row, ids.data(), tfs.data(), static_cast<uint32_t>(ids.size()))
                   : rlr_bm25_set_doc(ix_, row, ids.data(), tfs.data(), static_cast<uint32_t>(ids.size())));
    }
    // the query's known term ids in bytewise order of the term strings (std::map<std::string> iterates in that order)
    std::vector<uint32_t> query_terms(const std::string &query) const
    {
        std::map<std::string, int> uniq;
        for (auto &t : tokenize(query)) uniq[t] = 1;
        std::vector<uint32_t> out;
        for (auto &kv : uniq) { auto it = vocab_.find(kv.first); if (it != vocab_.end()) out.push_back(it->second); }
        return out;
    }
};

class RagEngine {
    std::unique_ptr<LexicalIndex> lexical_;          // built by enable_lexical(): validate_index_sync, :1375-1389
    std::unique_ptr<DeviceLexicalIndex> dev_lexical_; // or by enable_lexical_on_device(): the same index, postings on the GPU
    rlr_store *store_ = nullptr;                     // one GPU ...
    rlr_cluster *cluster_ = nullptr;                 // ... or the same rows sharded over several GPUs, driven from this process
    std::vector<int> devices_;                       // > 1 entry: cluster
    std::vector<DocumentChunk> chunks_;              // row -> chunk
    std::unordered_map<std::string, uint32_t> row_of_;
    uint32_t dim_ = 0;
    int device_ = 0;
    bool needs_reindex_ = false;

    bool normalize_on_upload_ = false;               // sidecar loads: normalize (:1678-1680) runs on the device

    void upload(const std::vector<float> &rows)
    {
        if (store_) { rlr_store_destroy(store_); store_ = nullptr; }
        if (cluster_) { rlr_cluster_destroy(cluster_); cluster_ = nullptr; }
        const uint32_t flags = normalize_on_upload_ ? RLR_STORE_NORMALIZE_ON_UPLOAD : 0;
        if (devices_.size() > 1)
            check(rlr_cluster_create(devices_.data(), static_cast<uint32_t>(devices_.size()), dim_ ? dim_ : 1, chunks_.size(),
                                     rows.empty() ? nullptr : rows.data(), dim_, flags, nullptr, &cluster_));
        else
            check(rlr_store_create(device_, dim_ ? dim_ : 1, chunks_.size(), rows.empty() ? nullptr : rows.data(), dim_, 0, flags, &store_));
        row_of_.clear();
        for (uint32_t i = 0; i < chunks_.size(); ++i) row_of_[chunks_[i].id] = i;
    }
    SearchResult result(uint32_t row, float score, float emb, float lex) const
    {
        const DocumentChunk &c = chunks_.at(row);
        SearchResult r;                              // fallback-fill fields, :680-695
        r.text = c.text; r.score = score; r.document = c.document_name; r.chunk_id = c.id;
        r.chunk_index = c.chunk_index; r.page_number = c.page_number; r.section = c.section;
        r.embedding_score = emb; r.lexical_score = lex; r.initial_score = score; r.row = row;
        return r;
    }

public:
    explicit RagEngine(int device = 0) : device_(device) {}
    // the store sharded over `devices` (entry 0 = root), all driven from this process (rlr_cluster_*)
    explicit RagEngine(std::vector<int> devices) : devices_(std::move(devices)), device_(devices_.empty() ? 0 : devices_[0]) {}
    RagEngine(const RagEngine &) = delete;
    RagEngine &operator=(const RagEngine &) = delete;
    ~RagEngine() { dev_lexical_.reset(); if (store_) rlr_store_destroy(store_); if (cluster_) rlr_cluster_destroy(cluster_); }
    bool sharded() const { return cluster_ != nullptr; }

    size_t len() const { return chunks_.size(); }
    bool needs_reindex() const { return needs_reindex_; }

    // Index every chunk's text (what validate_index_sync does after a load, :1382-1389).
    void enable_lexical()
    {
        lexical_.reset(new LexicalIndex());
        for (auto &c : chunks_) lexical_->add_chunk(c.id, c.text);
    }
    const LexicalIndex *lexical() const { return lexical_.get(); }
    // The same, with the postings on the device: text queries then run BM25, blend, top-k and MMR as one device
    // sequence (rlr_search_text_*); over a cluster every shard scores its documents with the corpus-wide statistics
    // (rlr_cluster_search_text_*).
    void enable_lexical_on_device()
    {
        if (!store_ && !cluster_) throw Error(RLR_ERR_UNSUPPORTED, "device BM25 needs a store");
        if (cluster_) dev_lexical_.reset(new DeviceLexicalIndex(cluster_));
        else dev_lexical_.reset(new DeviceLexicalIndex(store_));
        for (uint32_t i = 0; i < chunks_.size(); ++i) dev_lexical_->add_chunk(i, chunks_[i].text);
    }

    // search / search_with_diversity for a TEXT query whose embedding the caller already has
    // (EmbeddingService is host HTTP, out of scope): runs lexical_index.score(query, 5 * top_k) (:505).
    std::vector<SearchResult> search_text(const std::string &query, const std::vector<float> &query_embedding, size_t top_k,
                                          const QueryWeights *weights = nullptr) const
    {
        const size_t k = std::max<size_t>(top_k, 1);
        return search(query_embedding, top_k, weights, lexical_ ? lexical_->score(query, 5 * k) : std::vector<std::pair<std::string, float>>{});
    }
    std::vector<SearchResult> search_text_with_diversity(const std::string &query, const std::vector<float> &query_embedding,
                                                         size_t top_k, float diversity_factor, const QueryWeights *weights = nullptr) const
    {
        float lam = diversity_factor;
        if (lam < 0.0f) lam = 0.0f;
        if (lam > 1.0f) lam = 1.0f;
        if (dev_lexical_ && !chunks_.empty()) {
            const rlr_resolved_weights w = resolve_weights(weights);
            const std::vector<uint32_t> terms = dev_lexical_->query_terms(query);
            const size_t cap = std::max<size_t>(top_k, 1);
            std::vector<uint32_t> rows(cap); std::vector<float> score(cap), emb(cap), lex(cap);
            uint32_t n = 0;
            check(cluster_ ? rlr_cluster_search_text_mmr(cluster_, dev_lexical_->cluster_handle(), query_embedding.data(),
                                                         static_cast<uint32_t>(query_embedding.size()), 0, static_cast<uint32_t>(top_k),
                                                         diversity_factor, &w, terms.data(), static_cast<uint32_t>(terms.size()),
                                                         rows.data(), score.data(), emb.data(), lex.data(), &n)
                           : rlr_search_text_mmr(store_, dev_lexical_->handle(), query_embedding.data(), static_cast<uint32_t>(query_embedding.size()), 0,
                                                 static_cast<uint32_t>(top_k), diversity_factor, &w, terms.data(), static_cast<uint32_t>(terms.size()),
                                                 rows.data(), score.data(), emb.data(), lex.data(), &n));
            std::vector<SearchResult> out;
            for (uint32_t i = 0; i < n; ++i) out.push_back(result(rows[i], score[i], emb[i], lex[i]));
            return out;
        }
        const size_t pool = lam == 0.0f ? std::max<size_t>(top_k, 1) : std::max<size_t>(3 * top_k, top_k + 10);   // :728-734
        return search_with_diversity(query_embedding, top_k, diversity_factor, weights,
                                     lexical_ ? lexical_->score(query, 5 * pool) : std::vector<std::pair<std::string, float>>{});
    }
    const std::vector<DocumentChunk> &chunks() const { return chunks_; }

    // load_from_disk + apply_loaded_state (:1520-1696) for the model-specific file.
    void load_from_disk(const std::string &data_dir, const std::string &model) { load_file(get_index_path(data_dir, model)); }
    void load_file(const std::string &path)
    {
        std::ifstream f(path, std::ios::binary);
        if (!f) throw Error(RLR_ERR_INVALID_ARG, "cannot open " + path);
        std::stringstream ss;
        ss << f.rdbuf();
        const std::string text = ss.str();
        const Json root = JsonParser(text).value();
        const Json *ver = root.get("version"), *ch = root.get("chunks");
        if (!ver || ver->kind != Json::Num || !ch || ch->kind != Json::Obj) throw Error(RLR_ERR_INVALID_ARG, "not a PersistedState file");
        chunks_.clear();
        std::vector<float> rows;
        dim_ = 0;
        if (ver->num < 2) {                          // :1664-1673 outdated: wipe, mark for reindex
            needs_reindex_ = true;
            upload(rows);
            return;
        }
        for (auto &kv : ch->obj) {
            const Json &c = kv.second;
            const Json *e = c.get("embedding");
            if (!e || e->kind != Json::Arr) throw Error(RLR_ERR_INVALID_ARG, "chunk without embedding");
            if (dim_ == 0) dim_ = static_cast<uint32_t>(e->arr.size());
            if (e->arr.size() != dim_) throw Error(RLR_ERR_DIM_MISMATCH, "mixed embedding dimensions in " + path);
            const size_t base = rows.size();
            for (auto &x : e->arr) rows.push_back(std::strtof(x.str.c_str(), nullptr));
            check(rlr_normalize(rows.data() + base, dim_));          // :1678-1680 re-normalise at load
            DocumentChunk d;
            d.id = c.get("id") && c.get("id")->kind == Json::Str ? c.get("id")->str : kv.first;
            if (auto *v = c.get("document_name")) d.document_name = v->str;
            if (auto *v = c.get("text")) d.text = v->str;
            if (auto *v = c.get("chunk_index")) d.chunk_index = static_cast<size_t>(v->num);
            if (auto *v = c.get("page_number")) d.page_number = static_cast<size_t>(v->num);   // #[serde(default)] -> 0
            if (auto *v = c.get("section")) if (v->kind == Json::Str) d.section = v->str;
            chunks_.push_back(std::move(d));
        }
        const Json *nr = root.get("needs_reindex"), *dh = root.get("document_hashes");
        needs_reindex_ = nr && nr->kind == Json::Bool && nr->b;
        if ((!dh || dh->obj.empty()) && !chunks_.empty()) needs_reindex_ = true;   // :1686-1691
        upload(rows);
    }

    // Binary sidecar `chunks_{model}.rlrbin` written by the Python mirror's save_sidecar:
    // 40-byte header (magic "RLRB200\0", u32 version, u32 dim, u64 n_rows, u64 meta_bytes, u64 reserved), the
    // n x dim f32 rows as stored, then a JSON blob {model, needs_reindex, document_hashes, chunks:[...] in row order}.
    // apply_loaded_state semantics: version gate (:1664), re-normalise every row (:1678, on the device), :1686 rule.
    void load_sidecar(const std::string &path)
    {
        std::ifstream f(path, std::ios::binary);
        if (!f) throw Error(RLR_ERR_INVALID_ARG, "cannot open " + path);
        char head[40];
        f.read(head, 40);
        if (f.gcount() != 40 || std::memcmp(head, "RLRB200\0", 8) != 0) throw Error(RLR_ERR_INVALID_ARG, path + ": not an rlr_b200 sidecar");
        uint32_t version, dim; uint64_t n, meta_bytes;
        std::memcpy(&version, head + 8, 4); std::memcpy(&dim, head + 12, 4);
        std::memcpy(&n, head + 16, 8); std::memcpy(&meta_bytes, head + 24, 8);
        chunks_.clear();
        dim_ = 0;
        std::vector<float> rows;
        if (version < 2) { needs_reindex_ = true; upload(rows); return; }
        rows.resize(n * dim);
        f.read(reinterpret_cast<char *>(rows.data()), static_cast<std::streamsize>(rows.size() * 4));
        std::string meta(meta_bytes, '\0');
        f.read(&meta[0], static_cast<std::streamsize>(meta_bytes));
        if (!f || f.peek() != std::char_traits<char>::eof()) throw Error(RLR_ERR_INVALID_ARG, path + ": truncated or corrupt sidecar");
        const Json root = JsonParser(meta).value();
        const Json *ch = root.get("chunks");
        if (!ch || ch->kind != Json::Arr || ch->arr.size() != n) throw Error(RLR_ERR_INVALID_ARG, path + ": chunk records do not match the row count");
        for (auto &c : ch->arr) {
            DocumentChunk d;
            if (auto *v = c.get("id")) d.id = v->str;
            if (auto *v = c.get("document_name")) d.document_name = v->str;
            if (auto *v = c.get("text")) d.text = v->str;
            if (auto *v = c.get("chunk_index")) d.chunk_index = static_cast<size_t>(v->num);
            if (auto *v = c.get("page_number")) d.page_number = static_cast<size_t>(v->num);
            if (auto *v = c.get("section")) if (v->kind == Json::Str) d.section = v->str;
            chunks_.push_back(std::move(d));
        }
        dim_ = dim;
        const Json *nr = root.get("needs_reindex"), *dh = root.get("document_hashes");
        needs_reindex_ = nr && nr->kind == Json::Bool && nr->b;
        if ((!dh || dh->obj.empty()) && !chunks_.empty()) needs_reindex_ = true;
        normalize_on_upload_ = true;
        upload(rows);
    }

    // rows already in memory (normalised here like insert does, :359)
    void load_rows(std::vector<DocumentChunk> chunks, std::vector<float> rows, uint32_t dim)
    {
        chunks_ = std::move(chunks);
        dim_ = dim;
        for (size_t i = 0; i < chunks_.size(); ++i) check(rlr_normalize(rows.data() + i * dim, dim));
        upload(rows);
    }

    // RagEngine::search, :470-701 (reranker absent).  `lexical`: what LexicalIndex::score returned.
    std::vector<SearchResult> search(const std::vector<float> &query_embedding, size_t top_k, const QueryWeights *weights = nullptr,
                                     const std::vector<std::pair<std::string, float>> &lexical = {}) const
    {
        if (chunks_.empty()) return {};                               // :476-478
        const rlr_resolved_weights w = resolve_weights(weights);      // :481
        top_k = std::max<size_t>(top_k, 1);                           // :490
        if (top_k > RLR_MAX_M) throw Error(RLR_ERR_UNSUPPORTED, "top_k too large");
        std::vector<uint32_t> lr; std::vector<float> ls;
        for (auto &kv : lexical) { auto it = row_of_.find(kv.first); if (it != row_of_.end()) { lr.push_back(it->second); ls.push_back(kv.second); } }
        std::vector<uint32_t> rows(top_k); std::vector<float> comb(top_k), emb(top_k), lex(top_k);
        uint32_t n = 0;
        const uint32_t qd = static_cast<uint32_t>(query_embedding.size()), nl = static_cast<uint32_t>(lr.size()), m = static_cast<uint32_t>(top_k);
        check(cluster_ ? rlr_cluster_search_topm(cluster_, query_embedding.data(), qd, 0, &w, lr.data(), ls.data(), nl, m, rows.data(), comb.data(), emb.data(), lex.data(), &n)
                       : rlr_search_topm(store_, query_embedding.data(), qd, 0, &w, lr.data(), ls.data(), nl, m, rows.data(), comb.data(), emb.data(), lex.data(), &n));
        std::vector<SearchResult> out;
        for (uint32_t i = 0; i < n; ++i) out.push_back(result(rows[i], comb[i], emb[i], lex[i]));
        return out;
    }

    // RagEngine::search_with_diversity, :717-759 (one fused device call)
    std::vector<SearchResult> search_with_diversity(const std::vector<float> &query_embedding, size_t top_k, float diversity_factor,
                                                    const QueryWeights *weights = nullptr,
                                                    const std::vector<std::pair<std::string, float>> &lexical = {}) const
    {
        if (chunks_.empty()) return {};
        const rlr_resolved_weights w = resolve_weights(weights);
        std::vector<uint32_t> lr; std::vector<float> ls;
        for (auto &kv : lexical) { auto it = row_of_.find(kv.first); if (it != row_of_.end()) { lr.push_back(it->second); ls.push_back(kv.second); } }
        const size_t cap = std::max<size_t>(top_k, 1);
        std::vector<uint32_t> rows(cap); std::vector<float> score(cap), emb(cap), lex(cap);
        uint32_t n = 0;
        const uint32_t qd = static_cast<uint32_t>(query_embedding.size()), nl = static_cast<uint32_t>(lr.size()), k = static_cast<uint32_t>(top_k);
        check(cluster_ ? rlr_cluster_search_mmr(cluster_, query_embedding.data(), qd, 0, k, diversity_factor, &w, lr.data(), ls.data(), nl,
                                                rows.data(), score.data(), emb.data(), lex.data(), &n)
                       : rlr_search_mmr(store_, query_embedding.data(), qd, 0, k, diversity_factor, &w, lr.data(), ls.data(), nl,
                                        rows.data(), score.data(), emb.data(), lex.data(), &n));
        std::vector<SearchResult> out;
        for (uint32_t i = 0; i < n; ++i) out.push_back(result(rows[i], score[i], emb[i], lex[i]));
        return out;
    }

    // RagEngine::get_embedding_candidates, :415-461
    std::vector<RerankerCandidate> get_embedding_candidates(const std::vector<float> &query_embedding, size_t count) const
    {
        if (chunks_.empty() || count == 0) return {};
        std::vector<uint32_t> rows(count); std::vector<float> sc(count);
        uint32_t n = 0;
        const uint32_t qd = static_cast<uint32_t>(query_embedding.size());
        check(cluster_ ? rlr_cluster_embedding_candidates(cluster_, query_embedding.data(), qd, 0, static_cast<uint32_t>(count), rows.data(), sc.data(), &n)
                       : rlr_embedding_candidates(store_, query_embedding.data(), qd, 0, static_cast<uint32_t>(count), rows.data(), sc.data(), &n));
        std::vector<RerankerCandidate> out;
        for (uint32_t i = 0; i < n; ++i) {
            const DocumentChunk &c = chunks_.at(rows[i]);
            out.push_back({c.id, c.document_name, c.text, c.page_number, c.section, sc[i]});
        }
        return out;
    }

    // add_document's store update, :347-386: drop the document's chunks, insert the new ones
    void replace_document(const std::string &document_name, std::vector<DocumentChunk> chunks, std::vector<float> embeddings)
    {
        if (cluster_) throw Error(RLR_ERR_UNSUPPORTED, "a sharded store is a bulk-loaded snapshot: reload it, or keep a live index on one GPU");
        std::vector<uint32_t> old;
        for (uint32_t i = 0; i < chunks_.size(); ++i) if (chunks_[i].document_name == document_name) old.push_back(i);
        if (lexical_) {                              // drop_stale + add_chunk, :1379, :382
            for (uint32_t i : old) lexical_->remove_chunk(chunks_[i].id);
            for (auto &c : chunks) lexical_->add_chunk(c.id, c.text);
        }
        if (!old.empty()) {
            std::vector<uint32_t> mf(old.size()), mt(old.size());
            uint64_t nm = 0;
            check(rlr_store_remove_rows(store_, old.data(), old.size(), mf.data(), mt.data(), &nm));
            for (uint64_t i = 0; i < nm; ++i) chunks_[mt[i]] = chunks_[mf[i]];
            chunks_.resize(chunks_.size() - old.size());
        }
        if (!chunks.empty()) {
            if (dim_ == 0) dim_ = static_cast<uint32_t>(embeddings.size() / chunks.size());
            if (!normalize_on_upload_)                   // else the store normalises what it is given
                for (size_t i = 0; i < chunks.size(); ++i) check(rlr_normalize(embeddings.data() + i * dim_, dim_));   // :359
            uint64_t first = 0;
            if (!store_ || chunks_.empty()) {
                chunks_ = std::move(chunks);
                upload(embeddings);
                return;
            }
            check(rlr_store_append(store_, chunks.size(), embeddings.data(), dim_, &first));
            for (auto &c : chunks) chunks_.push_back(std::move(c));
        }
        row_of_.clear();
        for (uint32_t i = 0; i < chunks_.size(); ++i) row_of_[chunks_[i].id] = i;
    }
};

// MCP tool `search_documents`, src/mcp_server.rs:81-110: parameter defaults and clamps
inline std::vector<SearchResult> search_documents(const RagEngine &engine, const std::vector<float> &query_embedding,
                                                  std::optional<size_t> top_k = std::nullopt,
                                                  std::optional<float> diversity_factor = std::nullopt,
                                                  const QueryWeights *weights = nullptr)
{
    const size_t k = std::min<size_t>(top_k.value_or(5), RLR_MAX_TOP_K);     // :85
    float lam = diversity_factor.value_or(0.3f);
    lam = std::min(std::max(lam, 0.0f), 1.0f);                               // :86
    return engine.search_with_diversity(query_embedding, k, lam, weights);
}

} // namespace rlr
