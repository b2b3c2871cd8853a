// engine_cli -- drives include/rlr_engine.hpp (the C++ host mirror of RagEngine) for the pytest suite.
//   engine_cli <chunks_{model}.json> <query.f32> <top_k> <diversity>   [replace <doc> <n_new> <seed> | text <query string>]
// Prints one JSON object: {"n":..,"needs_reindex":..,"results":[{"chunk_id":..,"row":..,"score_bits":..,"emb_bits":..}],
//                          "candidates":[{"chunk_id":..,"score_bits":..}]}
#include <cstdio>
#include <fstream>
#include <iostream>

#include "../../include/rlr_engine.hpp"

static uint32_t bits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

int main(int argc, char **argv)
{
    if (argc < 5) { fprintf(stderr, "usage: engine_cli index.json query.f32 top_k diversity\n"); return 2; }
    try {
        // RLR_CLI_DEVICES="0,1" (or "0,0,0": several shards on one GPU): the store sharded over those GPUs, one process
        std::vector<int> devs;
        if (const char *e = getenv("RLR_CLI_DEVICES"))
            for (const char *p = e; *p;) { devs.push_back(static_cast<int>(strtol(p, const_cast<char **>(&p), 10))); if (*p == ',') ++p; }
        rlr::RagEngine eng(devs.size() > 1 ? devs : std::vector<int>{devs.empty() ? 0 : devs[0]});
        const std::string idx = argv[1];
        if (idx.size() > 7 && idx.substr(idx.size() - 7) == ".rlrbin") eng.load_sidecar(idx);
        else eng.load_file(idx);
        std::ifstream qf(argv[2], std::ios::binary);
        std::vector<char> raw((std::istreambuf_iterator<char>(qf)), std::istreambuf_iterator<char>());
        std::vector<float> q(raw.size() / 4);
        memcpy(q.data(), raw.data(), q.size() * 4);
        const size_t top_k = std::strtoul(argv[3], nullptr, 10);
        const float lam = std::strtof(argv[4], nullptr);
        if (argc >= 9 && std::string(argv[5]) == "replace") {
            // replace_document(doc, n_new chunks with LCG embeddings) -- the pytest side regenerates the same values
            const std::string doc = argv[6];
            const size_t n_new = std::strtoul(argv[7], nullptr, 10);
            uint64_t st = std::strtoull(argv[8], nullptr, 10);
            std::vector<rlr::DocumentChunk> cs(n_new);
            std::vector<float> emb(n_new * q.size());
            for (size_t i = 0; i < n_new; ++i) { cs[i].id = doc + "#new" + std::to_string(i); cs[i].document_name = doc; cs[i].chunk_index = i; }
            for (auto &x : emb) {      // SimpleRng of the reference, src/rag_engine.rs:1781-1796
                st = st * 6364136223846793005ull + 1;
                const uint32_t b = static_cast<uint32_t>(st >> 32);
                x = static_cast<float>(b) / static_cast<float>(UINT32_MAX) * 2.0f - 1.0f;
            }
            eng.replace_document(doc, std::move(cs), std::move(emb));
        }
        const bool text_mode = argc >= 7 && std::string(argv[5]) == "text";
        if (text_mode) {                             // BM25 over the chunk texts (validate_index_sync)
            if (getenv("RLR_CLI_BM25_DEVICE")) eng.enable_lexical_on_device();   // postings scored on the GPU (rlr_bm25_*)
            else eng.enable_lexical();
        }
        auto res = text_mode ? eng.search_text_with_diversity(argv[6], q, top_k, lam) : eng.search_with_diversity(q, top_k, lam);
        auto cand = eng.get_embedding_candidates(q, 7);
        printf("{\"n\":%zu,\"needs_reindex\":%s,\"results\":[", eng.len(), eng.needs_reindex() ? "true" : "false");
        for (size_t i = 0; i < res.size(); ++i)
            printf("%s{\"chunk_id\":\"%s\",\"row\":%u,\"score_bits\":%u,\"emb_bits\":%u,\"lex_bits\":%u,\"page\":%zu}", i ? "," : "", res[i].chunk_id.c_str(),
                   res[i].row, bits(res[i].score), bits(*res[i].embedding_score), bits(*res[i].lexical_score), res[i].page_number);
        printf("],\"candidates\":[");
        for (size_t i = 0; i < cand.size(); ++i)
            printf("%s{\"chunk_id\":\"%s\",\"score_bits\":%u}", i ? "," : "", cand[i].chunk_id.c_str(), bits(cand[i].initial_score));
        printf("]}\n");
        return 0;
    } catch (const rlr::Error &e) {
        fprintf(stderr, "rlr::Error %d: %s\n", e.code, e.what());
        return 10 + e.code;
    }
}
