"""BASELINE config 1 input: a real `chunks_{model}.json` (version 2) holding N synthetic 768-d chunks.

    python tools/make_config1_index.py <data_dir> [--rows 10000] [--dim 768] [--model nomic-embed-text]

Schema: /root/reference/src/rag_engine.rs:46-59 (DocumentChunk), :35-42 (ChunkMetadata), :1479-1486
(PersistedState).  Embeddings come from the counter-hash generator shared by the oracle and the device
 and are written UN-normalised (scaled by 2.5) so that the re-normalise at load (:1678-1680)
is exercised; chunk text is empty, so BM25 is inert (:2112-2114).  Written with the product's own
save_to_disk twin (rust-local-rag_b200/engine.py:write_chunks_json)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make(data_dir, rows=10_000, dim=768, model="nomic-embed-text", seed=0x5EED0001, scale=2.5):
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import engine
    from oracle import orc
    emb = orc.synth_rows(rows, dim, kind=1, seed=seed, n_clusters=256)
    emb = (emb * np.float32(scale)).astype(np.float32)
    chunks = [engine.DocumentChunk(id=f"{i:08x}-0000-4000-8000-{i:012x}", document_name=f"doc{i % 37}.pdf", text="",
                                   chunk_index=i // 37, page_number=1 + i % 11, section=None,
                                   metadata={"page_range": [1 + i % 11, 1 + i % 11], "sentence_range": [0, 4],
                                             "section_title": None, "token_count": 200, "overlap_with_previous": 2})
              for i in range(rows)]
    path = engine.get_index_path(data_dir, model)
    os.makedirs(data_dir, exist_ok=True)
    engine.write_chunks_json(path, model, chunks, emb, False, {f"doc{j}.pdf": "%064x" % j for j in range(37)})
    return path, emb


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("data_dir")
    ap.add_argument("--rows", type=int, default=10_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--model", default="nomic-embed-text")
    a = ap.parse_args()
    p, _ = make(a.data_dir, a.rows, a.dim, a.model)
    print(p, os.path.getsize(p), "bytes")
