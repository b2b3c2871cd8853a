"""Host mirror of RagEngine over the CUDA path: chunks_{model}.json loading (N1), store mutation
for add_document (N2), and the reranker-in-the-middle flow (N3) -- SURVEY.md section 8(f)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F32 = np.float32


def _write_index(path, n, dim, version=2, hashes=True, seed=0):
    rng = np.random.default_rng(seed)
    chunks = {}
    for i in range(n):
        cid = f"{i:08x}-aaaa-bbbb-cccc-{i:012x}"
        chunks[cid] = {"id": cid, "document_name": f"doc{i % 7}.pdf", "text": "",
                       "embedding": [float(x) for x in (rng.standard_normal(dim) * 2.5).astype(F32)],  # NOT normalised
                       "chunk_index": i, "page_number": 1 + i % 9, "section": None,
                       "metadata": {"page_range": [1, 2], "sentence_range": [0, 3], "section_title": None,
                                    "token_count": 180, "overlap_with_previous": 2}}
    state = {"version": version, "model": "nomic-embed-text", "chunks": chunks, "needs_reindex": False}
    if hashes:
        state["document_hashes"] = {f"doc{j}.pdf": "ab" * 32 for j in range(7)}
    with open(path, "w") as f:
        json.dump(state, f, indent=2)
    return chunks


def test_load_chunks_json_and_search_documents(tmp_path, orc):
    """BASELINE config 1 in miniature: chunks_{model}.json -> re-normalise at load (:1678) -> search_documents."""
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import engine
    n, dim = 600, 96
    path = engine.get_index_path(str(tmp_path), "nomic-embed-text")
    assert os.path.basename(path) == "chunks_nomic-embed-text.json"
    chunks = _write_index(path, n, dim)
    eng = engine.RagEngine.load_from_disk(str(tmp_path), model="nomic-embed-text")
    assert len(eng.chunks) == n and not eng.needs_reindex
    rows = orc.normalize_rows(np.array([c["embedding"] for c in chunks.values()], F32))
    ids = list(chunks.keys())
    q = np.random.default_rng(5).standard_normal(dim).astype(F32)
    hits = engine.search_documents(eng, q)                                   # defaults: top_k=5, diversity 0.3
    ref = orc.search_with_diversity(rows, q, 5, 0.3, full_sort=True)
    assert [h.chunk_id for h in hits] == [ids[r] for r in ref[0]]
    assert np.array([h.score for h in hits], F32).tobytes() == ref[1].tobytes()
    assert all(h.embedding_score is not None and h.reranker_score is None and h.initial_score == h.score for h in hits)
    assert hits[0].document == chunks[ids[ref[0][0]]]["document_name"] and hits[0].page_number == 1 + ref[0][0] % 9
    hits = engine.search_documents(eng, q, top_k=500, diversity_factor=7.0)  # clamps: top_k <= 100, lambda <= 1
    assert len(hits) == 100
    ref = orc.search_with_diversity(rows, q, 100, 1.0, full_sort=True)
    assert [h.row for h in hits] == ref[0].tolist()
    cands = eng.get_embedding_candidates(q, 12)
    rr, ss = orc.embedding_candidates(rows, q, 12)
    assert [c["chunk_id"] for c in cands] == [ids[r] for r in rr]
    assert np.array([c["initial_score"] for c in cands], F32).tobytes() == ss.tobytes()


def test_outdated_or_unfingerprinted_index(tmp_path):
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import engine
    p = os.path.join(tmp_path, "chunks_m.json")
    _write_index(p, 10, 8, version=1)
    e = engine.RagEngine.from_chunks_json(p)                                 # :1664-1673 wipe + needs_reindex
    assert e.chunks == [] and e.needs_reindex and e.search(np.ones(8, F32), 5) == []
    _write_index(p, 10, 8, version=2, hashes=False)
    e = engine.RagEngine.from_chunks_json(p)                                 # :1686-1691
    assert len(e.chunks) == 10 and e.needs_reindex


def test_replace_document_keeps_store_and_table_in_sync(orc):
    """add_document (:347-386): drop the document's rows, append the new ones; results must equal the
    oracle run on the host-side picture of the store after every mutation."""
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import engine
    rng = np.random.default_rng(17)
    dim = 128
    docs = {}

    def make_doc(name, n):
        emb = (rng.standard_normal((n, dim)) * 3).astype(F32)
        return [engine.DocumentChunk(id=f"{name}-{i}-{rng.integers(1 << 30)}", document_name=name, chunk_index=i)
                for i in range(n)], emb

    eng = engine.RagEngine.from_rows(np.zeros((0, dim), F32))
    q = rng.standard_normal(dim).astype(F32)
    for step, (name, n) in enumerate([("a.pdf", 300), ("b.pdf", 1), ("c.pdf", 700), ("a.pdf", 50), ("d.pdf", 2000),
                                      ("c.pdf", 0), ("b.pdf", 129), ("a.pdf", 400)]):
        chunks, emb = make_doc(name, n)
        eng.replace_document(name, chunks, emb)
        docs[name] = (chunks, orc.normalize_rows(emb))
        # host-side picture: whatever order the engine says
        host = {c.id: e for cs, es in docs.values() for c, e in zip(cs, es)}
        assert sorted(c.id for c in eng.chunks) == sorted(host)
        rows = np.array([host[c.id] for c in eng.chunks], F32).reshape(len(eng.chunks), dim)
        assert eng.store.info().n_rows == len(eng.chunks)
        got = eng.search_with_diversity(q, 10, 0.4)
        ref = orc.search_with_diversity(rows, q, 10, 0.4, full_sort=True)
        assert [g.row for g in got] == ref[0].tolist(), step
        assert np.array([g.score for g in got], F32).tobytes() == ref[1].tobytes()
        assert [g.chunk_id for g in got] == [eng.chunks[r].id for r in ref[0]]


def test_reranker_in_the_middle_flow(orc):
    """With a reranker, search() cuts at initial_k = 3*top_k (:544), the host blends reranker and initial
    scores (:602-665), and MMR runs on the blended relevance (:794): rlr_search_topm -> host -> rlr_mmr."""
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import engine
    rng = np.random.default_rng(23)
    n, dim, top_k, lam = 30000, 384, 20, 0.6
    rows = orc.synth_rows(n, dim, kind=1, n_clusters=32)
    q = orc.synth_rows(1, dim, kind=1, seed=77, n_clusters=32)[0]
    s = engine.DeviceStore.from_rows(rows)
    w = engine.resolve_weights(None)
    pool = max(3 * top_k, top_k + 10)
    initial_k = 3 * pool
    r, comb, emb, lex = s.search_topm(q, initial_k, w)
    R, Cc, E, L = orc.search(rows, q, initial_k, threads=4)
    assert r.tobytes() == R.tobytes() and comb.tobytes() == Cc.tobytes()
    # fake LLM relevance in [0,1] per candidate, blended exactly as :618-627
    relev = rng.random(initial_k).astype(F32)
    max_r = max(F32(relev.max()), np.finfo(F32).eps); max_i = max(F32(comb.max()), np.finfo(F32).eps)
    blended = (F32(w.reranker) * (relev / max_r)).astype(F32) + (F32(w.initial) * (comb / max_i)).astype(F32)
    order = np.argsort(-blended.astype(np.float64), kind="stable")[:pool]   # re-sort, truncate(top_k = pool)
    cand_rows, cand_rel = r[order], blended[order].astype(F32)
    sel = s.mmr(cand_rows, cand_rel, top_k, lam)
    ref = orc.mmr(rows[cand_rows], cand_rel, top_k, lam, threads=4)
    assert sel.tobytes() == ref.tobytes()
    s.close()


def test_hybrid_search_with_bm25_index_on_chunk_text(orc):
    """N4: a string query goes through the host BM25 index (limit 5*top_k, :505), its (row, score) pairs
    are normalised by the max (:519-530) and blended in the scan kernel (:531-532).  Oracle side: the
    pure-Python LexicalIndex restatement feeding the C search oracle."""
    import random
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import engine
    from oracle import lexical as olex
    vocab = ("retrieval augmented generation embedding vector cosine similarity rust tokio server pdf chunk sentence "
             "overlap index search query rerank lexical memory bandwidth tensor kernel").split()
    rng, nrng = random.Random(3), np.random.default_rng(3)
    n, dim = 3000, 128
    rows = orc.normalize_rows(nrng.standard_normal((n, dim)).astype(F32))
    chunks = [engine.DocumentChunk(id=f"c{i}", document_name=f"d{i % 11}.pdf",
                                   text=" ".join(rng.choice(vocab) for _ in range(rng.randint(5, 40))), chunk_index=i)
              for i in range(n)]
    store = engine.DeviceStore.from_rows(rows)
    qvec = nrng.standard_normal(dim).astype(F32)
    eng = engine.RagEngine(chunks, store, embedder=lambda s: qvec, lexical="bm25")
    ref_idx = olex.LexicalIndex()
    for i, c in enumerate(chunks):
        ref_idx.add_chunk(i, c.text)
    for query, k, lam in (("memory bandwidth of the tensor kernel", 5, 0.3), ("rust tokio server", 20, 0.0),
                          ("zzz nothing matches", 5, 0.5), ("lexical rerank query", 100, 0.7)):
        pool = max(k, 1) if lam == 0.0 else max(3 * k, k + 10)
        pairs = ref_idx.score(query, 5 * pool)
        lr = np.array([p[0] for p in pairs], np.uint32)
        ls = np.array([p[1] for p in pairs], F32)
        hits = eng.search_with_diversity(query, k, lam)
        R, S, E, L = orc.search_with_diversity(rows, qvec, k, lam, lex_rows=lr if len(lr) else None,
                                               lex_scores=ls if len(ls) else None, full_sort=True)
        assert [h.row for h in hits] == R.tolist(), query
        assert np.array([h.score for h in hits], F32).tobytes() == S.tobytes()
        assert np.array([h.lexical_score for h in hits], F32).tobytes() == L.tobytes()
    # replace_document keeps the lexical index in sync (validate_index_sync, :1375-1389)
    new_chunks = [engine.DocumentChunk(id=f"n{i}", document_name="d3.pdf", text="quantum entanglement " * 3, chunk_index=i)
                  for i in range(4)]
    eng.replace_document("d3.pdf", new_chunks, nrng.standard_normal((4, dim)).astype(F32))
    assert not eng.lexical.contains("c3") and eng.lexical.contains("n0")
    hits = eng.search("quantum entanglement", 4, engine.QueryWeights(embedding=0.0, lexical=1.0))
    assert sorted(h.chunk_id for h in hits) == ["n0", "n1", "n2", "n3"] and all(h.lexical_score == 1.0 for h in hits)


def test_binary_sidecar_round_trip(tmp_path, orc):
    """N1: chunks_{model}.rlrbin carries the same PersistedState as the JSON file.  Loading it re-normalises
    every row like apply_loaded_state does (:1678-1680) -- on the device, and the bits must be the oracle's."""
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import engine
    n, dim = 1500, 200                                                       # dim not a multiple of 32: padded pitch
    path = engine.get_index_path(str(tmp_path), "nomic-embed-text")
    chunks = _write_index(path, n, dim, seed=4)
    eng = engine.RagEngine.load_from_disk(str(tmp_path), model="nomic-embed-text")
    side = engine.get_sidecar_path(str(tmp_path), "nomic-embed-text")
    assert os.path.basename(side) == "chunks_nomic-embed-text.rlrbin"
    eng.save_sidecar(side)
    assert os.path.getsize(side) < os.path.getsize(path) / 3                 # ~4 bytes per float instead of ~20
    eng2 = engine.RagEngine.from_sidecar(side, model="nomic-embed-text")
    assert [c.id for c in eng2.chunks] == [c.id for c in eng.chunks]
    assert eng2.document_hashes == eng.document_hashes and not eng2.needs_reindex
    assert eng2.chunks[7].page_number == eng.chunks[7].page_number and eng2.chunks[7].metadata == eng.chunks[7].metadata
    once = orc.normalize_rows(np.array([c["embedding"] for c in chunks.values()], F32))   # what the JSON load stored
    twice = orc.normalize_rows(once)                                                      # what the next load scans
    assert eng2.store.read_rows(np.arange(n)).tobytes() == twice.tobytes()
    q = np.random.default_rng(6).standard_normal(dim).astype(F32)
    hits = eng2.search_with_diversity(q, 20, 0.6)
    ref = orc.search_with_diversity(twice, q, 20, 0.6, full_sort=True)
    assert [h.row for h in hits] == ref[0].tolist()
    assert np.array([h.score for h in hits], F32).tobytes() == ref[1].tobytes()
    # replace_document on a normalise-on-upload store: rows are normalised exactly once
    new = np.random.default_rng(7).standard_normal((3, dim)).astype(F32) * 3
    eng2.replace_document("doc1.pdf", [engine.DocumentChunk(id=f"x{i}", document_name="doc1.pdf") for i in range(3)], new)
    m = len(eng2.chunks)
    assert eng2.store.read_rows(np.arange(m - 3, m)).tobytes() == orc.normalize_rows(new).tobytes()
    # corrupt / foreign files are refused
    with open(side, "r+b") as f:
        f.write(b"NOTMAGIC")
    with pytest.raises(ValueError):
        engine.RagEngine.from_sidecar(side)


def test_save_to_disk_round_trip_and_legacy_migration(tmp_path, orc):
    """N1 saver: save_to_disk (:1477-1518) writes what load_from_disk reads back bit for bit; a legacy `chunks.json`
    of the current model is migrated to `chunks_{model}.json` and preserved (:1592-1617, :1699-1706); an outdated
    (version < 2) index is wiped, marked for reindex, and the wipe is persisted (:1664-1673)."""
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import engine
    d = str(tmp_path)
    model = "nomic-embed-text"
    legacy = engine.get_legacy_path(d)
    chunks = _write_index(legacy, 300, 64, seed=4)                   # un-normalised embeddings, model nomic-embed-text
    legacy_bytes = open(legacy, "rb").read()
    eng = engine.RagEngine.load_from_disk(d, model=model)
    specific = engine.get_index_path(d, model)
    assert os.path.exists(specific) and open(legacy, "rb").read() == legacy_bytes       # migrated, legacy preserved
    rows = orc.normalize_rows(np.array([c["embedding"] for c in chunks.values()], F32))
    st = json.load(open(specific))
    assert st["version"] == 2 and st["model"] == model and list(st["chunks"]) == list(chunks)
    saved = np.array([c["embedding"] for c in st["chunks"].values()], np.float64).astype(F32)
    assert saved.tobytes() == rows.tobytes()                         # the file holds the NORMALISED rows the searches scan
    assert st["chunks"][eng.chunks[7].id]["metadata"]["token_count"] == 180
    q = np.random.default_rng(2).standard_normal(64).astype(F32)
    a = eng.search_with_diversity(q, 10, 0.4)
    eng2 = engine.RagEngine.load_from_disk(d, model=model)           # now from the model-specific file
    b = eng2.search_with_diversity(q, 10, 0.4)
    ref = orc.search_with_diversity(rows, q, 10, 0.4, full_sort=True)
    ids = list(chunks)
    assert [h.chunk_id for h in a] == [ids[r] for r in ref[0]]
    assert np.array([h.score for h in a], F32).tobytes() == ref[1].tobytes()
    # the reference re-normalises EVERY embedding at EVERY load (:1678-1680), also the already normalised ones it
    # saved: the reloaded engine scans normalize(normalize(x)), which may differ from normalize(x) in the last bit
    rows2 = orc.normalize_rows(rows)
    ref2 = orc.search_with_diversity(rows2, q, 10, 0.4, full_sort=True)
    assert [h.chunk_id for h in b] == [ids[r] for r in ref2[0]]
    assert np.array([h.score for h in b], F32).tobytes() == ref2[1].tobytes()
    # a different model never touches these files and starts fresh
    other = engine.RagEngine.load_from_disk(d, model="all-minilm")
    assert len(other.chunks) == 0 and not other.needs_reindex and not os.path.exists(engine.get_index_path(d, "all-minilm"))
    # outdated index: wiped + needs_reindex, persisted
    d2 = str(tmp_path / "old")
    os.makedirs(d2)
    _write_index(engine.get_index_path(d2, model), 20, 16, version=1)
    e3 = engine.RagEngine.load_from_disk(d2, model=model)
    assert len(e3.chunks) == 0 and e3.needs_reindex
    st = json.load(open(engine.get_index_path(d2, model)))
    assert st["version"] == 2 and st["chunks"] == {} and st["needs_reindex"] is True and "document_hashes" not in st


@pytest.mark.parametrize("devices", [None, [0, 0]])
def test_config1_10k_x_768_through_the_index_file(tmp_path, orc, devices):
    """BASELINE configs[0] at its stated size, end to end through the file format: a generated
    10,000 x 768 `chunks_nomic-embed-text.json` -> load_from_disk (re-normalise, :1678) -> search_documents(top_k=5,
    diversity=0.3) -> the oracle's answer on the same file contents.  `devices=[0, 0]`: the same through a
    two-shard cluster."""
    import sys
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import engine
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import make_config1_index
    path, emb = make_config1_index.make(str(tmp_path), rows=10_000, dim=768)
    assert os.path.basename(path) == "chunks_nomic-embed-text.json" and os.path.getsize(path) > 50_000_000
    eng = engine.RagEngine.load_from_disk(str(tmp_path), model="nomic-embed-text", devices=devices)
    assert len(eng.chunks) == 10_000 and not eng.needs_reindex
    rows = orc.normalize_rows(emb)                                   # what apply_loaded_state leaves in memory
    qs = orc.synth_rows(16, 768, kind=1, seed=0x5EED0002, n_clusters=256)
    for q in qs:
        hits = engine.search_documents(eng, q)                       # top_k = 5, diversity_factor = 0.3
        ref = orc.search_with_diversity(rows, q, 5, 0.3, full_sort=True)
        assert [h.row for h in hits] == ref[0].tolist()
        assert np.array([h.score for h in hits], F32).tobytes() == ref[1].tobytes()
        assert np.array([h.embedding_score for h in hits], F32).tobytes() == ref[2].tobytes()
        assert [h.chunk_id for h in hits] == [eng.chunks[r].id for r in ref[0]]
    hits = engine.search_documents(eng, qs[0], top_k=100, diversity_factor=0.7)
    ref = orc.search_with_diversity(rows, qs[0], 100, 0.7, full_sort=True)
    assert [h.row for h in hits] == ref[0].tolist()
    eng.store.close()
