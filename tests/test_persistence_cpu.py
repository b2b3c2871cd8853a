"""N1, host side: the `chunks_{model}.json` writer (save_to_disk, /root/reference/src/rag_engine.rs:1477-1518) and
the file-selection logic of load_from_disk (:1520-1652), both pure host code and testable without a GPU.  The
reference's own persistence tests (:2365-2667) pin the same facts: file naming, tmp + rename, legacy preserved."""
import json
import os

import numpy as np

F32 = np.float32


def _engine():
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import engine
    return engine


def _chunks(E, n):
    return [E.DocumentChunk(id=f"id-{i}", document_name=f"d{i % 2}.pdf", text=f'text "{i}"\né', chunk_index=i,
                            page_number=1 + i, section="S" if i % 2 else None,
                            metadata={"page_range": [1, 2], "token_count": 7}) for i in range(n)]


def test_writer_schema_roundtrip_and_atomic_rename(tmp_path):
    E = _engine()
    rng = np.random.default_rng(1)
    rows = (rng.standard_normal((4, 33)) * 10.0 ** rng.integers(-20, 20, (4, 33))).astype(F32)
    rows[0, :4] = [0.0, -0.0, 1e-45, 3.4e38]
    path = E.get_index_path(str(tmp_path), "nomic-embed-text:latest")
    assert os.path.basename(path) == "chunks_nomic-embed-text_latest.json"      # sanitize_model_name, :1435-1462
    E.write_chunks_json(path, "nomic-embed-text:latest", _chunks(E, 4), rows, True, {"d0.pdf": "00", "d1.pdf": "11"})
    assert os.listdir(tmp_path) == ["chunks_nomic-embed-text_latest.json"]        # the .json.tmp was renamed away
    text = open(path, encoding="utf-8").read()
    st = json.loads(text)
    assert list(st) == ["version", "model", "chunks", "needs_reindex", "document_hashes"]    # PersistedState field order
    assert st["version"] == 2 and st["model"] == "nomic-embed-text:latest" and st["needs_reindex"] is True
    c = st["chunks"]["id-1"]
    assert list(c) == ["id", "document_name", "text", "embedding", "chunk_index", "page_number", "section", "metadata"]
    assert list(c["metadata"]) == ["page_range", "sentence_range", "section_title", "token_count", "overlap_with_previous"]
    assert c["metadata"] == {"page_range": [1, 2], "sentence_range": None, "section_title": None, "token_count": 7,
                             "overlap_with_previous": 0}
    assert c["text"] == 'text "1"\né' and c["section"] == "S" and st["chunks"]["id-0"]["section"] is None
    for i in range(4):                                                            # every f32 survives the decimal round trip
        assert np.array(st["chunks"][f"id-{i}"]["embedding"], np.float64).astype(F32).tobytes() == rows[i].tobytes()
    assert text.startswith('{\n  "version": 2,\n  "model": ') and '\n      "embedding": [\n        ' in text   # to_string_pretty
    # document_hashes is omitted when empty (skip_serializing_if = "HashMap::is_empty")
    E.write_chunks_json(path, "m", _chunks(E, 1), rows[:1], False, {})
    assert "document_hashes" not in json.load(open(path))


def test_load_decision_follows_the_reference(tmp_path):
    E = _engine()
    d = str(tmp_path)
    model = "nomic-embed-text"
    assert E.decide_load(d, model) == E.LoadDecision(None, None)                 # nothing on disk: start fresh
    legacy = E.get_legacy_path(d)
    assert os.path.basename(legacy) == "chunks.json"
    rows = np.eye(3, 8, dtype=F32)
    # legacy file of ANOTHER model: left alone, start fresh (:1618-1626)
    E.write_chunks_json(legacy, "all-minilm", _chunks(E, 3), rows, False, {"d0.pdf": "00"})
    before = open(legacy, "rb").read()
    dec = E.decide_load(d, model)
    assert dec.state is None and not dec.migrate and not dec.needs_reindex
    assert open(legacy, "rb").read() == before
    # legacy file of THIS model: migrate (:1592-1617)
    E.write_chunks_json(legacy, model, _chunks(E, 3), rows, False, {"d0.pdf": "00"})
    dec = E.decide_load(d, model)
    assert dec.migrate and dec.source == legacy and len(dec.state["chunks"]) == 3
    # a model-specific file has priority over the legacy one (:1549)
    specific = E.get_index_path(d, model)
    E.write_chunks_json(specific, model, _chunks(E, 2), rows[:2], False, {"d0.pdf": "00"})
    dec = E.decide_load(d, model)
    assert dec.source == specific and not dec.migrate and len(dec.state["chunks"]) == 2
    # a corrupt model-specific file is kept and asks for a reindex (:1570-1582)
    open(specific, "w").write('{"version": 2, "model": ')
    dec = E.decide_load(d, model)
    assert dec.state is None and dec.needs_reindex and os.path.exists(specific)
    os.remove(specific)
    # pre-model legacy format (a bare id -> chunk map) with chunks in it: reindex required (:1627-1644)
    json.dump({"a": {"id": "a", "document_name": "x", "text": "", "embedding": [1.0], "chunk_index": 0}}, open(legacy, "w"))
    dec = E.decide_load(d, model)
    assert dec.state is None and dec.needs_reindex
    json.dump({}, open(legacy, "w"))
    assert not E.decide_load(d, model).needs_reindex
