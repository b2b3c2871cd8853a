"""The C++ host mirror (include/rlr_engine.hpp) above the C ABI: it must compile as plain C++17 with
g++, fail loudly without a GPU, and on a B200 return the oracle's results from a chunks_{model}.json."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rust-local-rag_b200")
F32 = np.float32


def _build_cli(tmp_path):
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import _build
    _build.build()
    _build.build_hostmirror()
    exe = os.path.join(tmp_path, "engine_cli")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cpp", "engine_cli.cpp"),
           "-L" + PKG, "-l:librlr_b200.so", "-l:librlr_hostmirror.so", "-Wl,-rpath," + PKG]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def _write_index(path, n, dim, seed=3, texts=None):
    rng = np.random.default_rng(seed)
    chunks = {}
    for i in range(n):
        cid = f"chunk-{i:05d}"
        chunks[cid] = {"id": cid, "document_name": f"doc{i % 5}.pdf", "text": texts[i] if texts else "téxt \"quoted\"\n",
                       "embedding": [float(x) for x in (rng.standard_normal(dim) * 1.7).astype(F32)],
                       "chunk_index": i, "page_number": 1 + i % 4, "section": None if i % 2 else "Intro",
                       "metadata": {"page_range": None, "sentence_range": None, "section_title": None, "token_count": 1,
                                    "overlap_with_previous": 0}}
    with open(path, "w") as f:
        json.dump({"version": 2, "model": "m", "chunks": chunks, "needs_reindex": False,
                   "document_hashes": {"doc0.pdf": "00"}}, f, indent=2)
    return chunks


def _gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_gpu(), reason="checks the no-device failure mode")
def test_cpp_host_compiles_and_fails_loudly_without_gpu(tmp_path):
    exe = _build_cli(str(tmp_path))
    idx = os.path.join(tmp_path, "chunks_m.json")
    _write_index(idx, 8, 16)
    q = os.path.join(tmp_path, "q.f32")
    np.ones(16, F32).tofile(q)
    res = subprocess.run([exe, idx, q, "5", "0.3"], capture_output=True, text=True)
    assert res.returncode == 12, (res.returncode, res.stderr)            # 10 + RLR_ERR_NO_DEVICE
    assert "no CPU fallback" in res.stderr or "CUDA" in res.stderr


@pytest.mark.gpu
def test_cpp_host_matches_oracle(tmp_path, orc):
    exe = _build_cli(str(tmp_path))
    n, dim = 900, 64
    idx = os.path.join(tmp_path, "chunks_m.json")
    chunks = _write_index(idx, n, dim)
    ids = list(chunks)
    rows = orc.normalize_rows(np.array([c["embedding"] for c in chunks.values()], F32))
    qv = np.random.default_rng(9).standard_normal(dim).astype(F32)
    qp = os.path.join(tmp_path, "q.f32")
    qv.tofile(qp)
    out = json.loads(subprocess.run([exe, idx, qp, "8", "0.5"], capture_output=True, text=True, check=True).stdout)
    ref = orc.search_with_diversity(rows, qv, 8, 0.5, full_sort=True)
    assert out["n"] == n and not out["needs_reindex"]
    assert [r["chunk_id"] for r in out["results"]] == [ids[r] for r in ref[0]]
    assert [r["score_bits"] for r in out["results"]] == ref[1].view(np.uint32).tolist()
    assert [r["emb_bits"] for r in out["results"]] == ref[2].view(np.uint32).tolist()
    assert [r["page"] for r in out["results"]] == [1 + int(r) % 4 for r in ref[0]]
    cr, cs = orc.embedding_candidates(rows, qv, 7)
    assert [c["chunk_id"] for c in out["candidates"]] == [ids[r] for r in cr]
    assert [c["score_bits"] for c in out["candidates"]] == cs.view(np.uint32).tolist()
    # replace_document("doc2.pdf", 37 new chunks): same SimpleRng stream regenerated here
    out = json.loads(subprocess.run([exe, idx, qp, "8", "0.5", "replace", "doc2.pdf", "37", "42"], capture_output=True,
                                    text=True, check=True).stdout)
    st, vals = 42, []
    for _ in range(37 * dim):
        st = (st * 6364136223846793005 + 1) % (1 << 64)
        vals.append(F32(F32(st >> 32) / F32(0xFFFFFFFF) * F32(2.0) - F32(1.0)))
    new_rows = orc.normalize_rows(np.array(vals, F32).reshape(37, dim))
    keep = [i for i in range(n) if chunks[ids[i]]["document_name"] != "doc2.pdf"]
    host = {ids[i]: rows[i] for i in keep}
    host.update({f"doc2.pdf#new{i}": new_rows[i] for i in range(37)})
    assert out["n"] == len(host)
    # row order after swap-removal is the engine's business: compare by chunk id through the oracle on any order
    order = list(host)
    ref = orc.search_with_diversity(np.array([host[k] for k in order], F32), qv, 8, 0.5, full_sort=True)
    assert [r["score_bits"] for r in out["results"]] == ref[1].view(np.uint32).tolist()
    assert sorted(r["chunk_id"] for r in out["results"]) == sorted(order[r] for r in ref[0])


@pytest.mark.gpu
def test_cpp_host_hybrid_text_query(tmp_path, orc):
    """rlr::RagEngine::search_text_with_diversity: BM25 over the chunk texts (rlr::LexicalIndex) blended in the
    scan kernel, against the Python restatement of LexicalIndex + the C search oracle."""
    import random
    from oracle import lexical as olex
    exe = _build_cli(str(tmp_path))
    vocab = "retrieval embedding vector cosine rust tokio server chunk sentence index search query rerank lexical Memory café".split()
    rng = random.Random(5)
    n, dim = 700, 64
    texts = [" ".join(rng.choice(vocab) for _ in range(rng.randint(4, 30))) + ".\n" for _ in range(n)]
    idx = os.path.join(tmp_path, "chunks_m.json")
    chunks = _write_index(idx, n, dim, texts=texts)
    ids = list(chunks)
    rows = orc.normalize_rows(np.array([c["embedding"] for c in chunks.values()], F32))
    qv = np.random.default_rng(11).standard_normal(dim).astype(F32)
    qp = os.path.join(tmp_path, "q.f32")
    qv.tofile(qp)
    ref_idx = olex.LexicalIndex()
    for i, t in enumerate(texts):
        ref_idx.add_chunk(i, t)
    for query, k, lam, on_device in (("memory of the café server", 6, 0.4, False), ("rust tokio", 10, 0.0, False),
                                     ("memory of the café server", 6, 0.4, True), ("rust tokio index", 10, 0.0, True),
                                     ("Memory MEMORY lexical rerank", 100, 0.7, True),
                                     ("memory of the café server", 6, 0.4, "0,0,0"), ("Memory MEMORY lexical rerank", 100, 0.7, "0,0")):
        # on_device: rlr::DeviceLexicalIndex -- the postings scored on the GPU, the whole text query one device sequence;
        # a device list: the same over a sharded store (rlr_cluster_bm25_*, rlr_cluster_search_text_mmr)
        env = dict(os.environ, RLR_CLI_BM25_DEVICE="1") if on_device else dict(os.environ)
        if isinstance(on_device, str):
            env["RLR_CLI_DEVICES"] = on_device
        out = json.loads(subprocess.run([exe, idx, qp, str(k), str(lam), "text", query], capture_output=True, text=True,
                                        check=True, env=env).stdout)
        pool = max(k, 1) if lam == 0.0 else max(3 * k, k + 10)
        pairs = ref_idx.score(query, 5 * pool)
        lr, ls = np.array([p[0] for p in pairs], np.uint32), np.array([p[1] for p in pairs], F32)
        ref = orc.search_with_diversity(rows, qv, k, lam, lex_rows=lr, lex_scores=ls, full_sort=True)
        assert [r["chunk_id"] for r in out["results"]] == [ids[r] for r in ref[0]], query
        assert [r["score_bits"] for r in out["results"]] == ref[1].view(np.uint32).tolist()
        assert [r["lex_bits"] for r in out["results"]] == ref[3].view(np.uint32).tolist()


@pytest.mark.gpu
def test_cpp_host_loads_the_binary_sidecar(tmp_path, orc):
    """rlr::RagEngine::load_sidecar reads what the Python mirror's save_sidecar wrote; rows are re-normalised on the
    device at load (:1678-1680), so results are the oracle's on the twice-normalised rows."""
    from rust_local_rag_b200 import engine
    exe = _build_cli(str(tmp_path))
    n, dim = 500, 40
    idx = os.path.join(tmp_path, "chunks_m.json")
    chunks = _write_index(idx, n, dim, seed=8)
    ids = list(chunks)
    eng = engine.RagEngine.from_chunks_json(idx, model="m")
    side = os.path.join(tmp_path, "chunks_m.rlrbin")
    eng.save_sidecar(side)
    twice = orc.normalize_rows(orc.normalize_rows(np.array([c["embedding"] for c in chunks.values()], F32)))
    qv = np.random.default_rng(12).standard_normal(dim).astype(F32)
    qp = os.path.join(tmp_path, "q.f32")
    qv.tofile(qp)
    out = json.loads(subprocess.run([exe, side, qp, "9", "0.4"], capture_output=True, text=True, check=True).stdout)
    ref = orc.search_with_diversity(twice, qv, 9, 0.4, full_sort=True)
    assert out["n"] == n and not out["needs_reindex"]
    assert [r["chunk_id"] for r in out["results"]] == [ids[r] for r in ref[0]]
    assert [r["score_bits"] for r in out["results"]] == ref[1].view(np.uint32).tolist()
    assert [r["page"] for r in out["results"]] == [1 + int(r) % 4 for r in ref[0]]


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0,0,0", "0,1"])
def test_cpp_host_drives_a_sharded_store_from_one_process(tmp_path, orc, devices):
    """The reference is ONE process (src/main.rs:140-167).  The compiled C++ host mirror with
    RLR_CLI_DEVICES drives a row-sharded store through rlr_cluster_* -- no torchrun, no NCCL, no Python --
    and must print the oracle's unsharded answer, BM25 blend included ("0,1": two GPUs, real peer memory)."""
    if devices == "0,1" and _ngpu() < 2:
        pytest.skip(f"needs 2 GPUs on the box (has {_ngpu()})")
    exe = _build_cli(str(tmp_path))
    n, dim = 2500, 96
    idx = os.path.join(tmp_path, "chunks_m.json")
    chunks = _write_index(idx, n, dim, seed=21)
    ids = list(chunks)
    rows = orc.normalize_rows(np.array([c["embedding"] for c in chunks.values()], F32))
    qv = np.random.default_rng(4).standard_normal(dim).astype(F32)
    qp = os.path.join(tmp_path, "q.f32")
    qv.tofile(qp)
    env = dict(os.environ, RLR_CLI_DEVICES=devices)
    for k, lam in ((8, 0.5), (100, 0.7), (5, 0.0)):
        out = json.loads(subprocess.run([exe, idx, qp, str(k), str(lam)], capture_output=True, text=True, check=True, env=env).stdout)
        ref = orc.search_with_diversity(rows, qv, k, lam, full_sort=True)
        assert out["n"] == n
        assert [r["chunk_id"] for r in out["results"]] == [ids[r] for r in ref[0]], (devices, k, lam)
        assert [r["score_bits"] for r in out["results"]] == ref[1].view(np.uint32).tolist()
        assert [r["emb_bits"] for r in out["results"]] == ref[2].view(np.uint32).tolist()
        R, S = orc.embedding_candidates(rows, qv, 7)
        assert [c["chunk_id"] for c in out["candidates"]] == [ids[r] for r in R]
        assert [c["score_bits"] for c in out["candidates"]] == S.view(np.uint32).tolist()
    # replace_document on a sharded store is refused loudly (a cluster is a bulk-loaded snapshot)
    res = subprocess.run([exe, idx, qp, "5", "0.3", "replace", "doc1.pdf", "3", "7"], capture_output=True, text=True, env=env)
    assert res.returncode == 10 + 6, (res.returncode, res.stderr)       # RLR_ERR_UNSUPPORTED
