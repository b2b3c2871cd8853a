"""LexicalIndex::score (/root/reference/src/rag_engine.rs:2169-2227) ON THE DEVICE (rlr_bm25_*) and
the text-query entry points that chain it into the blend (rlr_search_text_topm / _mmr).  Oracle: oracle/lexical.py (the
pure-Python restatement; the reference holds no BM25 test, so it is pinned by source text) for the BM25 scores, and
oracle.search / search_with_diversity fed with the oracle's BM25 pairs for the search results.  Bar: bit-exact scores,
identical rows (ties: lower row / lower key, the same deterministic choice on both sides)."""
import random

import numpy as np
import pytest

from oracle import lexical as olex

pytestmark = pytest.mark.gpu
F32 = np.float32

VOCAB = ("retrieval augmented generation embedding vector cosine similarity rust tokio axum server pdf chunk sentence "
         "overlap index search query rerank lexical bm25 the and of to in a is it on at by an be GPU kernel HBM "
         "bandwidth tensor Memory memory MEMORY naïve café straße Ελλάδα москва 2024 42 x86 ΟΔΥΣΣΕΥΣ İstanbul").split()
QUERIES = ["memory bandwidth of the GPU kernel", "rust tokio axum server", "zzz unknown words", "the and the and",
           "Memory MEMORY memory", "café naïve москва straße", "", "a an it", "retrieval augmented generation with bm25 rerank"]


def W(e=0.7, l=0.3):
    from rust_local_rag_b200.engine import ResolvedWeights
    return ResolvedWeights(F32(e), F32(l), F32(0.7), F32(0.3))


def same(a, b):
    return np.asarray(a).tobytes() == np.asarray(b).tobytes()


def _corpus(seed, n_docs, zipf=False):
    rng = random.Random(seed)
    weights = [1.0 / (i + 1) for i in range(len(VOCAB))] if zipf else None
    return [" ".join(rng.choices(VOCAB, weights, k=rng.randint(0, 60))) for _ in range(n_docs)]


@pytest.fixture(scope="module")
def eng():
    from rust_local_rag_b200 import engine
    return engine


@pytest.fixture(autouse=True, params=["forward", "inverted"])
def index_kind(request, monkeypatch):
    """Every test runs over both device layouts: the forward index (CSR by row, small stores) and the inverted index
    (CSR by term, one launch per query term; stores beyond 262144 rows).  The knob is read at each index (re)build."""
    monkeypatch.setenv("RLR_BM25_INDEX", request.param)
    return request.param


def _build(eng, rows, docs, flags=0):
    store = eng.DeviceStore.from_rows(rows, flags=flags)
    ix = eng.DeviceLexicalIndex(store)
    ref = olex.LexicalIndex()
    for i, d in enumerate(docs):
        ix.add_chunk(i, d)
        ref.add_chunk(i, d)
    return store, ix, ref


@pytest.mark.parametrize("n,zipf", [(300, False), (5000, True), (40000, True)])
def test_device_bm25_scores_bit_identical_to_restatement(eng, orc, n, zipf):
    docs = _corpus(n, n, zipf)
    rows = orc.synth_rows(n, 64, kind=0, seed=n)
    store, ix, ref = _build(eng, rows, docs)
    assert ix.stats() == (ref.total_docs, ref.total_length, len(ref.term_postings))
    for query in QUERIES:
        for limit in (1, 25, 75, 1500, 4500):
            got, want = ix.score(query, limit), ref.score(query, limit)
            assert [r for r, _ in got] == [k for k, _ in want], (n, query, limit)
            assert all(F32(a).tobytes() == F32(b).tobytes() for (_, a), (_, b) in zip(got, want)), (n, query, limit)
    ix.close(); store.close()


@pytest.mark.parametrize("flags_name", ["latency-path", "regular-path"])
def test_text_search_equals_oracle_with_oracle_bm25_pairs(eng, rlr, orc, flags_name):
    """search / search_with_diversity for a text query, everything on the device, against the oracle's search fed with
    the oracle's BM25 pairs (lexical_index.score(query, 5 * top_k), :505)."""
    n, dim = 6000, 384
    docs = _corpus(7, n, zipf=True)
    rows = orc.synth_rows(n, dim, kind=1, n_clusters=64)
    store, ix, ref = _build(eng, rows, docs, flags=0 if flags_name == "latency-path" else rlr.RLR_STORE_NO_LATENCY_PATH)
    qs = orc.synth_rows(len(QUERIES), dim, kind=1, seed=0x5EED0002, n_clusters=64)
    blended = False
    for query, q in zip(QUERIES, qs):
        terms = ix.query_terms(query)
        for k, lam in ((5, 0.3), (100, 0.7), (10, 0.0), (0, 0.5)):
            pool = max(k, 1) if lam == 0.0 else max(3 * k, k + 10)
            pairs = ref.score(query, 5 * pool)
            lr = np.array([r for r, _ in pairs], np.uint32); ls = np.array([s for _, s in pairs], F32)
            got = store.search_text_mmr(q, k, lam, W(), ix.handle, terms)
            want = orc.search_with_diversity(rows, q, k, lam, lex_rows=lr if len(lr) else None, lex_scores=ls if len(lr) else None, full_sort=True)
            for a, b in zip(got, want):
                assert same(a, b), (flags_name, query, k, lam)
            blended |= bool((want[3] != 0).any())
        for m in (15, 45, 900):
            pairs = ref.score(query, 5 * m)
            lr = np.array([r for r, _ in pairs], np.uint32); ls = np.array([s for _, s in pairs], F32)
            got = store.search_text_topm(q, m, W(), ix.handle, terms)
            want = orc.search(rows, q, m, lex_rows=lr if len(lr) else None, lex_scores=ls if len(lr) else None, full_sort=True)
            for a, b in zip(got, want):
                assert same(a, b), (flags_name, query, m)
    assert blended, "no BM25 term ever reached a result"
    ix.close(); store.close()


def test_device_bm25_follows_mutation(eng, orc):
    """add_chunk replaces (:2107-2109), remove_chunk (:2140-2167), and the index follows the store's row moves."""
    rng = random.Random(3)
    n = 400
    docs = _corpus(11, n)
    rows = orc.synth_rows(n, 64, kind=0)
    store, ix, ref = _build(eng, rows, docs)
    keys = list(range(n))                                  # ref key of the document currently at each row
    for step in range(200):
        r = rng.randrange(n)
        if rng.random() < 0.5:
            d = rng.choice(docs)
            ix.add_chunk(r, d); ref.add_chunk(keys[r], d)
        else:
            ix.remove_chunk(r); ref.remove_chunk(keys[r])
        if step % 40 == 39:
            assert ix.stats() == (ref.total_docs, ref.total_length, len(ref.term_postings))
            got, want = ix.score("embedding search memory kernel", 50), ref.score("embedding search memory kernel", 50)
            assert sorted((keys[r], F32(s).tobytes()) for r, s in got) == sorted((k, F32(s).tobytes()) for k, s in want)
    # remove rows from the store: the tail moves into the holes and the index follows
    gone = sorted(rng.sample(range(n), 60))
    for r in gone:
        ix.remove_chunk(r); ref.remove_chunk(keys[r])
    mf, mt = store.remove_rows(gone)
    for f, t in zip(mf.tolist(), mt.tolist()):
        ix.move(f, t)
        keys[t] = keys[f]
    got, want = ix.score("the memory of the kernel", 1500), ref.score("the memory of the kernel", 1500)
    assert len(got) == len(want) and max(r for r, _ in got) < n - 60
    assert sorted((keys[r], F32(s).tobytes()) for r, s in got) == sorted((k, F32(s).tobytes()) for k, s in want)
    ix.close(); store.close()


def test_engine_text_queries_with_device_bm25(eng, orc):
    """The host mirror with lexical="bm25-device": RagEngine.search_with_diversity("text") == the same engine with the
    host-side twin (lexical="bm25"), which is checked against the oracle elsewhere."""
    n, dim = 3000, 128
    docs = _corpus(5, n, zipf=True)
    rows = orc.synth_rows(n, dim, kind=1, n_clusters=32)
    qv = {q: v for q, v in zip(QUERIES, orc.synth_rows(len(QUERIES), dim, kind=1, seed=0x5EED0002, n_clusters=32))}
    chunks = [eng.DocumentChunk(id=f"c{i}", document_name="d.pdf", text=docs[i], chunk_index=i) for i in range(n)]
    a = eng.RagEngine(chunks, eng.DeviceStore.from_rows(rows), embedder=lambda s: qv[s], lexical="bm25-device")
    b = eng.RagEngine(chunks, eng.DeviceStore.from_rows(rows), embedder=lambda s: qv[s], lexical="bm25")
    for query in QUERIES:
        for k, lam in ((5, 0.3), (20, 0.0)):
            ra, rb = a.search_with_diversity(query, k, lam), b.search_with_diversity(query, k, lam)
            assert [(x.row, F32(x.score).tobytes(), F32(x.lexical_score).tobytes()) for x in ra] == \
                   [(x.row, F32(x.score).tobytes(), F32(x.lexical_score).tobytes()) for x in rb], (query, k, lam)
        ra, rb = a.search(query, 30), b.search(query, 30)
        assert [(x.row, F32(x.score).tobytes()) for x in ra] == [(x.row, F32(x.score).tobytes()) for x in rb], query


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _cluster_device_sets():
    sets = [pytest.param([0, 0, 0], id="3-shards-on-gpu0"), pytest.param([0], id="1-shard")]
    for g in (2, 4, 8):
        sets.append(pytest.param(list(range(g)), id=f"{g}-gpus",
                                 marks=pytest.mark.skipif(_ngpu() < g, reason=f"needs {g} GPUs on the box (has {_ngpu()})")))
    return sets


@pytest.mark.parametrize("devices", _cluster_device_sets())
def test_cluster_bm25_equals_one_index_over_all_rows(eng, orc, devices):
    """rlr_cluster_bm25_*: one device index per shard, scored with the statistics of the whole corpus, ranked lists merged
    on the host.  Scores and order must not depend on the sharding: bit-equal to the pure-Python restatement over all
    documents, and the text search over the cluster bit-equal to the oracle's search fed with those pairs."""
    n, dim = 9000, 256
    docs = _corpus(11, n, zipf=True)
    rows = orc.synth_rows(n, dim, kind=1, n_clusters=64)
    g = len(devices)
    plan = None if g == 1 else [n - (g - 1) * (n // g + 7)] + [n // g + 7] * (g - 1)       # uneven shards
    cl = eng.ClusterStore.from_rows(rows, devices=devices, shard_rows=plan)
    ix = eng.DeviceLexicalIndex(cl)
    ref = olex.LexicalIndex()
    for i, d in enumerate(docs):
        ix.add_chunk(i, d)
        ref.add_chunk(i, d)
    assert ix.stats() == (ref.total_docs, ref.total_length, len(ref.term_postings))
    for query in QUERIES:
        for limit in (1, 75, 1500):
            got, want = ix.score(query, limit), ref.score(query, limit)
            assert [r for r, _ in got] == [k for k, _ in want], (devices, query, limit)
            assert all(F32(a).tobytes() == F32(b).tobytes() for (_, a), (_, b) in zip(got, want)), (devices, query, limit)
    qs = orc.synth_rows(len(QUERIES), dim, kind=1, seed=0x5EED0002, n_clusters=64)
    blended = False
    for query, q in zip(QUERIES, qs):
        terms = ix.query_terms(query)
        for k, lam in ((5, 0.3), (100, 0.7), (10, 0.0)):
            pool = max(k, 1) if lam == 0.0 else max(3 * k, k + 10)
            pairs = ref.score(query, 5 * pool)
            lr = np.array([r for r, _ in pairs], np.uint32); ls = np.array([s for _, s in pairs], F32)
            got = cl.search_text_mmr(q, k, lam, W(), ix.handle, terms)
            want = orc.search_with_diversity(rows, q, k, lam, lex_rows=lr if len(lr) else None, lex_scores=ls if len(lr) else None, full_sort=True)
            for a, b in zip(got, want):
                assert same(a, b), (devices, query, k, lam)
            blended |= bool((want[3] != 0).any())
        pairs = ref.score(query, 5 * 45)
        lr = np.array([r for r, _ in pairs], np.uint32); ls = np.array([s for _, s in pairs], F32)
        got = cl.search_text_topm(q, 45, W(), ix.handle, terms)
        want = orc.search(rows, q, 45, lex_rows=lr if len(lr) else None, lex_scores=ls if len(lr) else None, full_sort=True)
        for a, b in zip(got, want):
            assert same(a, b), (devices, query)
    assert blended
    # mutation: documents replaced / removed on two different shards change the GLOBAL statistics for every shard
    for r in (3, n - 5):
        ix.add_chunk(r, "bandwidth bandwidth bandwidth HBM kernel")
        ref.add_chunk(r, "bandwidth bandwidth bandwidth HBM kernel")
    ix.remove_chunk(n // 2); ref.remove_chunk(n // 2)
    assert ix.stats() == (ref.total_docs, ref.total_length, len(ref.term_postings))
    got, want = ix.score("HBM bandwidth kernel", 75), ref.score("HBM bandwidth kernel", 75)
    assert [r for r, _ in got] == [k for k, _ in want]
    assert all(F32(a).tobytes() == F32(b).tobytes() for (_, a), (_, b) in zip(got, want))
    ix.close(); cl.close()


def test_engine_text_queries_over_a_cluster_with_device_bm25(eng, orc):
    """RagEngine(lexical="bm25-device") over a ClusterStore == the same engine over one store."""
    n, dim = 4000, 128
    docs = _corpus(5, n, zipf=True)
    rows = orc.synth_rows(n, dim, kind=1, n_clusters=32)
    qv = {q: v for q, v in zip(QUERIES, orc.synth_rows(len(QUERIES), dim, kind=1, seed=0x5EED0002, n_clusters=32))}
    chunks = [eng.DocumentChunk(id=f"c{i}", document_name="d.pdf", text=docs[i], chunk_index=i) for i in range(n)]
    one = eng.RagEngine(chunks, eng.DeviceStore.from_rows(rows), embedder=lambda s: qv[s], lexical="bm25-device")
    many = eng.RagEngine(chunks, eng.ClusterStore.from_rows(rows, devices=[0, 0]), embedder=lambda s: qv[s], lexical="bm25-device")
    for query in QUERIES:
        for k, lam in ((5, 0.3), (20, 0.0)):
            ra, rb = one.search_with_diversity(query, k, lam), many.search_with_diversity(query, k, lam)
            assert [(x.row, F32(x.score).tobytes(), F32(x.lexical_score).tobytes()) for x in ra] == \
                   [(x.row, F32(x.score).tobytes(), F32(x.lexical_score).tobytes()) for x in rb], (query, k, lam)
        ra, rb = one.search(query, 30), many.search(query, 30)
        assert [(x.row, F32(x.score).tobytes()) for x in ra] == [(x.row, F32(x.score).tobytes()) for x in rb], query
