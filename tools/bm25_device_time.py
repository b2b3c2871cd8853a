"""N4 measurement: LexicalIndex::score on the device (rlr_bm25_*) vs on the host (the hash-map twin), and a whole text
query (BM25 + blend + top-5 + MMR) both ways, on the corpus of tools/bm25_cost.py.  usage: bm25_device_time.py [n_docs]"""
import random
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import rust_local_rag_b200  # noqa
from rust_local_rag_b200 import binding as B, engine

n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
dim = 768
rng = random.Random(1)
vocab = [f"w{rng.randrange(10**6):06d}" for _ in range(30000)]
weights = [1.0 / (i + 1) for i in range(len(vocab))]
docs = [" ".join(rng.choices(vocab, weights, k=200)) for _ in range(n_docs)]
queries = [" ".join(rng.choices(vocab, weights, k=8)) for _ in range(64)]
store = engine.DeviceStore.synthetic(n_docs, dim, kind=1, n_clusters=256)
qv = np.random.default_rng(0).standard_normal((64, dim)).astype(np.float32)
w = engine.ResolvedWeights(np.float32(0.7), np.float32(0.3), np.float32(0.7), np.float32(0.3))
t0 = time.perf_counter()
dev = engine.DeviceLexicalIndex(store)
for i, d in enumerate(docs):
    dev.add_chunk(i, d)
t_dev = time.perf_counter() - t0
t0 = time.perf_counter()
host = engine.LexicalIndex()
for i, d in enumerate(docs):
    host.add_chunk(str(i), d)
t_host = time.perf_counter() - t0
print(f"{n_docs} chunks x 200 tokens: index build {t_dev:.1f}s (device index, host tokenizer) / {t_host:.1f}s (host twin)")


def p50(f, n=200):
    for i in range(5):
        f(i)
    lat = []
    for i in range(n):
        t0 = time.perf_counter()
        f(i)
        lat.append(time.perf_counter() - t0)
    return 1e6 * statistics.median(lat)


for k in (5, 100):
    pool = max(3 * k, k + 10)
    limit = 5 * pool
    terms = [dev.query_terms(q) for q in queries]
    a = p50(lambda i: dev.score(queries[i % 64], limit))
    b = p50(lambda i: host.score(queries[i % 64], limit), n=30)
    c = p50(lambda i: store.search_text_mmr(qv[i % 64], k, 0.3 if k == 5 else 0.7, w, dev.handle, terms[i % 64]))

    def host_query(i):
        pairs = host.score(queries[i % 64], limit)
        lr = np.array([int(r) for r, _ in pairs], np.uint32); ls = np.array([s for _, s in pairs], np.float32)
        return store.search_mmr(qv[i % 64], k, 0.3 if k == 5 else 0.7, w, lr, ls)
    d = p50(host_query, n=30)
    same = all(np.array_equal(store.search_text_mmr(qv[i], k, 0.3, w, dev.handle, terms[i])[0], host_query(i)[0]) for i in range(8)) if k == 5 else None
    print(f"top_k={k}: LexicalIndex::score(limit {limit}) p50 {a:.0f} us on the device vs {b:.0f} us on the host; "
          f"whole text query (BM25 + blend + top-k + MMR) p50 {c:.0f} us vs {d:.0f} us"
          + (f"; identical rows: {same}" if same is not None else ""))
