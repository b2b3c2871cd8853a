"""What LexicalIndex::score (/root/reference/src/rag_engine.rs:2169-2225) costs on the HOST at the corpus shape of
tools/cluster_text_time.py (Zipf vocabulary of 50000, 60 tokens per chunk, 8-term queries drawn from the same term
ranks), through the host-mirror twin (librlr_hostmirror.so: hash-map postings like the reference's).  Host only.

    python tools/bm25_host_cost_large.py [n_docs]        (default 1,000,000; the cost is linear in n_docs)"""
import ctypes as C
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import rust_local_rag_b200  # noqa
from rust_local_rag_b200 import binding as B

n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
V, T = 50_000, 60
lib = B.load_hostmirror()
rng = np.random.default_rng(7)
p = 1.0 / np.arange(1, V + 1)
cdf = np.cumsum(p / p.sum())
words = [f"w{i:05d}" for i in range(V)]
lx = C.c_void_p()
B.check_hm(lib.rlr_lexical_create(C.byref(lx)))
t0 = time.perf_counter()
CH = 50_000
for r0 in range(0, n_docs, CH):
    c = min(CH, n_docs - r0)
    ids = np.minimum(np.searchsorted(cdf, rng.random((c, T))), V - 1)
    for d in range(c):
        text = " ".join([words[i] for i in ids[d]]).encode()
        B.check_hm(lib.rlr_lexical_add_chunk(lx, r0 + d, text, len(text)))
build = time.perf_counter() - t0
limit = 1500
keys, scores, n = np.zeros(limit, np.uint64), np.zeros(limit, np.float32), C.c_uint32(0)
lat = []
ranks = [3, 17, 120, 450, 2000, 9000, 20000, 40000, 11, 64, 300, 5000]
for _ in range(20):
    q = " ".join(words[i] for i in rng.choice(ranks, 8, replace=False)).encode()
    t0 = time.perf_counter()
    B.check_hm(lib.rlr_lexical_score(lx, q, len(q), limit, keys.ctypes.data_as(C.c_void_p), scores.ctypes.data_as(C.c_void_p), limit, C.byref(n)))
    lat.append(time.perf_counter() - t0)
print(f"{n_docs} chunks x {T} tokens, vocabulary {V} (index built in {build:.0f}s): host LexicalIndex::score(8-term query, limit {limit}) "
      f"p50 {1e3 * statistics.median(lat):.1f} ms, min {1e3 * min(lat):.1f} ms, max {1e3 * max(lat):.1f} ms ({n.value} results)")
lib.rlr_lexical_destroy(lx)
