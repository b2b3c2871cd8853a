"""N4 decision data ("BM25 postings scoring on the device"): what the host-side
LexicalIndex::score (/root/reference/src/rag_engine.rs:2169-2225) costs per query on a synthetic corpus, through the
host-mirror twin (librlr_hostmirror.so), next to the bytes the device needs from it.  Host only; run anywhere."""
import ctypes as C
import random
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import rust_local_rag_b200  # noqa
from rust_local_rag_b200 import binding as B

lib = B.load_hostmirror()
rng = random.Random(1)
vocab = [f"w{rng.randrange(10**6):06d}" for _ in range(30000)]
weights = [1.0 / (i + 1) for i in range(len(vocab))]                  # Zipf-ish term frequencies
for n_docs in (10_000, 100_000):
    lx = C.c_void_p()
    B.check_hm(lib.rlr_lexical_create(C.byref(lx)))
    t0 = time.perf_counter()
    for d in range(n_docs):
        text = " ".join(rng.choices(vocab, weights, k=200)).encode()   # ~200 tokens per chunk (:245)
        B.check_hm(lib.rlr_lexical_add_chunk(lx, d, text, len(text)))
    build = time.perf_counter() - t0
    for k in (5, 100):
        limit = 5 * max(3 * k, k + 10)                                 # 5 * pool, :505 via :734
        keys, scores, n = np.zeros(limit, np.uint64), np.zeros(limit, np.float32), C.c_uint32(0)
        lat = []
        for _ in range(200):
            q = " ".join(rng.choices(vocab, weights, k=8)).encode()
            t0 = time.perf_counter()
            B.check_hm(lib.rlr_lexical_score(lx, q, len(q), limit, keys.ctypes.data_as(C.c_void_p), scores.ctypes.data_as(C.c_void_p), limit, C.byref(n)))
            lat.append(time.perf_counter() - t0)
        print(f"{n_docs} chunks x ~200 tokens (index built in {build:.1f}s): score(query of 8 terms, limit {limit}) p50 "
              f"{1e6 * statistics.median(lat):.0f} us, p90 {1e6 * sorted(lat)[180]:.0f} us; output <= {limit} pairs = {limit * 8} bytes to the device")
    lib.rlr_lexical_destroy(lx)
