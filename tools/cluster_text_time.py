"""Dev tool: a TEXT query (`search_documents`: BM25 + embedding blend + top-k + MMR) over a store sharded across GPUs,
with the BM25 postings scored on every shard's GPU (rlr_cluster_bm25_*, rlr_cluster_search_text_mmr).

    python tools/cluster_text_time.py [n_docs] [n_gpus] [dim] [shards_on_gpu0]

Documents are synthetic term-id bags (Zipf vocabulary), the store is the bench's clustered synthetic rows.  Checks, at
full size: every (row, score) LexicalIndex::score returns is re-scored on the host with the reference's f32 formula in
the same term order (bit-equal), the list is in rank order, and no document of a large random sample outranks the last
returned one.  Prints one JSON line."""
import ctypes
import json
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import rust_local_rag_b200  # noqa
from rust_local_rag_b200 import binding as B, engine

F = np.float32
n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n_gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 768
shards0 = int(sys.argv[4]) if len(sys.argv) > 4 else 0          # >0: that many shards, all on GPU 0 (development)
devices = [0] * shards0 if shards0 else list(range(n_gpus))
V, T = 50_000, 60                                               # vocabulary, tokens per document
TOP_K, LAM = 100, 0.7

t0 = time.time()
cl = engine.ClusterStore.synthetic(n_docs, dim, devices=devices, kind=B.RLR_SYNTH_CLUSTERED, seed=0x5EED0001, centroid_seed=0x5EED00C0,
                                   n_clusters=4096, sigma=0.65)
ix = engine.DeviceLexicalIndex(cl)
rng = np.random.default_rng(7)
p = 1.0 / np.arange(1, V + 1)
cdf = np.cumsum(p / p.sum())
doc_off = [np.zeros(1, np.uint64)]
doc_ids, doc_tfs = [], []
CH = 200_000
base = 0
blk = None
for r0 in range(0, n_docs, CH):
    c = min(CH, n_docs - r0)
    if blk is None or c != CH:
        # one block of CH documents is generated once and repeated over the row ranges: the statistics are those of
        # n_docs independent documents, the build stays short, and equal documents on DIFFERENT shards tie exactly --
        # the cross-shard tie order (lower row first) is part of what is checked below
        a = np.searchsorted(cdf, rng.random((c, T))).astype(np.uint32)
        np.minimum(a, V - 1, out=a)
        a.sort(axis=1)
        first = np.ones(a.shape, bool)
        first[:, 1:] = a[:, 1:] != a[:, :-1]
        ids = a[first]
        pos = np.flatnonzero(first.ravel())
        tfs = np.diff(np.append(pos, c * T)).astype(np.uint32)
        off = np.concatenate([[0], np.cumsum(first.sum(1))]).astype(np.uint64)
        if c == CH:
            blk = (ids, tfs, off)
    else:
        ids, tfs, off = blk
    ix.add_documents_csr(r0, off, ids, tfs)
    doc_off.append(off[1:] + np.uint64(base))
    base += int(off[-1])
    doc_ids.append(ids); doc_tfs.append(tfs)
doc_off = np.concatenate(doc_off); doc_ids = np.concatenate(doc_ids); doc_tfs = np.concatenate(doc_tfs)
total_docs, total_len, n_terms = ix.stats()
t_build = time.time() - t0

w = engine.ResolvedWeights(F(0.7), F(0.3), F(0.7), F(0.3))
qs = engine.DeviceStore.synthetic(16, dim, kind=B.RLR_SYNTH_CLUSTERED, seed=0x5EED0002, centroid_seed=0x5EED00C0, n_clusters=4096, sigma=0.65)
qh = qs.read_rows(np.arange(16)); qs.close()
term_sets = [np.array(sorted(rng.choice([3, 17, 120, 450, 2000, 9000, 20000, 40000, 11, 64, 300, 5000], 8, replace=False)), np.uint32) for _ in range(16)]

t0 = time.time()
first_call = cl.search_text_mmr(qh[0], TOP_K, LAM, w, ix.handle, term_sets[0])      # builds the device CSR of every shard
t_sync = time.time() - t0


def score_ids(terms, limit):
    rows, sc, n = np.zeros(limit, np.uint32), np.zeros(limit, np.float32), ctypes.c_uint32(0)
    B.check(ix._fn("score")(ix.handle, B.ptr(terms), len(terms), limit, B.ptr(rows), B.ptr(sc), limit, ctypes.byref(n)))
    return rows[:n.value], sc[:n.value]


libm = ctypes.CDLL("libm.so.6"); libm.logf.restype = ctypes.c_float; libm.logf.argtypes = [ctypes.c_float]
df = np.bincount(doc_ids, minlength=V)
doc_len_all = np.add.reduceat(doc_tfs.astype(np.int64), doc_off[:-1].astype(np.int64))


def host_scores(terms, rows):
    """LexicalIndex::score's formula (:2188-2219) for `rows`, numpy f32, terms in the given order."""
    avg = F(total_len) / F(total_docs)
    k1, b = F(1.5), F(0.75)
    acc = np.zeros(len(rows), F); hit = np.zeros(len(rows), bool)
    dl = doc_len_all[rows].astype(F)
    x = k1 * ((F(1.0) - b) + b * (dl / avg))
    for t in terms:
        d = F(df[t])
        if d == 0:
            continue
        idf = F(max(libm.logf(ctypes.c_float(float((F(total_docs) - d + F(0.5)) / (d + F(0.5))))), 0.0))
        tf = np.zeros(len(rows), F)
        for i, r in enumerate(rows):
            lo, hi = int(doc_off[r]), int(doc_off[r + 1])
            j = lo + np.searchsorted(doc_ids[lo:hi], t)
            if j < hi and doc_ids[j] == t:
                tf[i] = doc_tfs[j]
        m = tf > 0
        sc = (idf * (tf * (k1 + F(1.0)))) / (tf + x)
        acc = np.where(m, acc + sc.astype(F), acc).astype(F)
        hit |= m
    return acc, hit


ok = True
limit = 5 * max(3 * TOP_K, TOP_K + 10)
for qi in range(3):
    rows, sc = score_ids(term_sets[qi], limit)
    want, hit = host_scores(term_sets[qi], rows)
    ok &= bool(hit.all()) and want.tobytes() == sc.tobytes()
    keys = list(zip((-sc).tolist(), rows.tolist()))
    ok &= keys == sorted(keys) and len(set(rows.tolist())) == len(rows)
    sample = rng.choice(n_docs, 20000, replace=False)
    sample = sample[~np.isin(sample, rows)]
    s2, h2 = host_scores(term_sets[qi], sample)
    if len(rows) == limit:
        ok &= bool((s2[h2] <= sc[-1]).all())

lat_text, lat_emb, lat_score = [], [], []
for i in range(40):
    t0 = time.perf_counter(); cl.search_text_mmr(qh[i % 16], TOP_K, LAM, w, ix.handle, term_sets[i % 16]); lat_text.append(time.perf_counter() - t0)
for i in range(40):
    t0 = time.perf_counter(); cl.search_mmr(qh[i % 16], TOP_K, LAM, w); lat_emb.append(time.perf_counter() - t0)
for i in range(40):
    t0 = time.perf_counter(); score_ids(term_sets[i % 16], limit); lat_score.append(time.perf_counter() - t0)
# the text query must equal the pairs form fed with LexicalIndex::score's output
rows, sc = score_ids(term_sets[1], limit)
a = cl.search_text_mmr(qh[1], TOP_K, LAM, w, ix.handle, term_sets[1])
b = cl.search_mmr(qh[1], TOP_K, LAM, w, rows, sc)
ok &= all(x.tobytes() == y.tobytes() for x, y in zip(a, b)) and bool((a[3] != 0).any())
print(json.dumps({"workload": f"text query (8 terms, Zipf vocabulary of {V}) over {n_docs} chunks of {T} tokens (a block of {CH} documents repeated) x {dim}-d, top_k={TOP_K} diversity={LAM}, "
                              f"{len(devices)} shard(s) on devices {devices}", "api": "rlr_cluster_search_text_mmr (C ABI, host buffers, one process)",
                  "index": {"total_docs": total_docs, "total_length": total_len, "terms": n_terms, "postings": int(len(doc_ids)),
                            "build_s": round(t_build, 1), "first_query_incl_device_csr_upload_s": round(t_sync, 2)},
                  "p50_ms": {"text_query": 1e3 * statistics.median(lat_text), "embedding_only_query": 1e3 * statistics.median(lat_emb),
                             "lexical_score_only": 1e3 * statistics.median(lat_score)},
                  "lexical_limit": limit, "parity_ok": bool(ok),
                  "parity_what": "3 queries: every returned (row, score) re-scored on the host in f32 (bit-equal), rank order, no sampled "
                                 "document outranks the last one; the text query == the pairs form fed with that list (bit-equal)"}))
ix.close(); cl.close()
sys.exit(0 if ok else 3)
