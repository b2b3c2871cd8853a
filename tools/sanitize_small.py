"""Small end-to-end pass over every kernel, for compute-sanitizer (memcheck / racecheck)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import rust_local_rag_b200  # noqa
from rust_local_rag_b200 import binding as B, engine
from oracle import orc

F32 = np.float32
w = engine.ResolvedWeights(F32(0.7), F32(0.3), F32(0.7), F32(0.3))
rng = np.random.default_rng(0)
n, dim = 6000, 64
rows = orc.normalize_rows(rng.standard_normal((n, dim)).astype(F32))
q = rng.standard_normal(dim).astype(F32)
s = engine.DeviceStore.from_rows(rows, flags=B.RLR_STORE_KEEP_F16)
a = s.search_mmr(q, 20, 0.5, w)
b = orc.search_with_diversity(rows, q, 20, 0.5, full_sort=True)
assert a[0].tobytes() == b[0].tobytes()
a = s.search_topm(q, 900, w, np.array([5, 77, 4000], np.uint32), np.array([1.0, 2.0, 0.5], F32))
b = orc.search(rows, q, 900, lex_rows=np.array([5, 77, 4000], np.uint32), lex_scores=np.array([1.0, 2.0, 0.5], F32), full_sort=True)
assert a[0].tobytes() == b[0].tobytes()
a = s.search_mmr(q, 20, 0.5, w, flags=B.RLR_SEARCH_F16)
b = orc.search_with_diversity(rows.astype(np.float16).astype(F32), q, 20, 0.5, full_sort=True)
assert a[0].tobytes() == b[0].tobytes()
r, sc, nn = s.search_batch(rng.standard_normal((8, dim)).astype(F32), 10, flags=B.RLR_BATCH_EXACT_RESCORE)
assert (nn == 10).all()
mf, mt = s.remove_rows([1, 2, 3, n - 1])
s.append(rows[:5])
s.search_mmr(q, 5, 0.3, w)
s.close()
s2 = engine.DeviceStore.synthetic(3000, 96, kind=1, n_clusters=8)
s2.search_mmr(rng.standard_normal(96).astype(F32), 100, 0.7, w)
s2.close()
print("SANITIZE_SMALL_OK")
