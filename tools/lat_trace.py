"""Dev tool: latency-path timing at BASELINE config 1 (10k x 768, top_k=5, diversity=0.3).
RLR_DEBUG_LAT_TRACE=1 prints in-kernel phase timestamps for the first requests."""
import sys, time, statistics
import numpy as np
sys.path.insert(0, ".")
import rust_local_rag_b200  # noqa
from rust_local_rag_b200 import binding as B, engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 5
lam = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
w = engine.ResolvedWeights(np.float32(0.7), np.float32(0.3), np.float32(0.7), np.float32(0.3))
for flags, name in ((0, "latency path"), (B.RLR_STORE_NO_LATENCY_PATH, "regular path")):
    s = engine.DeviceStore.synthetic(n, 768, kind=1, n_clusters=256, flags=flags)
    qs = np.random.default_rng(0).standard_normal((64, 768)).astype(np.float32)
    for i in range(20):
        s.search_mmr(qs[i], k, lam, w)
    lat = []
    for i in range(2000):
        t0 = time.perf_counter()
        s.search_mmr(qs[i % 64], k, lam, w)
        lat.append(time.perf_counter() - t0)
    s.search_mmr(qs[0], k, lam, w, flags=B.RLR_WANT_TIMINGS)
    t = s.last_timings()
    print(f"{name}: n={n} k={k} lam={lam}: p50 {1e6*statistics.median(lat):.1f} us, p99 {1e6*sorted(lat)[1980]:.1f} us, min {1e6*min(lat):.1f} us; "
          f"device total {1e3*t.total_ms:.1f} us (scan {1e3*t.scan_ms:.1f}, mmr {1e3*t.mmr_ms:.1f}), launches {t.launches}")
    s.close()
