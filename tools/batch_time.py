"""Dev tool: time rlr_search_batch (tcgen05 path) and report TFLOP/s."""
import sys
import time
import numpy as np
sys.path.insert(0, ".")
import rust_local_rag_b200  # noqa
from rust_local_rag_b200 import binding as B, engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
m = int(sys.argv[4]) if len(sys.argv) > 4 else 100
s = engine.DeviceStore.synthetic(n, dim, kind=1, flags=B.RLR_STORE_KEEP_F16 | B.RLR_STORE_KEEP_BF16)
qs = np.random.default_rng(0).standard_normal((nq, dim)).astype(np.float32)
for flags, name in ((B.RLR_WANT_TIMINGS | B.RLR_BATCH_F16, "binary16 operands"), (B.RLR_WANT_TIMINGS | B.RLR_BATCH_F16 | B.RLR_BATCH_EXACT_RESCORE, "binary16 + exact rescore"),
                    (B.RLR_WANT_TIMINGS | B.RLR_BATCH_BF16, "bfloat16 operands"), (B.RLR_WANT_TIMINGS | B.RLR_BATCH_TF32, "tf32 over the f32 store")):
    for rep in range(3):
        t0 = time.perf_counter()
        s.search_batch(qs, m, flags=flags)
        dt = time.perf_counter() - t0
        t = s.last_timings()
    flop = 2.0 * nq * n * dim
    print(f"n={n} dim={dim} nq={nq} m={m} [{name}]: wall {dt*1e3:.2f} ms, device {t.total_ms:.2f} ms (contraction+prunes {t.scan_ms:.2f}, rescore {t.merge_ms:.2f}), "
          f"{flop / (t.scan_ms * 1e-3) / 1e12:.1f} TFLOP/s, {nq / dt:.0f} queries/s, launches {t.launches}")
s.close()
