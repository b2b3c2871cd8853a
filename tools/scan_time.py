"""Dev tool: time the scan kernel alone (rlr_time_scan) and the fused search for a few sizes."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import rust_local_rag_b200  # noqa
from rust_local_rag_b200 import binding as B, engine

lib = B.load()
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1_000_000]
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
f16 = len(sys.argv) > 3 and sys.argv[3] == "f16"
import torch
for n in sizes:
    s = engine.DeviceStore.synthetic(n, dim, kind=1, flags=B.RLR_STORE_F16_ONLY if f16 else 0)
    ctx = C.c_void_p()
    B.check(lib.rlr_ctx_create(s.handle, C.byref(ctx)))
    q = torch.zeros(4096 + 64, device="cuda")
    q[:dim] = torch.nn.functional.normalize(torch.randn(dim, device="cuda"), dim=0)
    torch.cuda.synchronize()
    for m in (300, 900):
        ms = C.c_float()
        B.check(lib.rlr_time_scan(ctx, C.c_void_p(q.data_ptr()), m, 20, None, C.byref(ms)))
        gb = n * dim * (2 if f16 else 4) / 1e9
        print(f"n={n} dim={dim} {'f16' if f16 else 'f32'} m={m}: scan {ms.value:.4f} ms  {gb / ms.value * 1e3:.1f} GB/s")
    w = engine.ResolvedWeights(np.float32(0.7), np.float32(0.3), np.float32(0.7), np.float32(0.3))
    qh = q[:dim].cpu().numpy()
    for _ in range(3):
        s.search_mmr(qh, 100, 0.7, w)
    t0 = time.perf_counter()
    for _ in range(20):
        s.search_mmr(qh, 100, 0.7, w, flags=B.RLR_WANT_TIMINGS)
    dt = (time.perf_counter() - t0) / 20
    t = s.last_timings()
    print(f"  e2e search_mmr k=100: {dt*1e3:.3f} ms/query  (scan {t.scan_ms:.3f} merge {t.merge_ms:.3f} mmr {t.mmr_ms:.3f} total {t.total_ms:.3f} launches {t.launches})")
    lib.rlr_ctx_destroy(ctx)
    s.close()
