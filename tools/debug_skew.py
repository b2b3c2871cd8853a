import sys
import numpy as np
sys.path.insert(0, ".")
import rust_local_rag_b200  # noqa
from rust_local_rag_b200 import engine
from oracle import orc
F32 = np.float32
W = engine.ResolvedWeights(F32(1.0), F32(0.0), F32(0.7), F32(0.3))
cases = [(70000, 900), (70000, 1024), (70000, 300), (19000, 1000)]
sel = int(sys.argv[1]) if len(sys.argv) > 1 else -1
for (n, m) in (cases if sel < 0 else [cases[sel]]):
    dim = 64
    rng = np.random.default_rng(n + m)
    q = orc.normalize(rng.standard_normal(dim).astype(F32))
    noise = rng.standard_normal((n, dim)).astype(F32)
    scale = (np.arange(n, dtype=F32) / F32(n))[:, None]
    rows = orc.normalize_rows(q[None, :] + scale * noise)
    for name, R in (("fwd", rows), ("rev", rows[::-1].copy())):
        s = engine.DeviceStore.from_rows(R)
        for rep in range(3):
            got = s.search_topm(q, m, W)
            ref = orc.search(R, q, m, w_embed=1.0, w_lex=0.0, full_sort=False, threads=4)
            ok = got[0].tobytes() == ref[0].tobytes()
            if not ok:
                g, r = got[0], ref[0]
                bad = np.nonzero(g[:min(len(g), len(r))] != r[:min(len(g), len(r))])[0]
                print(n, m, name, rep, "MISMATCH len", len(g), len(r), "first bad", bad[:5], "set equal", set(g.tolist()) == set(r.tolist()),
                      "missing", sorted(set(r.tolist()) - set(g.tolist()))[:10], "extra", sorted(set(g.tolist()) - set(r.tolist()))[:10], flush=True)
            else:
                print(n, m, name, rep, "ok", flush=True)
        s.close()
