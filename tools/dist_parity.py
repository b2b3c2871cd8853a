"""Multi-GPU parity check (run under torchrun, NCCL): row-sharded search over G GPUs must be
bit-identical to the single-GPU search over the whole corpus and to the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_local_rag_b200  # noqa: E402,F401
from rust_local_rag_b200 import binding as B, engine, dist as rdist  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    mode = os.environ.get("RLR_DIST_MODE", "fused")      # fused | peers | reduce
    # the fused mode is tested with an uneven (tail-balanced style) split
    plan = rdist.ShardPlan(n, world, rank, head_rows=(n // world) // 3 if mode == "fused" else None)
    kw = dict(kind=B.RLR_SYNTH_CLUSTERED, seed=11, centroid_seed=12, n_clusters=64, sigma=0.65)
    shard = engine.DeviceStore.synthetic(plan.n_local, dim, device=lr, row_base=plan.row0, **kw)
    backend = rdist.CudaBackend(shard, dev)
    use_peers = mode in ("fused", "peers")
    if use_peers:
        backend.open_peers(dist.group.WORLD, plan)
    if mode == "fused":
        backend.open_mailbox(dist.group.WORLD, m_cap=320, ring=2)
    pitch = shard.info().pitch
    qs = engine.DeviceStore.synthetic(8, dim, device=lr, **{**kw, "seed": 13})
    q_host = qs.read_rows(np.arange(8))
    full = engine.DeviceStore.synthetic(n, dim, device=lr, **kw) if rank == 0 else None
    w = engine.ResolvedWeights(np.float32(0.7), np.float32(0.3), np.float32(0.7), np.float32(0.3))
    ok = True
    if rank == 0:
        from oracle import orc
        rows_host = orc.synth_rows(n, dim, kind=1, seed=11, centroid_seed=12, n_clusters=64, sigma=0.65)
    for qi in range(8):
        for (k, lam) in ((100, 0.7), (5, 0.3), (5, 0.0), (0, 0.5)):
            p_cap = max(rdist.pool_size(k, rdist.clamp_lambda(lam)), 1)
            bufs = rdist.Buffers(world, p_cap, pitch, dev)
            q = torch.zeros(B.RLR_MAX_DIM + 64, device=dev)
            q[:dim] = torch.from_numpy(q_host[qi]).to(dev)
            res, res_n = rdist.sharded_search(backend, dist.group.WORLD, bufs, q, k, lam, float(w.embedding), float(w.lexical))
            torch.cuda.synchronize()
            if rank == 0:
                got = rdist.decode_result(res, int(res_n.item()))
                one = full.search_mmr(q_host[qi], k, lam, w, flags=B.RLR_QUERY_PRENORMALIZED)
                ref = orc.search_with_diversity(rows_host, q_host[qi], k, lam, normalize_query=False, threads=8)
                for a, b, c in zip(got[:3], one[:3], ref[:3]):
                    if a.tobytes() != b.tobytes() or a.tobytes() != c.tobytes():
                        ok = False
                        print(f"MISMATCH q={qi} k={k} lam={lam}\n sharded={a[:8]}\n single ={b[:8]}\n oracle ={c[:8]}")
    if mode == "fused" and backend.mailbox_status() != 0:
        ok = False
        print(f"rank {rank}: a mailbox wait timed out")
    flag = torch.tensor([1 if ok else 0], device=dev)
    if mode == "fused":
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.broadcast(flag, 0)
    if rank == 0:
        print("DIST_PARITY_OK" if ok else "DIST_PARITY_FAIL", f"world={world} n={n} dim={dim} mode={mode}")
    dist.barrier()
    backend.close(dist.group.WORLD)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
