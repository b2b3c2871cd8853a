#!/usr/bin/env python
"""bench.py -- queries/s of `search_with_diversity(top_k=100)` over 10M x 768 f32 chunks.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A step is one query (one pass of the hot path over the whole corpus).  The corpus is the
BASELINE.json metric's: 10,000,000 x 768 synthetic unit vectors (clustered, seeded), rows
sharded contiguously over the N ranks (strong scaling: total work fixed).  `value` times the
path with queries already resident in HBM (at N=1 with two queries in flight, each on its own
workspace and stream; at N>1 one in flight, rank 0's merge/MMR tail hidden by giving rank 0 fewer
rows); `e2e` times the public host-buffer API one query at a time (C-ABI `rlr_search_mmr` at N=1,
the sharded searcher at N>1) with the H2D copy of the query and the D2H read of the result inside
the timed region.  At N>1 the per-GPU lists are exchanged by the scan kernels themselves (peer
stores into rank 0's HBM; RLR_DIST_MODE=peers|reduce selects the NCCL paths).  Prints ONE JSON
line on rank 0's stdout; everything else goes to stderr.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED_STORE, SEED_QUERY, SEED_CENTROID = 0x5EED0001, 0x5EED0002, 0x5EED00C0
N_CLUSTERS, SIGMA = 4096, 0.65
N_QUERIES = 128
METRIC = "search_with_diversity queries/sec (top_k=100 MMR, 10Mx768 f32 chunks)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--top-k", type=int, default=100)
    ap.add_argument("--diversity", type=float, default=0.7)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(a, world):
    return {
        "workload": f"single-query top_k={a.top_k} diversity={a.diversity} MMR over {a.rows}x{a.dim} f32 chunks "
                    f"(BASELINE configs[{1 if a.rows == 1_000_000 else 2}]), rows sharded contiguously over {world} GPU(s)",
        "rows": a.rows, "dim": a.dim, "top_k": a.top_k, "diversity": a.diversity,
        "pool": max(3 * a.top_k, a.top_k + 10), "weights": [0.7, 0.3],
        "distribution": f"clustered: normalize(centroid[row % {N_CLUSTERS}] + {SIGMA}*U[-1,1)), splitmix64 counter hash, "
                        f"seeds store={SEED_STORE:#x} query={SEED_QUERY:#x} centroid={SEED_CENTROID:#x}",
        "queries": N_QUERIES,
        "l2": "inputs larger than L2 (store shard >= 3.8 GB vs 126 MB L2)",
        "parallelism": f"rows/{world}",
        "queries_in_flight": max(1, int(os.environ.get("RLR_BENCH_LANES", "2" if world == 1 else "1"))),
        "exchange": {"fused": "fused: each GPU's scan kernel stores its top-300 list into rank 0's HBM mailbox (NVLink peer "
                              "stores + release flag, no collective call); rank 0 merges in a waiting kernel and its MMR reads "
                              "pool rows from peer HBM (CUDA IPC); rank 0 owns fewer rows so that its scan + merge/MMR tail "
                              "equals the other ranks' scan (tail-balanced sharding)",
                     "peers": "NCCL all-gather of per-GPU top-300 lists; MMR on rank 0 reads pool rows from peer HBM (CUDA IPC / NVLink)",
                     "reduce": "NCCL all-gather + int32 reduce of pool rows"}[dist_mode()] if world > 1 else "none (single GPU)",
    }


def dist_mode():
    m = os.environ.get("RLR_DIST_MODE", "fused")
    if m not in ("fused", "peers", "reduce"):
        raise SystemExit(f"RLR_DIST_MODE={m!r}: expected fused | peers | reduce")
    return m


# ----------------------------------------------------------------------------------------
# clocks: sampled with NVML during the timed regions
# ----------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t is not None:
            self._stop.set()
            self._t.join()
            self._t = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle (the reference is Rust; no toolchain here)
# ----------------------------------------------------------------------------------------
def cpu_reference(a, steps, warmup, budget_s):
    """Times oracle.search_with_diversity (bit-faithful restatement of the reference's
    single-process search, all host threads) on the bench workload.  Returns dict."""
    from oracle import orc
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1, which would make
    # omp_get_max_threads() say 1; the oracle takes its thread count explicitly (num_threads clause)
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    threads = max(orc.max_threads(), avail)
    rows_n = a.rows
    t0 = time.perf_counter()
    rows = orc.synth_rows(rows_n, a.dim, kind=1, seed=SEED_STORE, centroid_seed=SEED_CENTROID,
                          n_clusters=N_CLUSTERS, sigma=SIGMA, threads=threads)
    gen_s = time.perf_counter() - t0
    qs = orc.synth_rows(N_QUERIES, a.dim, kind=1, seed=SEED_QUERY, centroid_seed=SEED_CENTROID,
                        n_clusters=N_CLUSTERS, sigma=SIGMA, threads=1)
    # probe one query to size the sample
    t0 = time.perf_counter()
    orc.search_with_diversity(rows, qs[0], a.top_k, a.diversity, threads=threads)
    probe = time.perf_counter() - t0
    sample_rows = rows_n
    if probe * (steps + warmup) > budget_s:
        sample_rows = max(100_000, int(rows_n * budget_s / (probe * (steps + warmup))))
        rows = rows[:sample_rows]
    for i in range(warmup):
        orc.search_with_diversity(rows, qs[i % N_QUERIES], a.top_k, a.diversity, threads=threads)
    lat = []
    for i in range(steps):
        t0 = time.perf_counter()
        orc.search_with_diversity(rows, qs[(warmup + i) % N_QUERIES], a.top_k, a.diversity, threads=threads)
        lat.append(time.perf_counter() - t0)
    total = sum(lat)
    sample = (f"{steps} queries, each the full path (scan + top-{max(3 * a.top_k, a.top_k + 10)} + literal O(k^2 P D) MMR) "
              f"over {sample_rows} of {rows_n} rows x {a.dim}, {threads} OpenMP threads; host rows generated in {gen_s:.1f}s")
    if sample_rows != rows_n:
        sample += f"; ROWS SUBSAMPLED x{rows_n / sample_rows:.1f} to bound run time: value is for the subsample"
    return {"value": steps / total, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample,
            "ms_per_step": 1e3 * total / steps, "p50_ms": 1e3 * statistics.median(lat), "sample_rows": sample_rows}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference(a, a.steps, a.warmup, budget_s=200.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "queries/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, a.gpus),
        "p50_latency_ms": r["p50_ms"],
        "cpu_baseline": {"value": r["value"], "unit": "queries/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference arm = CPU oracle (C restatement of rag_engine.rs search/MMR arithmetic, OpenMP over rows); "
                "the Rust reference cannot be built here (no cargo/rustc) and additionally clones every chunk and "
                "sorts N fat tuples per query, so this is a lower bound on its time",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# own arm
# ----------------------------------------------------------------------------------------
def run_b200(a, guard=None):
    import numpy as np
    import torch
    import torch.distributed as dist

    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import binding as B, engine
    from rust_local_rag_b200 import dist as rdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {a.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    lib = B.load()

    mode = dist_mode() if world > 1 else "single"
    p_cap = max(3 * a.top_k, a.top_k + 10, 1)
    w_e, w_l = float(np.float32(0.7)), float(np.float32(0.3))

    def build_shard(plan):
        nonlocal mode
        st = engine.DeviceStore.synthetic(plan.n_local, a.dim, kind=B.RLR_SYNTH_CLUSTERED, seed=SEED_STORE,
                                          centroid_seed=SEED_CENTROID, n_clusters=N_CLUSTERS, sigma=SIGMA,
                                          device=local_rank, row_base=plan.row0)
        be = rdist.CudaBackend(st, dev)
        try:
            if mode in ("fused", "peers"):
                be.open_peers(group, plan)       # rank 0 maps the peer shards (CUDA IPC over NVLink)
            if mode == "fused":
                be.open_mailbox(group, m_cap=p_cap, ring=4)
        except B.RlrError as e:                  # raised on EVERY rank (dist._agree): no peer access on this box
            if rank == 0:
                print(f"[bench] peer-memory setup failed ({e}); falling back to the NCCL all-gather + reduce path",
                      file=sys.stderr, flush=True)
            be.close()
            be = rdist.CudaBackend(st, dev)
            mode = "reduce"
            os.environ["RLR_DIST_MODE"] = "reduce"   # workload_config reports what actually ran
        return st, be, rdist.Buffers(world, p_cap, st.info().pitch, dev)

    plan = rdist.ShardPlan(a.rows, world, rank)
    store, backend, bufs = build_shard(plan)
    info = store.info()
    pitch = info.pitch
    # queries: same generator, different noise seed (clustered around the same centroids)
    qstore = engine.DeviceStore.synthetic(N_QUERIES, a.dim, kind=B.RLR_SYNTH_CLUSTERED, seed=SEED_QUERY,
                                          centroid_seed=SEED_CENTROID, n_clusters=N_CLUSTERS, sigma=SIGMA,
                                          device=local_rank)
    q_host = qstore.read_rows(np.arange(N_QUERIES))
    qstore.close()
    qcap = B.RLR_MAX_DIM + 64
    q_pinned = torch.zeros((N_QUERIES, qcap), dtype=torch.float32).pin_memory()
    q_pinned[:, :a.dim] = torch.from_numpy(q_host)
    q_dev = q_pinned.to(dev)

    balance = None
    if mode == "fused" and os.environ.get("RLR_DIST_BALANCE", "1") == "1":
        # tail-balanced sharding: measure rank 0's merge + MMR tail (flags of a finished query are
        # already set, so re-running the two steps times the tail alone) and the local scan rate,
        # then give rank 0 that many fewer rows and rebuild the shards.
        res0 = rdist.sharded_search(backend, group, bufs, q_dev[0], a.top_k, a.diversity, w_e, w_l)
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)
        cal = torch.zeros(2, dtype=torch.float64, device=dev)
        if rank == 0:
            lam = rdist.clamp_lambda(a.diversity)
            m = rdist.pool_size(a.top_k, lam)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            for it in range(reps + 3):
                if it == 3:
                    e0.record()
                backend.mailbox_merge(backend._seq, m, bufs.pool[:m], bufs.pool_n)
                if lam != 0.0:
                    backend.mmr_peers(bufs.pool[:m], bufs.pool_n, m, a.top_k, lam, bufs.sel_pos, bufs.sel_n, bufs.result)
            e1.record()
            torch.cuda.synchronize(dev)
            tail_ms = e0.elapsed_time(e1) / reps
            iso = C.c_float(0)
            B.check(lib.rlr_time_scan(backend.ctx, C.c_void_p(q_dev[0].data_ptr()), p_cap, 10,
                                      C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), C.byref(iso)))
            cal[0], cal[1] = tail_ms, plan.n_local / iso.value
        dist.broadcast(cal, 0, group=group)
        tail_ms, rows_per_ms = float(cal[0].item()), float(cal[1].item())
        head = rdist.ShardPlan.balanced_head_rows(a.rows, world, tail_ms * rows_per_ms)
        balance = {"tail_ms": tail_ms, "scan_rows_per_ms": rows_per_ms, "rank0_rows": head,
                   "other_rank_rows": (a.rows - head) // (world - 1)}
        backend.close(group)
        dist.barrier(group=group)
        store.close()
        del bufs
        dist.barrier(group=group)
        plan = rdist.ShardPlan(a.rows, world, rank, head_rows=head)
        store, backend, bufs = build_shard(plan)
    result_host = torch.zeros((p_cap, 2), dtype=torch.int64).pin_memory()
    n_host = torch.zeros(1, dtype=torch.int32).pin_memory()
    q_stage = torch.zeros(qcap, dtype=torch.float32, device=dev)

    # `value` keeps LANES queries in flight per rank (each lane: its own workspace, result buffers and CUDA
    # stream), as a server with concurrent searches does: while the last CTA of one scan merges the per-CTA lists
    # and the MMR kernels run, the next query's scan already streams rows on the other SMs.  Latency (e2e, p50)
    # is measured one query at a time further down.
    # Sharded runs keep one query in flight: there rank 0's tail is already hidden by giving rank 0 fewer rows
    # (measured at N=2: 471.8 q/s balanced + 1 lane, 469.8 balanced + 2 lanes, 481.6 even shards + 2 lanes but with
    # e2e/p50 1.5 % worse; the balanced split serves both numbers).
    lanes = max(1, int(os.environ.get("RLR_BENCH_LANES", "2" if world == 1 else "1")))
    for _ in range(lanes - 1):
        backend.add_lane()
    lane_bufs = [bufs] + [rdist.Buffers(world, p_cap, pitch, dev) for _ in range(lanes - 1)]
    lane_streams = [torch.cuda.Stream(dev) for _ in range(lanes)]

    def step_device(i, b=None):
        b = bufs if b is None else b
        q = q_dev[i % N_QUERIES]
        if world == 1:
            backend.search_mmr(q, a.top_k, a.diversity, w_e, w_l, b.result, b.sel_n)
            return b.result, b.sel_n
        return rdist.sharded_search(backend, group, b, q, a.top_k, a.diversity, w_e, w_l)

    def run_steps(first, count):
        cur = torch.cuda.current_stream(dev)
        for s_ in lane_streams:
            s_.wait_stream(cur)
        for i in range(count):
            lane = i % lanes
            backend.use_lane(lane)
            with torch.cuda.stream(lane_streams[lane]):
                step_device(first + i, lane_bufs[lane])
        for s_ in lane_streams:
            cur.wait_stream(s_)
        backend.use_lane(0)

    def sync_all():
        if world > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t.item())

    clocks = ClockSampler(local_rank)

    # ---- value: queries resident in HBM, device-timed ----
    run_steps(0, a.warmup)
    sync_all()
    l0 = backend.launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.start()
    ev0.record()
    run_steps(a.warmup, a.steps)
    ev1.record()
    sync_all()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = backend.launches() - l0
    value = a.steps / (dev_ms * 1e-3)

    # ---- e2e: host buffers through the public API, H2D + D2H inside the timed region ----
    wts = engine.ResolvedWeights(np.float32(0.7), np.float32(0.3), np.float32(0.7), np.float32(0.3))
    scan_ms_samples, stage_samples = [], []

    def step_e2e(i, timed):
        qi = i % N_QUERIES
        if world == 1:
            out = store.search_mmr(q_host[qi], a.top_k, a.diversity, wts, flags=B.RLR_WANT_TIMINGS if timed else 0)
            if timed:
                t = store.last_timings()
                scan_ms_samples.append(t.scan_ms)
                stage_samples.append((t.scan_ms, t.merge_ms, t.mmr_ms, t.total_ms))
            return out
        q_stage.copy_(q_pinned[qi], non_blocking=True)
        res, n = rdist.sharded_search(backend, group, bufs, q_stage, a.top_k, a.diversity, w_e, w_l)
        if rank == 0:
            result_host.copy_(res[:p_cap], non_blocking=True)
            n_host.copy_(n, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return None

    for i in range(a.warmup):
        step_e2e(i, False)
    sync_all()
    lat = []
    t_start = time.perf_counter()
    for i in range(a.steps):
        t0 = time.perf_counter()
        step_e2e(a.warmup + i, True)
        lat.append(time.perf_counter() - t0)
    sync_all()
    e2e_s = max_over_ranks(time.perf_counter() - t_start)
    clocks.stop()
    e2e_value = a.steps / e2e_s
    h2d = (pitch + 64) * 4 if world == 1 else qcap * 4
    d2h = max(a.top_k, 1) * 16 + 4 if world == 1 else p_cap * 16 + 4

    # ---- roofline: the scan kernel (dominant) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured copy)") if "hbm_gbs" in peaks \
        else (6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)")
    iso_ms = C.c_float(0)
    B.check(lib.rlr_time_scan(backend.ctx, C.c_void_p(q_dev[0].data_ptr()), p_cap, 20,
                              C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), C.byref(iso_ms)))
    algo_bytes = plan.n_local * a.dim * 4
    if scan_ms_samples:
        scan_ms = statistics.mean(scan_ms_samples)
        how = "CUDA events around the scan kernel inside every timed e2e call (mean)"
    else:
        scan_ms = iso_ms.value
        how = "20 back-to-back launches after the timed region, CUDA events on the launch stream"
    achieved = algo_bytes / (scan_ms * 1e-3) / 1e9
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel from the committed
    # `ncu --set full` capture; used only when that capture is of exactly this launch shape (else null)
    traffic, traffic_src = None, None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["scan_topm_kernel"]
        if t["rows"] == plan.n_local and t["dim"] == a.dim and t["elem_bytes"] == 4:
            traffic, traffic_src = t["dram_bytes_read"] + t["dram_bytes_write"], t["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": "scan_topm_kernel", "bytes_per_launch": algo_bytes,
                "ms_per_launch": scan_ms, "isolated_ms_per_launch": iso_ms.value,
                "isolated_GBps": algo_bytes / (iso_ms.value * 1e-3) / 1e9, "peak_source": peak_src, "how": how,
                "frac_of_nominal_8TBps": achieved / 8000.0}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = cpu_reference(a, steps=6, warmup=1, budget_s=60.0)
        cpu = {"value": r["value"], "unit": "queries/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "p50_ms": r["p50_ms"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a, world),
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "p50_latency_ms": 1e3 * statistics.median(lat), "p99_latency_ms": 1e3 * sorted(lat)[int(0.99 * (len(lat) - 1))],
                    "api": "rlr_search_mmr (C ABI, host buffers)" if world == 1 else
                           "rust_local_rag_b200.dist.sharded_search (pinned query H2D, result D2H on rank 0)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks.summary(),
        }
        if balance is not None:
            line["config"]["tail_balance"] = balance
        if stage_samples:
            m = [statistics.mean(x[j] for x in stage_samples) for j in range(4)]
            line["stage_ms"] = {"scan": m[0], "merge": m[1], "mmr": m[2], "device_total": m[3]}
        if guard is not None:
            guard.restore()
        print(json.dumps(line), flush=True)
        if guard is not None:
            guard.__enter__()
    timeouts = backend.mailbox_status() if mode == "fused" else 0
    backend.close(group)
    if world > 1:
        dist.barrier(group=group)
    store.close()
    if world > 1:
        dist.barrier(group=group)
        dist.destroy_process_group()
    if timeouts:
        raise SystemExit(f"rank {rank}: a mailbox wait timed out (status {timeouts}); the numbers above are invalid")


class StdoutToStderr:
    """Libraries (NCCL's version banner, for one) print to fd 1; the contract is ONE JSON line on
    stdout.  Everything written to fd 1 while this is active goes to stderr instead."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def restore(self):
        if self.saved is not None:
            sys.stdout.flush()
            os.dup2(self.saved, 1)
            os.close(self.saved)
            self.saved = None

    def __exit__(self, *exc):
        self.restore()
        return False


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        with StdoutToStderr() as guard:
            run_b200(args, guard)
