#!/usr/bin/env python
"""bench.py -- queries/s of `search_with_diversity(top_k=100)` over 10M x 768 f32 chunks.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A step is one query (one pass of the hot path over the whole corpus).  The corpus is the
BASELINE.json metric's: 10,000,000 x 768 synthetic unit vectors (clustered, seeded), rows
sharded contiguously over the N GPUs (strong scaling: total work fixed).

  value   the path with queries already resident in HBM, CUDA events, max over ranks.  N=1: two
          queries in flight (two workspaces + streams).  N>1: one process per GPU (torchrun), the
          per-GPU lists exchanged by the scan kernels themselves (peer stores into rank 0's HBM).
  e2e     the public host-buffer C-ABI call, one query at a time, H2D of the query and D2H of the
          result inside the timed region: `rlr_search_mmr` at N=1, `rlr_cluster_search_mmr` at
          N>1 -- ONE process (rank 0) driving all N GPUs, as the reference is one process; the
          other ranks free their shards and wait on a CPU barrier.
  parity  the answers are checked IN THIS RUN and the run fails (exit 3) on a mismatch: every
          returned row of the first queries is regenerated on the CPU and re-scored with the
          oracle's sequential f32 dot (bit-equal), at N=1 the complete results are compared with the
          oracle's over all 10M rows, and `parity.digest` (sha256 of rows + score bits) must be the
          same string at every N.
Prints ONE JSON line on rank 0's stdout; everything else goes to stderr.  `extra_configs` holds
BASELINE configs 1, 2, 4 (and 5 at N=8), each with its own clock record.
"""
import argparse
import ctypes as C
import hashlib
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
N_PARITY = 8            # queries whose answers are verified in every run
METRIC = "search_with_diversity queries/sec (top_k=100 MMR, 10Mx768 f32 chunks)"
SYNTH = dict(kind=1, seed=SEED_STORE, centroid_seed=SEED_CENTROID, n_clusters=N_CLUSTERS, sigma=SIGMA)


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


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
    ap.add_argument("--extras", default="auto", choices=["auto", "none", "all"],
                    help="extra_configs (BASELINE configs 1, 2, 4, 5): auto = at N=1 and N=8")
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
        "queries_in_flight": max(1, int(os.environ.get("RLR_BENCH_LANES", "2"))),
        "exchange": {"fused": "fused: each GPU's scan kernel stores its top-300 list into rank 0's HBM mailbox (NVLink peer "
                              "stores + release flag, no collective call); rank 0 merges in a waiting kernel and its MMR reads "
                              "pool rows from peer HBM; rank 0 owns fewer rows so that its scan + merge/MMR tail "
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
    def __init__(self, index, interval=0.02):
        self.interval = interval
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
            self._stop.wait(self.interval)

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


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ----------------------------------------------------------------------------------------
# parity: verify answers in the run that reports the numbers
# ----------------------------------------------------------------------------------------
def result_digest(results):
    h = hashlib.sha256()
    import numpy as np
    for rows, score, _emb in results:
        h.update(np.ascontiguousarray(rows, np.uint32).tobytes())
        h.update(np.ascontiguousarray(score, np.float32).tobytes())
    return h.hexdigest()


def verify_results(results, queries_used, a, label):
    """Size-independent check of search results against the oracle's arithmetic: every returned row is
    regenerated on the CPU (counter hash + reference normalize) and re-scored with the oracle's sequential
    f32 dot; embedding_score must be bit-equal, score must be fl(fl(0.7*e) + fl(0.3*0)) (:531-532), rows must be
    distinct and every result must hold top_k entries.  Returns (ok, failures)."""
    import numpy as np
    from oracle import orc
    fails = []
    w_e, w_l = np.float32(0.7), np.float32(0.3)
    for qi, ((rows, score, emb), q) in enumerate(zip(results, queries_used)):
        want_n = min(max(a.top_k, 1), a.rows)
        if len(rows) != want_n:
            fails.append(f"{label} q{qi}: {len(rows)} results, expected {want_n}")
            continue
        if len(set(rows.tolist())) != len(rows) or (len(rows) and int(rows.max()) >= a.rows):
            fails.append(f"{label} q{qi}: duplicate or out-of-range rows")
        for r, s, e in zip(rows.tolist(), score, emb):
            row = orc.synth_rows(1, a.dim, row0=int(r), threads=1, **SYNTH)[0]
            e_ref = np.float32(orc.dot(q, row))
            s_ref = np.float32(np.float32(w_e * e_ref) + np.float32(w_l * np.float32(0.0)))
            if e_ref.tobytes() != np.float32(e).tobytes() or s_ref.tobytes() != np.float32(s).tobytes():
                fails.append(f"{label} q{qi} row {r}: emb {float(e)!r} vs oracle {float(e_ref)!r}, score {float(s)!r} vs {float(s_ref)!r}")
                break
    return (not fails), fails[:5]


# ----------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle (the reference is Rust; no toolchain here)
# ----------------------------------------------------------------------------------------
def cpu_reference(a, steps, warmup, budget_s, keep_results=False):
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
    rows = orc.synth_rows(rows_n, a.dim, threads=threads, **SYNTH)
    gen_s = time.perf_counter() - t0
    qs = orc.synth_rows(N_QUERIES, a.dim, threads=1, **{**SYNTH, "seed": SEED_QUERY})
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
    lat, results = [], {}
    for i in range(steps):
        qi = (warmup + i) % N_QUERIES
        t0 = time.perf_counter()
        r = orc.search_with_diversity(rows, qs[qi], a.top_k, a.diversity, threads=threads)
        lat.append(time.perf_counter() - t0)
        if keep_results:
            results[qi] = (r[0].copy(), r[1].copy(), r[2].copy())
    total = sum(lat)
    sample = (f"{steps} queries, each the full path (scan + top-{max(3 * a.top_k, a.top_k + 10)} + literal O(k^2 P D) MMR) "
              f"over {sample_rows} of {rows_n} rows x {a.dim}, {threads} OpenMP threads; host rows generated in {gen_s:.1f}s")
    if sample_rows != rows_n:
        sample += f"; ROWS SUBSAMPLED x{rows_n / sample_rows:.1f} to bound run time: value is for the subsample"
    return {"value": steps / total, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample,
            "ms_per_step": 1e3 * total / steps, "p50_ms": 1e3 * statistics.median(lat), "sample_rows": sample_rows,
            "results": results}


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
    group = cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
        cpu_group = dist.new_group(backend="gloo")     # host-side waits that must not occupy the GPUs
    lib = B.load()

    mode = dist_mode() if world > 1 else "single"
    p_cap = max(3 * a.top_k, a.top_k + 10, 1)
    w_e, w_l = float(np.float32(0.7)), float(np.float32(0.3))
    peaks = load_peaks()

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
                log(f"peer-memory setup failed ({e}); falling back to the NCCL all-gather + reduce path")
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
        rdist.sharded_search(backend, group, bufs, q_dev[0], a.top_k, a.diversity, w_e, w_l)
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
    lanes = max(1, int(os.environ.get("RLR_BENCH_LANES", "2")))
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

    # ---- parity, device-resident path: the same calls as `value`, results brought back and verified ----
    device_results = []
    for qi in range(N_PARITY):
        res, n = step_device(qi)
        torch.cuda.synchronize(dev)
        if rank == 0:
            r, s, e, _ = rdist.decode_result(res, int(n.item()))
            device_results.append((r, s, e))
    sync_all()

    # ---- throughput mode (N=1): Q queries answered by ONE pass over the rows (query groups, scan_topm.cu) ----
    throughput = None
    if world == 1 and os.environ.get("RLR_BENCH_MULTI", "1") == "1":
        try:
            throughput = run_throughput_mode(a, backend, store, q_dev, q_host, dev, p_cap, w_e, w_l, lanes)
        except B.RlrError as e:
            throughput = {"error": str(e)}

    # ---- e2e: host buffers through the public API, H2D + D2H inside the timed region ----
    wts = engine.ResolvedWeights(np.float32(0.7), np.float32(0.3), np.float32(0.7), np.float32(0.3))
    scan_ms_samples, stage_samples = [], []
    e2e_multiprocess = None
    cluster = None
    cluster_info = None

    if world > 1:
        # (a) the one-process-per-GPU path end to end (kept as a second mode): pinned query H2D on every rank,
        #     sharded search, result D2H on rank 0
        def step_mp(i):
            qi = i % N_QUERIES
            q_stage.copy_(q_pinned[qi], non_blocking=True)
            res, n = rdist.sharded_search(backend, group, bufs, q_stage, a.top_k, a.diversity, w_e, w_l)
            if rank == 0:
                result_host.copy_(res[:p_cap], non_blocking=True)
                n_host.copy_(n, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
        for i in range(a.warmup):
            step_mp(i)
        sync_all()
        lat = []
        t_start = time.perf_counter()
        for i in range(a.steps):
            t0 = time.perf_counter()
            step_mp(a.warmup + i)
            lat.append(time.perf_counter() - t0)
        sync_all()
        mp_s = max_over_ranks(time.perf_counter() - t_start)
        e2e_multiprocess = {"value": a.steps / mp_s, "unit": "queries/s", "p50_latency_ms": 1e3 * statistics.median(lat),
                            "api": "rust_local_rag_b200.dist.sharded_search under torchrun (pinned query H2D on every rank, "
                                   "result D2H on rank 0)"}
        # (b) the drop-in path: ONE process (this rank 0) drives all GPUs through the C ABI.  The other
        #     ranks free their shards and wait on a CPU (gloo) barrier so that their GPUs are idle.
        timeouts = backend.mailbox_status() if mode == "fused" else 0
        backend.close(group)
        dist.barrier(group=group)
        store.close()
        del bufs, lane_bufs
        torch.cuda.empty_cache()
        dist.barrier(group=group)
        if timeouts:
            raise SystemExit(f"rank {rank}: a mailbox wait timed out (status {timeouts}); the run is invalid")
        if rank == 0:
            shard_rows = None
            if balance is not None:
                shard_rows = [rdist.ShardPlan(a.rows, world, r, head_rows=balance["rank0_rows"]).n_local for r in range(world)]
            cluster = engine.ClusterStore.synthetic(a.rows, a.dim, kind=B.RLR_SYNTH_CLUSTERED, seed=SEED_STORE,
                                                    centroid_seed=SEED_CENTROID, n_clusters=N_CLUSTERS, sigma=SIGMA,
                                                    devices=list(range(world)), shard_rows=shard_rows)
            ci = cluster.cluster_info()
            cluster_info = {"shard_rows": [int(ci.shard_rows[g]) for g in range(ci.n_shards)],
                            "devices": [int(ci.device[g]) for g in range(ci.n_shards)]}

    api_results = []
    e2e = None
    if rank == 0:
        target = store if world == 1 else cluster

        def step_e2e(i, timed):
            qi = i % N_QUERIES
            out = target.search_mmr(q_host[qi], a.top_k, a.diversity, wts, flags=B.RLR_WANT_TIMINGS if timed else 0)
            if timed:
                t = target.last_timings()
                if world == 1:
                    scan_ms_samples.append([t.scan_ms])
                else:
                    scan_ms_samples.append(cluster.last_scan_ms())
                stage_samples.append((t.scan_ms, t.merge_ms, t.mmr_ms, t.total_ms))
            return out

        for i in range(a.warmup):
            step_e2e(i, False)
        # (1) ONE caller, one query at a time: the latency of a call (and the stage timings for the roofline)
        lat = []
        t_start = time.perf_counter()
        for i in range(a.steps):
            t0 = time.perf_counter()
            step_e2e(a.warmup + i, True)
            lat.append(time.perf_counter() - t0)
        single_s = time.perf_counter() - t_start
        # (2) CALLERS host threads calling concurrently, as concurrent searches under the reference's read lock do
        #     (src/mcp_server.rs:89,377): every call still carries its own H2D query copy and D2H result read.  While
        #     one query's merge + MMR tail runs, the next query's scan already streams rows.
        callers = max(1, int(os.environ.get("RLR_BENCH_CALLERS", "2")))
        errs, lat_multi = [], []
        start = threading.Barrier(callers + 1)

        def caller(t):
            try:
                for i in range(3):      # untimed: every caller's lane (workspaces, streams, mailbox) is created here
                    step_e2e(t + i, False)
                start.wait()
                for i in range(t, a.steps, callers):
                    t0 = time.perf_counter()
                    step_e2e(a.warmup + i, False)
                    lat_multi.append(time.perf_counter() - t0)
            except Exception as e:      # noqa: BLE001
                errs.append(repr(e))

        th = [threading.Thread(target=caller, args=(t,)) for t in range(callers)]
        [t.start() for t in th]
        start.wait()
        t_start = time.perf_counter()
        [t.join() for t in th]
        e2e_s = time.perf_counter() - t_start
        if errs:
            raise SystemExit(f"concurrent e2e callers failed: {errs[:2]}")
        q_bytes = (((a.dim + 63) & ~63) + 128) * 4
        e2e = {"value": a.steps / e2e_s, "unit": "queries/s", "callers": callers,
               "h2d_bytes_per_step": q_bytes * world, "d2h_bytes_per_step": max(a.top_k, 1) * 16 + 16,
               "p50_latency_ms": 1e3 * statistics.median(lat), "p99_latency_ms": 1e3 * sorted(lat)[int(0.99 * (len(lat) - 1))],
               "single_caller": {"value": a.steps / single_s, "unit": "queries/s", "p50_latency_ms": 1e3 * statistics.median(lat)},
               "p50_latency_ms_with_concurrent_callers": 1e3 * statistics.median(lat_multi),
               "note": f"value: {callers} host threads calling concurrently (each call: query H2D, search, result D2H); p50/p99: ONE "
                       "caller, one query at a time (the latency of a call from idle GPUs)",
               "api": "rlr_search_mmr (C ABI, host buffers)" if world == 1 else
                      f"rlr_cluster_search_mmr (C ABI, host buffers; ONE process drives {world} GPUs, peer-memory mailbox + peer-pointer MMR)"}
        if cluster_info is not None:
            e2e["cluster"] = cluster_info
        # parity, public API path: the call a user makes (it normalises the query, :494)
        for qi in range(N_PARITY):
            r, s, e, _ = target.search_mmr(q_host[qi], a.top_k, a.diversity, wts)
            api_results.append((r.copy(), s.copy(), e.copy()))
        if world > 1 and os.environ.get("RLR_BENCH_MULTI", "1") == "1":
            # throughput mode through the cluster's C-ABI call: Q queries per pass over every GPU's shard.  The root
            # now carries Q merge + MMR tails per pass, so this mode gets its own tail-balanced plan (Q = 3).
            try:
                singles = [target.search_mmr(q_host[i], a.top_k, a.diversity, wts) for i in range(3)]
                tp_cluster, tp_plan = cluster, cluster_info["shard_rows"]
                if balance is not None:
                    head3 = rdist.ShardPlan.balanced_head_rows(a.rows, world, 3 * balance["tail_ms"] * balance["scan_rows_per_ms"])
                    tp_plan = [rdist.ShardPlan(a.rows, world, r, head_rows=head3).n_local for r in range(world)]
                    cluster.close()
                    cluster = None
                    tp_cluster = engine.ClusterStore.synthetic(a.rows, a.dim, kind=B.RLR_SYNTH_CLUSTERED, seed=SEED_STORE,
                                                               centroid_seed=SEED_CENTROID, n_clusters=N_CLUSTERS, sigma=SIGMA,
                                                               devices=list(range(world)), shard_rows=tp_plan)
                throughput = api_throughput_mode(a, tp_cluster, q_host, wts, callers, singles,
                                                 f"rlr_cluster_search_mmr_multi (C ABI, host buffers, one process, {world} GPUs)")
                throughput["shard_rows"] = tp_plan
                if tp_cluster is not cluster:
                    tp_cluster.close()
            except B.RlrError as e:
                throughput = {"error": str(e)}
    clocks.stop()

    # ---- roofline: the scan kernel (dominant) ----
    peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured copy)") if "hbm_gbs" in peaks \
        else (6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)")
    roofline = None
    if rank == 0:
        if world == 1:
            iso_ms = C.c_float(0)
            B.check(lib.rlr_time_scan(backend.ctx, C.c_void_p(q_dev[0].data_ptr()), p_cap, 20,
                                      C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), C.byref(iso_ms)))
            rows_of = [plan.n_local]
            iso = iso_ms.value
        else:
            rows_of = cluster_info["shard_rows"]
            iso = None
        per_shard_ms = [statistics.mean(x[g] for x in scan_ms_samples) for g in range(len(rows_of))]
        # the dominant launch: the largest shard's scan (every non-root GPU runs one of that size per query)
        g_dom = max(range(len(rows_of)), key=lambda g: rows_of[g])
        algo_bytes = rows_of[g_dom] * a.dim * 4
        scan_ms = per_shard_ms[g_dom]
        achieved = algo_bytes / (scan_ms * 1e-3) / 1e9
        # traffic: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel from the committed
        # `ncu --set full` capture, used only when that capture is of exactly this launch shape AND of this source tree
        traffic, traffic_src = None, None
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["scan_topm_kernel"]
            if t["rows"] == rows_of[g_dom] and t["dim"] == a.dim and t["elem_bytes"] == 4 and t.get("scan_src_sha16") == scan_source_sha16():
                traffic, traffic_src = t["dram_bytes_read"] + t["dram_bytes_write"], t["source"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "kernel": "scan_topm_kernel", "bytes_per_launch": algo_bytes,
                    "ms_per_launch": scan_ms, "peak_source": peak_src,
                    "how": "CUDA events around the scan kernel on its launch stream inside every timed e2e call (mean)",
                    "frac_of_nominal_8TBps": achieved / 8000.0,
                    "per_gpu": [{"rows": rows_of[g], "scan_ms": per_shard_ms[g],
                                 "GBps": rows_of[g] * a.dim * 4 / (per_shard_ms[g] * 1e-3) / 1e9} for g in range(len(rows_of))]}
        if iso is not None:
            roofline["isolated_ms_per_launch"] = iso
            roofline["isolated_GBps"] = algo_bytes / (iso * 1e-3) / 1e9

    # ---- cpu baseline (rank 0, N=1) + full-oracle parity ----
    cpu = None
    parity = None
    if rank == 0:
        from oracle import orc
        fails = []
        q_norm = [orc.normalize(q_host[qi]) for qi in range(N_PARITY)]
        ok_api, f = verify_results(api_results, q_norm, a, "api")
        fails += f
        ok_dev, f = verify_results(device_results, [q_host[qi] for qi in range(N_PARITY)], a, "device")
        fails += f
        oracle_compared = 0
        if world == 1 and not a.no_cpu_baseline:
            n_base = 6
            r = cpu_reference(a, steps=n_base, warmup=1, budget_s=60.0, keep_results=True)
            cpu = {"value": r["value"], "unit": "queries/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
                   "p50_ms": r["p50_ms"]}
            if r["sample_rows"] == a.rows:
                for qi, (R, S, E) in r["results"].items():
                    if qi < N_PARITY:
                        got = api_results[qi]
                    else:
                        g = store.search_mmr(q_host[qi], a.top_k, a.diversity, wts)
                        got = (g[0], g[1], g[2])
                    oracle_compared += 1
                    if not (got[0].tobytes() == R.tobytes() and got[1].tobytes() == S.tobytes() and got[2].tobytes() == E.tobytes()):
                        fails.append(f"api q{qi}: result differs from the oracle's search_with_diversity over all {a.rows} rows")
        parity = {"checked": N_PARITY, "ok": not fails, "digest": result_digest(api_results),
                  "device_digest": result_digest(device_results),
                  "rows_rescored": sum(len(x[0]) for x in api_results) + sum(len(x[0]) for x in device_results),
                  "oracle_full_results_compared": oracle_compared,
                  "what": "digest = sha256(rows, score bits) of the first 8 queries through the public C-ABI call (must be identical at "
                          "every N); device_digest = same for the device-resident path `value` times; every returned row regenerated "
                          "on the CPU and re-scored with the oracle's sequential f32 dot (bit-equal emb and blended score); at N=1 the "
                          "complete result lists are also compared with the oracle's search over all rows",
                  "failures": fails}

    # ---- extra configs (BASELINE configs 1, 2, 4, 5), each with its own clock record ----
    extras = None
    want_extras = a.extras == "all" or (a.extras == "auto" and world in (1, 8) and a.rows == 10_000_000)
    if world > 1:
        if cluster is not None:
            cluster.close()
            cluster = None
        torch.cuda.empty_cache()
        dist.barrier(group=cpu_group)          # ranks != 0 have been waiting here, GPUs idle
    else:
        backend.close(group)
        store.close()
        torch.cuda.empty_cache()
    if want_extras:
        try:
            extras = run_extras(a, rank, world, local_rank, dev, group, peaks, cpu_group)
        except Exception as e:      # noqa: BLE001 -- extras must never take the headline line down with them
            extras = {"error": repr(e)}
            log("extra configs failed:", repr(e))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a, world),
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity": parity,
            "clocks": clocks.summary(),
        }
        if e2e_multiprocess is not None:
            line["e2e_multiprocess"] = e2e_multiprocess
        if balance is not None:
            line["config"]["tail_balance"] = balance
        if stage_samples:
            m = [statistics.mean(x[j] for x in stage_samples) for j in range(4)]
            line["stage_ms"] = {"scan": m[0], "merge_incl_wait": m[1], "mmr": m[2], "device_total": m[3],
                                "note": "CUDA events inside the timed e2e calls; at N>1 scan = slowest GPU, merge_incl_wait = root's "
                                        "scan end -> merged pool (waits for the slowest GPU), measured on the root's stream"}
        if throughput is not None:
            line["throughput_mode"] = throughput
        if extras is not None:
            line["extra_configs"] = extras
        if guard is not None:
            guard.restore()
        print(json.dumps(line), flush=True)
        if guard is not None:
            guard.__enter__()
    if world > 1:
        dist.barrier(group=group)
        dist.destroy_process_group()
    if rank == 0 and throughput is not None and throughput.get("parity_ok") is False:
        log("PARITY FAILURE in throughput mode")
        raise SystemExit(3)
    if rank == 0 and parity is not None and not parity["ok"]:
        log("PARITY FAILURE:", *parity["failures"])
        raise SystemExit(3)


def api_throughput_mode(a, target, q_host, wts, callers, singles, api):
    """Throughput mode through the public host-buffer call: `callers` host threads, each call carries Q queries that are
    answered by ONE pass over the rows.  Separate from the single-query headline."""
    out = {"what": "Q independent top_k=%d diversity=%.1f searches per call, answered by ONE scan of the store per GPU; "
                   "queries/s counts queries" % (a.top_k, a.diversity), "api": api, "callers": callers}
    ok = True
    for nq in (2, 3):
        calls = max(4, a.steps // nq)
        errs, lat = [], []
        start = threading.Barrier(callers + 1)

        def caller(t):
            try:
                for i in range(3):
                    target.search_mmr_multi(q_host[[(t + i + j) % N_QUERIES for j in range(nq)]], a.top_k, a.diversity, wts)
                start.wait()
                for i in range(t, calls, callers):
                    idx = [(i * nq + j) % N_QUERIES for j in range(nq)]
                    t0 = time.perf_counter()
                    target.search_mmr_multi(q_host[idx], a.top_k, a.diversity, wts)
                    lat.append(time.perf_counter() - t0)
            except Exception as e:      # noqa: BLE001
                errs.append(repr(e))
                try:
                    start.abort()
                except Exception:       # noqa: BLE001
                    pass

        th = [threading.Thread(target=caller, args=(t,)) for t in range(callers)]
        [t.start() for t in th]
        try:
            start.wait()
        except threading.BrokenBarrierError:
            pass
        t_start = time.perf_counter()
        [t.join() for t in th]
        wall = time.perf_counter() - t_start
        if errs:
            out[f"q{nq}"] = {"error": errs[0]}
            ok = False
            continue
        got = target.search_mmr_multi(q_host[:nq], a.top_k, a.diversity, wts)
        same = all(got[j][0].tobytes() == singles[j][0].tobytes() and got[j][1].tobytes() == singles[j][1].tobytes() for j in range(nq))
        ok &= same
        out[f"q{nq}"] = {"queries_per_pass": nq, "value": calls * nq / wall, "unit": "queries/s",
                         "p50_latency_ms_per_call": 1e3 * statistics.median(lat),
                         "identical_to_single_query_results": bool(same)}
    out["parity_ok"] = bool(ok)
    return out


def run_throughput_mode(a, backend, store, q_dev, q_host, dev, p_cap, w_e, w_l, lanes):
    """Separate from the single-query headline (never folded into `value`): Q = 2 and 3 queries per pass over the rows.
    Each (row, query) dot is still its own sequential f32 chain, so every answer is bit-identical to the single-query
    one -- checked here against the single-query results of the same queries."""
    import numpy as np
    import torch
    from rust_local_rag_b200 import binding as B, dist as rdist, engine
    wts = engine.ResolvedWeights(np.float32(0.7), np.float32(0.3), np.float32(0.7), np.float32(0.3))
    out = {"what": "rlr_search_mmr_multi: Q independent top_k=%d diversity=%.1f searches answered by ONE scan of the store (one group of "
                   "consumer warps per query over the same shared-memory tiles); queries/s counts queries, HBM bytes per query = store/Q" % (a.top_k, a.diversity)}
    ok = True
    singles = [store.search_mmr(q_host[i], a.top_k, a.diversity, wts) for i in range(6)]
    for nq in (2, 3):
        bufs = [[rdist.Buffers(1, p_cap, store.info().pitch, dev) for _ in range(nq)] for _ in range(lanes)]
        streams = [torch.cuda.Stream(dev) for _ in range(lanes)]

        def step(i, lane):
            backend.use_lane(lane)
            qs = [q_dev[(i * nq + j) % N_QUERIES] for j in range(nq)]
            backend.search_mmr_multi(qs, a.top_k, a.diversity, w_e, w_l, [b.result for b in bufs[lane]], [b.sel_n for b in bufs[lane]])

        def run(first, count):
            cur = torch.cuda.current_stream(dev)
            for s_ in streams:
                s_.wait_stream(cur)
            for i in range(count):
                with torch.cuda.stream(streams[i % lanes]):
                    step(first + i, i % lanes)
            for s_ in streams:
                cur.wait_stream(s_)
            backend.use_lane(0)

        calls = max(4, a.steps // nq)
        run(0, max(3, a.warmup))
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(a.warmup, calls)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        # host-buffer API, one caller: latency of a call that carries nq queries
        for i in range(3):
            store.search_mmr_multi(q_host[[j % N_QUERIES for j in range(nq)]], a.top_k, a.diversity, wts)     # leases + warms nq workspaces
        lat = []
        for i in range(calls):
            idx = [(i * nq + j) % N_QUERIES for j in range(nq)]
            t0 = time.perf_counter()
            store.search_mmr_multi(q_host[idx], a.top_k, a.diversity, wts)
            lat.append(time.perf_counter() - t0)
        got = store.search_mmr_multi(q_host[:nq], a.top_k, a.diversity, wts)
        same = all(got[j][0].tobytes() == singles[j][0].tobytes() and got[j][1].tobytes() == singles[j][1].tobytes() for j in range(nq))
        ok &= same
        out[f"q{nq}"] = {"queries_per_pass": nq, "queries_in_flight": nq * lanes,
                         "value": calls * nq / (ms * 1e-3), "unit": "queries/s", "ms_per_pass": ms / calls,
                         "hbm_GBps_algorithmic": a.rows * a.dim * 4 / (ms / calls * 1e-3) / 1e9,
                         "e2e_single_caller": {"value": calls * nq / sum(lat), "unit": "queries/s",
                                               "p50_latency_ms_per_call": 1e3 * statistics.median(lat)},
                         "identical_to_single_query_results": bool(same)}
    out["parity_ok"] = bool(ok)
    return out


def scan_source_sha16():
    h = hashlib.sha256()
    for f in ("scan_topm.cu", "common.cuh", "kernels.cuh", "sort_regs.cuh"):
        h.update(open(os.path.join(ROOT, "rust-local-rag_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


# ----------------------------------------------------------------------------------------
# extra configs
# ----------------------------------------------------------------------------------------
def run_extras(a, rank, world, local_rank, dev, group, peaks, cpu_group=None):
    """BASELINE configs other than the headline one, measured in the same driver-visible run.  Every record
    carries its own `clocks` sample.  N=1: configs 1, 2 and the per-GPU shape of config 4.  N=8: config 4
    sharded over the 8 GPUs and config 5 (100M x 768 f16)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from rust_local_rag_b200 import binding as B, engine, dist as rdist
    from oracle import orc

    out = {}
    wts = engine.ResolvedWeights(np.float32(0.7), np.float32(0.3), np.float32(0.7), np.float32(0.3))
    kw = dict(kind=B.RLR_SYNTH_CLUSTERED, seed=SEED_STORE, centroid_seed=SEED_CENTROID, n_clusters=N_CLUSTERS, sigma=SIGMA)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t.item())

    def queries(n, dim):
        qs = engine.DeviceStore.synthetic(n, dim, device=local_rank, **{**kw, "seed": SEED_QUERY})
        h = qs.read_rows(np.arange(n))
        qs.close()
        return h

    def single_query_config(name, rows_n, dim, k, lam, steps, oracle_rows):
        """configs 1 and 2 on one GPU through rlr_search_mmr; parity: complete results vs the oracle."""
        clk = ClockSampler(local_rank, interval=0.005)
        st = engine.DeviceStore.synthetic(rows_n, dim, device=local_rank, **kw)
        qh = queries(64, dim)
        for i in range(5):
            st.search_mmr(qh[i], k, lam, wts)
        clk.start()
        lat, scan = [], []
        t_start = time.perf_counter()
        for i in range(steps):                     # the latency of a call: no event records, no timing reads
            t0 = time.perf_counter()
            st.search_mmr(qh[i % 64], k, lam, wts)
            lat.append(time.perf_counter() - t0)
        total = time.perf_counter() - t_start
        # the same call with its ctypes arguments built once, as a compiled host would hold them: what is left is the
        # C-ABI call itself (tools/lat_bench.c measures it from C, no interpreter at all)
        import ctypes as C
        cap = max(k, 1)
        o_rows = np.empty(cap, np.uint32); o_sc = np.empty(cap, np.float32); o_em = np.empty(cap, np.float32); o_lx = np.empty(cap, np.float32)
        o_n = C.c_uint32(0)
        wc = B.ResolvedWeightsC(wts.embedding, wts.lexical, wts.reranker, wts.initial)
        fn = st._hot("search_mmr")
        pre = [(st._h, B.ptr(qh[i]), dim, 0, k, float(lam), C.byref(wc), None, None, 0, B.ptr(o_rows), B.ptr(o_sc), B.ptr(o_em), B.ptr(o_lx), C.byref(o_n))
               for i in range(64)]
        lat_raw = []
        for i in range(steps):
            a = pre[i % 64]
            t0 = time.perf_counter()
            rc = fn(*a)
            lat_raw.append(time.perf_counter() - t0)
            if rc != 0:
                raise SystemExit("rlr_search_mmr failed in the prebuilt-argument loop")
        launches = 0
        for i in range(min(steps, 50)):            # stage timings from a separate pass (CUDA events on the launch stream)
            st.search_mmr(qh[i % 64], k, lam, wts, flags=B.RLR_WANT_TIMINGS)
            t = st.last_timings()
            scan.append(t.scan_ms)
            launches = int(t.launches)
        clk.stop()
        # parity: the oracle over the same rows (generated on the host), complete result lists
        ok, compared = True, 0
        if oracle_rows:
            host = orc.synth_rows(rows_n, dim, threads=max(orc.max_threads(), 1), **SYNTH)
            for qi in range(4):
                got = st.search_mmr(qh[qi], k, lam, wts)
                ref = orc.search_with_diversity(host, qh[qi], k, lam, threads=max(orc.max_threads(), 1))
                compared += 1
                ok &= all(x.tobytes() == y.tobytes() for x, y in zip(got[:3], ref[:3]))
            del host
        scan_ms = statistics.mean(scan)
        rec = {"workload": f"{name}: single-query top_k={k} diversity={lam} MMR over {rows_n}x{dim} f32 chunks on 1 GPU, rlr_search_mmr (host buffers)",
               "queries_per_s_e2e": steps / total, "p50_latency_ms": 1e3 * statistics.median(lat),
               "p99_latency_ms": 1e3 * sorted(lat)[int(0.99 * (len(lat) - 1))],
               "p50_latency_ms_prebuilt_ctypes_args": 1e3 * statistics.median(lat_raw),
               "launches_per_query": launches,
               "scan_ms": scan_ms, "scan_GBps": rows_n * dim * 4 / (scan_ms * 1e-3) / 1e9,
               "roofline": {"bound": "hbm", "achieved": rows_n * dim * 4 / (scan_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": rows_n * dim * 4 / (scan_ms * 1e-3) / 1e9 / hbm_peak},
               "parity": {"ok": bool(ok), "oracle_full_results_compared": compared},
               "steps": steps, "clocks": clk.summary()}
        st.close()
        return rec, ok

    all_ok = True
    if world == 1:
        rec, ok = single_query_config("config 1", 10_000, 768, 5, 0.3, 300, True)
        out["config1_10k_k5"] = rec
        all_ok &= ok
        rec, ok = single_query_config("config 2", 1_000_000, 768, 100, 0.7, 100, True)
        out["config2_1m_k100"] = rec
        all_ok &= ok

    # ---- config 1 as a TEXT query: BM25 (LexicalIndex::score) on the device vs on the host, then the same search ----
    if world == 1:
        import random as _random
        from oracle import lexical as olex
        clk = ClockSampler(local_rank, interval=0.005)
        n_t, dim_t, k_t, lam_t = 10_000, 768, 5, 0.3
        rng = _random.Random(1)
        vocab = [f"w{rng.randrange(10**6):06d}" for _ in range(30000)]
        zipf = [1.0 / (i + 1) for i in range(len(vocab))]
        docs = [" ".join(rng.choices(vocab, zipf, k=200)) for _ in range(n_t)]
        tq = [" ".join(rng.choices(vocab, zipf, k=8)) for _ in range(32)]
        st = engine.DeviceStore.synthetic(n_t, dim_t, device=local_rank, **kw)
        qh = queries(32, dim_t)
        dev_ix = engine.DeviceLexicalIndex(st)
        host_ix = engine.LexicalIndex()
        for i, d in enumerate(docs):
            dev_ix.add_chunk(i, d)
            host_ix.add_chunk(str(i), d)
        limit = 5 * max(3 * k_t, k_t + 10)
        terms = [dev_ix.query_terms(q) for q in tq]

        def host_query(i):
            pairs = host_ix.score(tq[i % 32], limit)
            lr = np.array([int(r) for r, _ in pairs], np.uint32); ls = np.array([s_ for _, s_ in pairs], np.float32)
            return st.search_mmr(qh[i % 32], k_t, lam_t, wts, lr, ls)

        def p50(f, n):
            for i in range(5):
                f(i)
            lat = []
            for i in range(n):
                t0 = time.perf_counter()
                f(i)
                lat.append(time.perf_counter() - t0)
            return 1e3 * statistics.median(lat)

        clk.start()
        dev_ms = p50(lambda i: st.search_text_mmr(qh[i % 32], k_t, lam_t, wts, dev_ix.handle, terms[i % 32]), 300)
        clk.stop()
        host_ms = p50(host_query, 20)
        bm_dev_ms = p50(lambda i: dev_ix.score(tq[i % 32], limit), 100)
        bm_host_ms = p50(lambda i: host_ix.score(tq[i % 32], limit), 20)
        # parity: device BM25 + search == host-twin pairs + search (two product paths), and == the oracle's search fed with
        # the pure-Python oracle's BM25 pairs for two queries (the restatement is slow on 10k x 200-token chunks)
        okt = True
        for i in range(8):
            a_, b_ = st.search_text_mmr(qh[i], k_t, lam_t, wts, dev_ix.handle, terms[i]), host_query(i)
            okt &= all(x.tobytes() == y.tobytes() for x, y in zip(a_, b_))
        ref_ix = olex.LexicalIndex()
        for i, d in enumerate(docs):
            ref_ix.add_chunk(i, d)
        host_rows = orc.synth_rows(n_t, dim_t, threads=max(orc.max_threads(), 1), **SYNTH)
        for i in range(2):
            pairs = ref_ix.score(tq[i], limit)
            lr = np.array([r for r, _ in pairs], np.uint32); ls = np.array([s_ for _, s_ in pairs], np.float32)
            want = orc.search_with_diversity(host_rows, qh[i], k_t, lam_t, lex_rows=lr, lex_scores=ls, threads=max(orc.max_threads(), 1))
            got = st.search_text_mmr(qh[i], k_t, lam_t, wts, dev_ix.handle, terms[i])
            okt &= all(x.tobytes() == y.tobytes() for x, y in zip(got, want))
        out["config1_text_query_bm25_on_device"] = {
            "workload": f"search_documents for a TEXT query (8 terms, Zipf vocabulary) over {n_t} chunks of 200 tokens x {dim_t}-d, top_k={k_t} "
                        f"diversity={lam_t}: LexicalIndex::score(query, {limit}) + blend + top-k + MMR",
            "device_bm25": {"p50_latency_ms": dev_ms, "api": "rlr_search_text_mmr: BM25 scoring, top-5k selection, normalisation, scan + blend, MMR on one stream",
                            "lexical_score_only_p50_ms": bm_dev_ms},
            "host_bm25": {"p50_latency_ms": host_ms, "api": "host-mirror LexicalIndex twin (hash-map postings like the reference's) + rlr_search_mmr",
                          "lexical_score_only_p50_ms": bm_host_ms},
            "speedup_whole_query": host_ms / dev_ms,
            "parity": {"ok": bool(okt), "what": "8 queries: device-BM25 path == host-pairs path bit for bit; 2 queries == the oracle's search fed with "
                                                "the pure-Python BM25 restatement's pairs"},
            "clocks": clk.summary()}
        all_ok &= bool(okt)
        dev_ix.close(); host_ix.close(); st.close()
        del host_rows

    # ---- config 4: batched queries on the tensor cores, three operand precisions ----
    if world in (1, 8):
        n4 = 1_250_000 * world
        dim4, nq, m4 = 1024, 1024, 100
        plan = rdist.ShardPlan(n4, world, rank)
        st = engine.DeviceStore.synthetic(plan.n_local, dim4, device=local_rank, row_base=plan.row0,
                                          flags=B.RLR_STORE_KEEP_F16 | B.RLR_STORE_KEEP_BF16, **kw)
        qh = queries(nq, dim4)
        burst, sustained = peaks.get("bf16_tflops", 1636.3), peaks.get("bf16_tflops_sustained", 1371.9)
        sub = min(plan.n_local, 100_000)
        rows_f32 = st.read_rows(np.arange(plan.row0, plan.row0 + sub)) if rank == 0 else None

        def rounded(x, prec):
            if prec == "f16":
                return x.astype(np.float16).astype(np.float64)
            if prec == "bf16":
                return torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16).double().numpy()
            return (np.ascontiguousarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)

        recs = {}
        for prec, pflag in (("f16", B.RLR_BATCH_F16), ("bf16", B.RLR_BATCH_BF16), ("tf32", B.RLR_BATCH_TF32)):
            clk = ClockSampler(local_rank, interval=0.005)
            flags = B.RLR_QUERY_PRENORMALIZED | B.RLR_WANT_TIMINGS | pflag
            for _ in range(3):
                res = rdist.sharded_search_batch(st, group, qh, m4, flags, dev)
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier(group=group)
            clk.start()
            wall, dev_ms = [], []
            for _ in range(20):
                if world > 1:
                    dist.barrier(group=group)
                t0 = time.perf_counter()
                res = rdist.sharded_search_batch(st, group, qh, m4, flags, dev)
                wall.append(time.perf_counter() - t0)
                dev_ms.append(st.last_timings().scan_ms)
            clk.stop()
            wall_s = max_over_ranks(statistics.median(wall))
            gemm_ms = max_over_ranks(statistics.median(dev_ms))
            flop_gpu = 2.0 * nq * plan.n_local * dim4
            tf = flop_gpu / (gemm_ms * 1e-3) / 1e12
            rows_g, scores_g, n_g = res
            if rank == 0:
                worst, ok4 = 0.0, True
                ref = rounded(qh[:8], prec) @ rounded(rows_f32, prec).T
                exact = qh[:8].astype(np.float64) @ rows_f32.astype(np.float64).T
                worst_exact = 0.0
                for q in range(8):
                    inside = (rows_g[q] >= plan.row0) & (rows_g[q] < plan.row0 + sub)
                    sel, got = rows_g[q][inside], scores_g[q][inside]
                    if len(sel):
                        worst = max(worst, float(np.abs(ref[q, sel - plan.row0] - got.astype(np.float64)).max()))
                        worst_exact = max(worst_exact, float(np.abs(exact[q, sel - plan.row0] - got.astype(np.float64)).max()))
                    ok4 &= bool((np.diff(scores_g[q][:n_g[q]].astype(np.float64)) <= 0).all()) and int(n_g[q]) == m4
                ok4 &= worst <= 1e-5
                # the tensor-rate peak of a precision: the measured dense bf16 figures for the 16-bit kinds; tf32 runs at
                # half that rate on this part (no tf32 figure in MEASURED_PEAKS.json: half of bf16, stated)
                pk_b, pk_s = (burst, sustained) if prec != "tf32" else (burst / 2, sustained / 2)
                recs[prec] = {
                    "queries_per_s_e2e": nq / wall_s, "ms_per_batch_e2e": wall_s * 1e3, "contraction_ms_per_gpu": gemm_ms,
                    "flop_per_gpu": flop_gpu, "tflops_per_gpu": tf, "tflops_aggregate": tf * world,
                    "roofline": {"bound": "tensor", "achieved": tf, "peak": pk_b, "unit": "TFLOP/s", "frac": tf / pk_b,
                                 "peak_sustained": pk_s, "frac_of_sustained": tf / pk_s,
                                 "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst) / bf16_tflops_sustained"
                                                + (" halved: tf32 issues at half the bf16 rate" if prec == "tf32" else "")},
                    "parity": {"ok": bool(ok4), "max_abs_dev_from_fp64_contraction_of_rounded_inputs": worst, "stated_tolerance": 1e-5,
                               "max_abs_dev_from_exact_scores": worst_exact},
                    "steps": 20, "clocks": clk.summary()}
                all_ok &= ok4
        if rank == 0:
            out["config4_batched"] = {
                "workload": f"batched {nq} queries x {n4}x{dim4} chunks, top-{m4} per query, tcgen05 cta_group::2 contraction with f32 "
                            f"accumulation in TMEM, rows sharded over {world} GPU(s)"
                            + (" (the per-GPU shape of BASELINE configs[3])" if world == 1 else " (BASELINE configs[3])"),
                "operands": {"f16": "binary16 copy of the store (kind::f16)", "bf16": "bfloat16 copy (kind::f16)",
                             "tf32": "the f32 store itself, no copy (kind::tf32)"},
                "parity_what": "8 queries: every returned row inside a 100k-row slice against the fp64 contraction of the inputs rounded "
                               "as that precision rounds them; lists sorted and full",
                "exchange": "one NCCL all-gather of nq x m u64 keys per rank + per-query device merge" if world > 1 else "none",
                **recs}
        st.close()
        torch.cuda.empty_cache()

    # ---- config 4 again, from ONE process: rlr_cluster_search_batch (rank 0 drives all GPUs; the others wait on the CPU) ----
    if world == 8:
        if rank == 0:
            clk = ClockSampler(local_rank, interval=0.005)
            n4, dim4, nq, m4 = 10_000_000, 1024, 1024, 100
            cl = engine.ClusterStore.synthetic(n4, dim4, devices=list(range(world)), flags=B.RLR_STORE_KEEP_F16,
                                               shard_rows=[rdist.ShardPlan(n4, world, r).n_local for r in range(world)], **kw)
            qh = queries(nq, dim4)
            fl = B.RLR_QUERY_PRENORMALIZED | B.RLR_BATCH_F16
            for _ in range(3):
                res = cl.search_batch(qh, m4, flags=fl)
            clk.start()
            wall = []
            for _ in range(20):
                t0 = time.perf_counter()
                res = cl.search_batch(qh, m4, flags=fl)
                wall.append(time.perf_counter() - t0)
            clk.stop()
            w = statistics.median(wall)
            rows_g, scores_g, n_g = res
            sub = 100_000
            rows_f32 = cl.read_rows(np.arange(sub))
            ref = qh[:8].astype(np.float16).astype(np.float64) @ rows_f32.astype(np.float16).astype(np.float64).T
            worst, okc = 0.0, True
            for q in range(8):
                inside = rows_g[q] < sub
                if inside.any():
                    worst = max(worst, float(np.abs(ref[q, rows_g[q][inside]] - scores_g[q][inside].astype(np.float64)).max()))
                okc &= bool((np.diff(scores_g[q][:n_g[q]].astype(np.float64)) <= 0).all()) and int(n_g[q]) == m4
            okc &= worst <= 1e-5
            out["config4_batched_one_process"] = {
                "workload": f"batched {nq} queries x {n4}x{dim4} chunks, top-{m4}, binary16 operands, 8 GPUs driven by ONE process",
                "api": "rlr_cluster_search_batch (C ABI, host buffers): per-GPU contraction, peer copies of the key lists to the root, per-query device merge",
                "queries_per_s_e2e": nq / w, "ms_per_batch_e2e": w * 1e3, "tflops_aggregate_e2e": 2.0 * nq * n4 * dim4 / w / 1e12,
                "parity": {"ok": bool(okc), "max_abs_dev_from_fp64_contraction_of_rounded_inputs": worst, "stated_tolerance": 1e-5},
                "steps": 20, "clocks": clk.summary()}
            all_ok &= bool(okc)
            cl.close()
            torch.cuda.empty_cache()
        if cpu_group is not None:
            dist.barrier(group=cpu_group)

    # ---- config 5: 100M x 768 binary16 store over 8 GPUs, fused exchange ----
    if world == 8:
        clk = ClockSampler(local_rank)
        dim5, n5, k, lam, p_cap = 768, 100_000_000, 100, 0.7, 300
        w_e, w_l = float(np.float32(0.7)), float(np.float32(0.3))
        head = rdist.ShardPlan.balanced_head_rows(n5, world, 0.1 * 7.0e9 / (dim5 * 2))
        plan = rdist.ShardPlan(n5, world, rank, head_rows=head)
        st = engine.DeviceStore.synthetic(plan.n_local, dim5, device=local_rank, row_base=plan.row0, flags=B.RLR_STORE_F16_ONLY, **kw)
        be = rdist.CudaBackend(st, dev)
        be.open_peers(group, plan)
        be.open_mailbox(group, m_cap=p_cap, ring=4)
        bufs = rdist.Buffers(world, p_cap, st.info().pitch, dev)
        qh = queries(32, dim5)
        qd = torch.zeros((32, B.RLR_MAX_DIM + 64), device=dev)
        qd[:, :dim5] = torch.from_numpy(qh).to(dev)
        for i in range(5):
            rdist.sharded_search(be, group, bufs, qd[i], k, lam, w_e, w_l)
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)
        clk.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(40):
            rdist.sharded_search(be, group, bufs, qd[(5 + i) % 32], k, lam, w_e, w_l)
        e1.record()
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)
        clk.stop()
        ms = max_over_ranks(e0.elapsed_time(e1)) / 40
        res, res_n = rdist.sharded_search(be, group, bufs, qd[0], k, lam, w_e, w_l)
        torch.cuda.synchronize(dev)
        if rank == 0:
            got_rows, got_score, got_emb, _ = rdist.decode_result(res, int(res_n.item()))
            ok5, worst = len(got_rows) == k, 0.0
            for r, e in zip(got_rows, got_emb):
                row32 = orc.synth_rows(1, dim5, row0=int(r), threads=1, **SYNTH)[0]
                row16 = row32.astype(np.float16).astype(np.float32)
                ok5 &= np.float32(orc.dot(qh[0], row16)).tobytes() == np.float32(e).tobytes()
                worst = max(worst, abs(float(orc.dot(qh[0], row32)) - float(e)))
            bytes_gpu = (n5 - head) // (world - 1) * dim5 * 2
            out["config5_100m_f16"] = {
                "workload": f"single-query top_k={k} diversity={lam} MMR over {n5}x{dim5} binary16 chunks ({n5 * dim5 * 2 / 1e9:.1f} GB), "
                            f"rows sharded over {world} GPUs, fused exchange, tail-balanced (BASELINE configs[4])",
                "queries_per_s": 1e3 / ms, "ms_per_query": ms,
                "roofline": {"bound": "hbm", "achieved": bytes_gpu / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": bytes_gpu / (ms * 1e-3) / 1e9 / hbm_peak,
                             "note": "per GPU: one non-root shard's bytes / whole step time (lower bound on the scan kernel's rate)"},
                "parity": {"ok": bool(ok5), "what": "every selected row regenerated on the CPU, rounded to binary16 and re-scored by the "
                                                    "oracle: bit-equal (the f16 store is the reference run on the rounded rows)",
                           "max_abs_deviation_from_f32_scores": worst, "stated_f16_tolerance": 2e-4,
                           "mailbox_timeouts": be.mailbox_status()},
                "steps": 40, "clocks": clk.summary()}
            all_ok &= bool(ok5)
        be.close(group)
        dist.barrier(group=group)
        st.close()
    # ---- the 10M x 768 store as a TEXT corpus: BM25 scored on the 8 GPUs + the same search, ONE process (rank 0 runs
    # tools/cluster_text_time.py in a child process; every rank's own stores are closed by now) ----
    if world == 8:
        if rank == 0:
            import subprocess
            torch.cuda.empty_cache()
            try:
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cluster_text_time.py"), "10000000", "8", "768"],
                                   capture_output=True, text=True, timeout=600, cwd=ROOT)
                rec = json.loads(r.stdout.strip().splitlines()[-1])
                rec["exit_code"] = r.returncode
                out["config3_text_query_bm25_on_8_gpus"] = rec
                all_ok &= bool(rec.get("parity_ok")) and r.returncode == 0
            except Exception as e:                     # reported, and it fails the run's parity flag
                out["config3_text_query_bm25_on_8_gpus"] = {"error": repr(e)[:300]}
                all_ok = False
        if cpu_group is not None:
            dist.barrier(group=cpu_group)
    if rank == 0:
        out["all_parity_ok"] = bool(all_ok)
    return out


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
