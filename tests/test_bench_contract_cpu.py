"""bench.py's output contract: the reference arm runs on CPU (small --rows here) and prints ONE JSON
line with the agreed keys; the committed bench lines under profiles/ carry the keys of the own arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")          # what torchrun exports: the arm must still use every core
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "30000",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["unit"] == "queries/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="3", WORLD_SIZE="8", LOCAL_RANK="3")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8", "--rows", "1000",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_own_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode != 0 and res.stdout.strip() == ""
    assert "no CPU fallback" in res.stderr or "CUDA" in res.stderr


def test_committed_bench_lines_carry_the_contract_keys():
    prof = os.path.join(ROOT, "profiles")
    for name in ("r01_bench_n1_default.json", "r01_bench_n2.json", "r01_bench_n4.json", "r01_bench_n8.json"):
        d = json.loads(open(os.path.join(prof, name)).read())
        assert BASE_KEYS | {"roofline", "clocks"} <= set(d), name
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
        assert d["roofline"]["bound"] == "hbm" and abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
        assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"]))
        assert d["scaling"] == "strong" and d["dtype"] == "f32" and d["data"] == "synthetic"
        if d["n_gpus"] == 1:
            assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
