"""In-tree build of the sm_100a shared library (no JIT cache: the built .so travels with
the repo snapshot to the GPU box)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librlr_b200.so")
SOURCES = ["api.cu", "cluster.cu", "scan_topm.cu", "merge.cu", "mmr.cu", "synth.cu", "batch_gemm.cu", "bm25.cu"]
HEADERS = ["common.cuh", "kernels.cuh", "sort_regs.cuh", "mmr_device.cuh", "api_internal.hpp", os.path.join("..", "..", "include", "rlr_b200.h")]

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",                 # belt and braces: exact paths use __fmul_rn/__fadd_rn anyway
    "-ccbin", "/usr/bin/g++",      # $CC in this image is a wrapper without OpenMP specs
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off",
    "-shared",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> librlr_b200.so with nvcc for sm_100a.  Returns the path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


# ---- host-mirror support (NOT the product): the BM25 / tokenizer twin the non-Rust host mirrors use for text queries
HM_DIR = os.path.join(HERE, "host_mirror")
HM_LIB = os.path.join(HERE, "librlr_hostmirror.so")
HM_DEPS = [os.path.join(HM_DIR, "lexical.cpp"), os.path.join(HM_DIR, "unicode_tables.inc"),
           os.path.join(HERE, "..", "include", "rlr_hostmirror.h")]


def build_hostmirror(force: bool = False) -> str:
    """g++ host_mirror/lexical.cpp -> librlr_hostmirror.so (plain C++17, no CUDA)."""
    if not force and os.path.exists(HM_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(HM_LIB) for d in HM_DEPS):
        return HM_LIB
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-ffp-contract=off",
           "-o", HM_LIB, HM_DEPS[0]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return HM_LIB
