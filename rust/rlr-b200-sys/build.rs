// build.rs -- drives nvcc for sm_100a and links the resulting shared library.
// NOT COMPILED HERE (no Rust toolchain in the image); mirrors rust-local-rag_b200/_build.py.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("rust-local-rag_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("librlr_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let sources = ["api.cu", "cluster.cu", "scan_topm.cu", "merge.cu", "mmr.cu", "synth.cu", "batch_gemm.cu", "bm25.cu"];
    let status = Command::new(&nvcc)
        .current_dir(&csrc)
        .args(["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false"])
        .args(["-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off", "-shared", "-o"])
        .arg(&lib)
        .args(sources)
        .status()
        .expect("failed to run nvcc");
    assert!(status.success(), "nvcc failed");
    for s in sources.iter().chain(["common.cuh", "kernels.cuh", "sort_regs.cuh", "mmr_device.cuh", "api_internal.hpp"].iter()) {
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/rlr_b200.h").display());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=rlr_b200");
}
