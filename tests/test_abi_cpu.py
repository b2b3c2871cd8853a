"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/rlr_b200.h
declares, its host-only helpers are bit-exact against the oracle, and -- with no GPU --
every compute entry point fails loudly instead of falling back."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rlr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rlr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound(rlr):
    lib = rlr.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rlr_b200.h but not exported"
        assert n in rlr.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(rlr.PROTOTYPES) == names
    assert lib.rlr_abi_version() == 1


def test_only_sm100a_code_in_library(rlr):
    import subprocess
    from rust_local_rag_b200 import _build
    out = subprocess.run(["cuobjdump", "-lelf", _build.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_host_normalize_bit_exact_vs_oracle(rlr, orc):
    lib = rlr.load()
    rng = np.random.default_rng(0)
    for dim in (1, 3, 384, 768, 1024):
        v = rng.standard_normal(dim).astype(np.float32)
        w = v.copy()
        rlr.check(lib.rlr_normalize(w.ctypes.data_as(C.POINTER(C.c_float)), dim))
        assert w.tobytes() == orc.normalize(v).tobytes()
    z = np.zeros(8, np.float32)
    rlr.check(lib.rlr_normalize(z.ctypes.data_as(C.POINTER(C.c_float)), 8))
    assert (z == 0).all()


def test_resolve_weights_matches_reference_tests(rlr, orc):
    from rust_local_rag_b200.engine import QueryWeights, resolve_weights
    d = resolve_weights(None)                                   # :3115-3125 defaults
    assert (d.embedding, d.lexical, d.reranker, d.initial) == tuple(np.float32(x) for x in (0.7, 0.3, 0.7, 0.3))
    r = resolve_weights(QueryWeights(embedding=0.9))            # :3140-3157 partial override
    assert r.embedding == np.float32(0.9) and r.lexical == d.lexical
    r = resolve_weights(QueryWeights(embedding=0.8, lexical=0.2, initial=0.4))  # :3159-3174
    assert (r.embedding, r.lexical, r.reranker, r.initial) == (np.float32(0.8), np.float32(0.2), d.reranker, np.float32(0.4))
    r = resolve_weights(QueryWeights(embedding=math.nan, lexical=-0.1, reranker=0.6, initial=1.5))  # :3176-3194
    assert (r.embedding, r.lexical, r.reranker, r.initial) == (d.embedding, d.lexical, np.float32(0.6), d.initial)
    for bad in (math.inf, -math.inf, -1e-6, 1.000001, 2.0):
        assert resolve_weights(QueryWeights(lexical=bad)).lexical == d.lexical
    assert resolve_weights(QueryWeights(embedding=0.0)).embedding == 0.0
    assert resolve_weights(QueryWeights(embedding=1.0)).embedding == 1.0
    for v in (0.5, 0.9, 0.0, 1.0, -0.0, math.nan, 1.5):
        assert resolve_weights(QueryWeights(initial=v)).initial == np.float32(orc.resolve_weight(v, 0.3))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="checks the no-device failure mode")
def test_no_device_fails_loudly_no_cpu_fallback(rlr):
    lib = rlr.load()
    rows = np.eye(4, dtype=np.float32)
    h = C.c_void_p()
    rc = lib.rlr_store_create(0, 4, 4, rows.ctypes.data_as(C.c_void_p), 4, 0, 0, C.byref(h))
    assert rc == rlr.RLR_ERR_NO_DEVICE
    assert not h.value
    assert b"no CPU fallback" in lib.rlr_last_error() or b"CUDA" in lib.rlr_last_error()
    from rust_local_rag_b200.engine import RagEngine
    with pytest.raises(rlr.RlrError) as ei:
        RagEngine.from_rows(rows)
    assert ei.value.code == rlr.RLR_ERR_NO_DEVICE


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "rust-local-rag_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "liborc" not in txt and "import orc" not in txt and "from oracle" not in txt, f
