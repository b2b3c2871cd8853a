"""The Rust -sys crate cannot be compiled here (no cargo/rustc).  The least we can do is keep
its extern block textually in sync with include/rlr_b200.h: same symbol set, same arity."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_decls():
    text = open(os.path.join(ROOT, "include", "rlr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(rlr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def _rust_decls():
    text = open(os.path.join(ROOT, "rust", "rlr-b200-sys", "src", "lib.rs")).read()
    out = {}
    for m in re.finditer(r"pub fn (rlr_[a-z0-9_]+)\(([^)]*)\)", text):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if not args else args.count(":")
    return out


def test_rust_extern_block_matches_header():
    h, r = _header_decls(), _rust_decls()
    assert set(h) == set(r), (sorted(set(h) - set(r)), sorted(set(r) - set(h)))
    for name in h:
        assert h[name] == r[name], (name, h[name], r[name])


def test_build_rs_compiles_the_same_sources_as_the_python_build():
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import _build
    text = open(os.path.join(ROOT, "rust", "rlr-b200-sys", "build.rs")).read()
    m = re.search(r"let sources = \[([^\]]*)\]", text)
    assert m, "sources list not found in build.rs"
    rust_sources = re.findall(r'"([^"]+)"', m.group(1))
    assert rust_sources == list(_build.SOURCES)


def _header_defines():
    text = open(os.path.join(ROOT, "include", "rlr_b200.h")).read()
    out = {}
    for m in re.finditer(r"^#define\s+(RLR_[A-Z0-9_]+)\s+(0x[0-9a-fA-F]+|\d+)u?\b", text, flags=re.M):
        out[m.group(1)] = int(m.group(2), 0)
    out.pop("RLR_B200_H", None)
    return out


def test_constants_agree_between_header_rust_and_python():
    from rust_local_rag_b200 import binding as B
    h = _header_defines()
    assert len(h) >= 20
    text = open(os.path.join(ROOT, "rust", "rlr-b200-sys", "src", "lib.rs")).read()
    r = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"pub const (RLR_[A-Z0-9_]+): \w+ = (0x[0-9a-fA-F]+|\d+);", text)}
    assert set(h) == set(r), (sorted(set(h) - set(r)), sorted(set(r) - set(h)))
    for k, v in h.items():
        assert r[k] == v, k
        if hasattr(B, k):
            assert getattr(B, k) == v, k
    for k in ("RLR_STORE_NORMALIZE_ON_UPLOAD", "RLR_BATCH_EXACT_RESCORE", "RLR_SEARCH_F16", "RLR_MAX_M", "RLR_IPC_HANDLE_BYTES"):
        assert hasattr(B, k), k
