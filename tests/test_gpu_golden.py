"""CUDA path against the committed golden vectors (see tests/test_golden_cpu.py)."""
import numpy as np
import pytest

from test_golden_cpu import inputs, load_cases

pytestmark = pytest.mark.gpu
F32 = np.float32


def test_cuda_path_reproduces_golden(orc):
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import binding as B, engine
    w = engine.ResolvedWeights(F32(0.7), F32(0.3), F32(0.7), F32(0.3))
    for c in load_cases():
        rows, q, lr, ls = inputs(orc, c)
        s = engine.DeviceStore.from_rows(rows)
        for g in c["results"]:
            r, sc, e, l = s.search_mmr(q, g["top_k"], g["diversity"], w, lr if g["lexical"] else None,
                                       ls if g["lexical"] else None, flags=B.RLR_QUERY_PRENORMALIZED)
            assert r.tolist() == g["rows"], (c["name"], g["top_k"], g["diversity"], g["lexical"])
            assert sc.view(np.uint32).tolist() == g["score_bits"] and e.view(np.uint32).tolist() == g["emb_bits"]
            assert l.view(np.uint32).tolist() == g["lex_bits"]
        tr, ts, te, _ = s.search_topm(q, 45, w, flags=B.RLR_QUERY_PRENORMALIZED)
        assert tr.tolist() == c["topm_45"]["rows"] and ts.view(np.uint32).tolist() == c["topm_45"]["score_bits"]
        assert te.view(np.uint32).tolist() == c["topm_45"]["emb_bits"]
        s.close()
