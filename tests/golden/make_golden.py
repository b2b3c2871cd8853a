"""Regenerates tests/golden/search_mmr_golden.json.

The reference is Rust and cannot be run in this image (no cargo/rustc), so these vectors do NOT come
from the reference: they are outputs of the CPU oracle (oracle/rlr_oracle.c, itself pinned on the
reference's own known-answer tests in tests/test_oracle_kat.py), frozen here so that any later drift
of the oracle OR of the CUDA path shows up as a diff against a committed file.  Inputs are
reproducible from the recorded parameters (counter-hash synthetic generator); outputs are stored as
raw f32 bit patterns.

    python tests/golden/make_golden.py        # rewrites the JSON next to this script
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

CASES = [
    dict(name="iid_small", n=64, dim=16, kind=0, n_clusters=1, seed=101, qseed=201),
    dict(name="clustered_300x48", n=300, dim=48, kind=1, n_clusters=8, seed=102, qseed=202),
    dict(name="clustered_2000x768", n=2000, dim=768, kind=1, n_clusters=32, seed=103, qseed=203),
    dict(name="odd_dim_129x33", n=129, dim=33, kind=0, n_clusters=1, seed=104, qseed=204),
]
QUERIES = [(5, 0.3), (5, 0.0), (100, 0.7), (0, 0.5), (7, 1.0)]


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32).tolist()


def main():
    out = {"note": "oracle-generated (NOT reference-generated) regression vectors; see make_golden.py", "cases": []}
    for c in CASES:
        kw = dict(kind=c["kind"], seed=c["seed"], centroid_seed=c["seed"] + 1000, n_clusters=c["n_clusters"], sigma=0.65)
        rows = orc.synth_rows(c["n"], c["dim"], threads=1, **kw)
        q = orc.synth_rows(1, c["dim"], threads=1, **{**kw, "seed": c["qseed"]})[0]
        lex_rows = np.array([3, 17, 40, c["n"] - 1], np.uint32)
        lex_scores = np.array([2.5, 0.75, 4.0, 1.25], np.float32)
        rec = dict(c, first_row_bits=bits(rows[0][:8]), query_bits=bits(q), lex_rows=lex_rows.tolist(),
                   lex_score_bits=bits(lex_scores), results=[])
        for k, lam in QUERIES:
            for use_lex in (False, True):
                r, s, e, l = orc.search_with_diversity(rows, q, k, lam, normalize_query=False, full_sort=True,
                                                       lex_rows=lex_rows if use_lex else None,
                                                       lex_scores=lex_scores if use_lex else None)
                rec["results"].append(dict(top_k=k, diversity=lam, lexical=use_lex, rows=r.tolist(), score_bits=bits(s),
                                           emb_bits=bits(e), lex_bits=bits(l)))
        tr, ts, te, tl = orc.search(rows, q, 45, normalize_query=False, full_sort=True)
        rec["topm_45"] = dict(rows=tr.tolist(), score_bits=bits(ts), emb_bits=bits(te))
        out["cases"].append(rec)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "search_mmr_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
