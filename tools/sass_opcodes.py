"""Per-kernel SASS opcode summary of the built product library (evidence for profiles/):

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt

Counts, per kernel in rust-local-rag_b200/librlr_b200.so, of the instructions that show what the kernel is made of:
UTMALDG (TMA tensor loads), UBLKCP (bulk async copies), UTCHMMA / UTCQMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR
(tcgen05.commit), SYNCS (mbarrier), LDGSTS (cp.async), REDUX, and FFMA vs FMUL/FADD (the exact-order kernels must
hold NO FFMA in their dot chains), plus the `arch =` lines of the fatbin."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rust-local-rag_b200", "librlr_b200.so")
OPS = ["UTMALDG", "UBLKCP", "UTCHMMA", "UTCQMMA", "LDTM", "UTCBAR", "SYNCS", "LDGSTS", "REDUX", "FFMA", "FMUL", "FADD",
       "HFMA2", "LDS", "LDG", "STG", "ATOMS", "ATOMG", "RED", "NANOSLEEP", "MEMBAR", "ACQBULK", "UCGABAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r"arch = (\S+)", out)))
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS images for {archs}")
    print("# exact-order kernels (scan_topm, mmr_pairwise, mmr_greedy, batch_rescore): FMUL + FADD chains, 0 FFMA.")
    print("# FFMA in normalize_* / synth_kernel sits inside the IEEE-correct __fdiv_rn / __fsqrt_rn expansions (results are")
    print("# bit-identical to the oracle: tests/test_gpu_parity.py).  Both batch_gemm*<true> (kind::tf32) and <false> (kind::f16)")
    print("# issue UTCHMMA: the operand format travels in the instruction descriptor.")
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            cur = re.sub(r"\(.*$", "", cur)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    counts[cur][o] += 1
            counts[cur]["_total"] += 1
    used = [o for o in OPS if any(c[o] for c in counts.values())]
    w = max(len(k) for k in counts) + 2
    print("kernel".ljust(w) + "".join(o.rjust(9) for o in ["insts"] + used))
    for k, c in counts.items():
        print(k.ljust(w) + str(c["_total"]).rjust(9) + "".join(str(c[o]).rjust(9) for o in used))


if __name__ == "__main__":
    sys.exit(main())
