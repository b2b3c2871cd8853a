"""ncu raw page CSV of ONE scan_topm_kernel launch -> profiles/r02_traffic.json (dram bytes of that launch).
bench.py reports them as `roofline.traffic` only when the launch shape AND the scan kernel's source hash match the
running tree, so a stale file can never annotate a different kernel.
usage: make_traffic.py <raw.csv> <rows> <dim> <out.json>"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def scan_source_sha16():
    h = hashlib.sha256()
    for f in ("scan_topm.cu", "common.cuh", "kernels.cuh", "sort_regs.cuh"):
        h.update(open(os.path.join(ROOT, "rust-local-rag_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def main():
    raw, rows, dim, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    r = list(csv.reader(open(raw)))
    hdr_i = next(i for i, x in enumerate(r) if x and x[0] == "ID")
    hdr, units, vals = r[hdr_i], r[hdr_i + 1], r[hdr_i + 2]

    def get(name):
        i = hdr.index(name)
        v = float(vals[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)

    d = {"scan_topm_kernel": {"rows": rows, "dim": dim, "elem_bytes": 4,
                              "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
                              "kernel": vals[hdr.index("Kernel Name")], "duration_ns": get("gpu__time_duration.sum"),
                              "scan_src_sha16": scan_source_sha16(),
                              "source": f"{os.path.basename(raw)} (ncu --set full, one launch, {rows} x {dim} f32 on one B200)"}}
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d))


if __name__ == "__main__":
    main()
