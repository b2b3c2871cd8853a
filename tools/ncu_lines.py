"""Dev tool: attribute ncu warp-stall samples (SASS source page CSV) to CUDA source lines using
nvdisasm -g line info.  usage: ncu_lines.py <sass_page.csv> <nvdisasm_-g.txt> <kernel-substring> [top]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# address -> (file, line)
amap, cur, infn = {}, None, False
for l in open(dis):
    if l.startswith(".text."):
        infn = kname in l
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        amap[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isamp, isrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
by_line, total = defaultdict(lambda: [0, defaultdict(int)]), 0
base = None
for r in rows[hi + 1:]:
    if len(r) != len(hdr):
        continue
    a = int(r[ia], 16) if not r[ia].isdigit() else int(r[ia])
    if base is None:
        base = a
    s = int(r[isamp] or 0)
    total += s
    key = amap.get(a - base, (("?", 0), ""))[0]
    by_line[key][0] += s
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            by_line[key][1][hdr[i]] += v
print("total samples", total)
for key, (s, st) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{key[0]:24s} L{key[1]:4d} {s:8d} {100 * s / max(total, 1):5.1f}%  {tops}")
