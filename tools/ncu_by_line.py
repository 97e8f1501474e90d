"""Joins an ncu source-page CSV (per-SASS-instruction counts) with nvdisasm line info and
prints per-source-line totals: warp instructions, thread instructions, SIMT efficiency, stall samples.
usage: ncu_by_line.py src.csv dis.txt kernel_mangled_substring"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis, kern = sys.argv[1:4]
# nvdisasm: track current file/line for each instruction offset within the kernel section
line_of = {}
cur = None
in_k = False
for ln in open(dis):
    if ln.startswith("//-") and ".text." in ln:
        in_k = kern in ln
        continue
    if not in_k:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        inl = "inlined" in ln
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
base = None
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    addr = int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]])
    if base is None:
        base = addr
    off = addr - base
    ie = int(float(r[ix["Instructions Executed"]] or 0))
    te = int(float(r[ix["Thread Instructions Executed"]] or 0))
    ss = int(float(r[ix["# Samples"]] or 0))
    key = line_of.get(off, ("?", 0))
    a = agg[key]
    a[0] += ie; a[1] += te; a[2] += ss
    tot[0] += ie; tot[1] += te; tot[2] += ss
print(f"total warp-inst {tot[0]:.3e} thread-inst {tot[1]:.3e} eff {tot[1]/tot[0]/32:.3f} samples {tot[2]}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 45]:
    print(f"{key[0]}:{key[1]:<5d} warp-inst {100*a[0]/tot[0]:5.1f}%  eff {a[1]/max(a[0],1)/32:5.2f}  stall-samples {100*a[2]/max(tot[2],1):5.1f}%")
