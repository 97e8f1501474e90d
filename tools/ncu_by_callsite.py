"""Like ncu_by_line.py, but attributes every SASS instruction to the OUTERMOST source line of the kernel
body it was inlined into (nvdisasm -gi), so library math (umath.h) is charged to the phase that called it.
usage: ncu_by_callsite.py src.csv dis_gi.txt kernel_mangled_substring kernel_file.cuh [top]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis, kern, kfile = sys.argv[1:5]
line_of = {}
cur = None
in_k = False
for ln in open(dis):
    if ln.startswith("//-") and ".text." in ln:
        in_k = kern in ln
        continue
    if not in_k:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        f, l, of, ol = m.group(1), int(m.group(2)), m.group(3), m.group(4)
        if of and of.endswith(kfile):
            cur = int(ol)
        elif f.endswith(kfile) and not of:
            cur = l
        elif of:
            cur = cur  # nested deeper than one level: keep the enclosing call site
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
    ie = int(float(r[ix["Instructions Executed"]] or 0))
    te = int(float(r[ix["Thread Instructions Executed"]] or 0))
    ss = int(float(r[ix["# Samples"]] or 0))
    a = agg[line_of.get(addr - base)]
    a[0] += ie; a[1] += te; a[2] += ss
    tot[0] += ie; tot[1] += te; tot[2] += ss
print(f"total warp-inst {tot[0]:.3e} thread-inst {tot[1]:.3e} eff {tot[1]/tot[0]/32:.3f}")
src = open(sys.argv[6]).read().splitlines() if len(sys.argv) > 6 else None
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 40]:
    text = src[key - 1].strip()[:90] if src and key and key <= len(src) else ""
    print(f"{kfile}:{key}  warp-inst {100*a[0]/tot[0]:5.1f}%  eff {a[1]/max(a[0],1)/32:5.2f}  stall-samples {100*a[2]/max(tot[2],1):5.1f}%   {text}")
