"""Summarise an `ncu --page source --csv` dump: stall samples and executed instructions by opcode, plus the hottest instructions.
usage: python tools/ncu_src_summary.py dump.csv [ntop]"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; data = [r for r in rows[1:] if r[1] != "Source"]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
tot = sum(int(r[iS]) for r in data); totI = sum(int(r[iI]) for r in data)
print("samples", tot, "warp instructions", totI)
g = defaultdict(lambda: [0, 0])
for r in data:
    op = r[iSrc].split()
    o = (op[0] if not op[0].startswith('@') else op[1]).split('.')[0]
    g[o][0] += int(r[iS]); g[o][1] += int(r[iI])
for k, v in sorted(g.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"{k:10s} samples {v[0]:7d} ({100*v[0]/max(tot,1):5.1f}%)  inst {v[1]:10d} ({100*v[1]/max(totI,1):5.1f}%)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for r in sorted(data, key=lambda r: -int(r[iS]))[:ntop]:
    st = sorted(((int(r[i]), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(r[0][-5:], r[iSrc][:64].ljust(64), r[iS], r[iI], st)
