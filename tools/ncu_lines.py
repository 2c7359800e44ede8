"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > x.csv ; python tools/ncu_lines.py x.csv [top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, h, data = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        h = r
    elif h and r[0].isdigit() and len(r) > 8 and r[2] == "-":
        isamp, iinst = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
        stalls = {k: int(r[h.index(k)] or 0) for k in h if k.startswith("stall_") and "(" not in k}
        data.append((int(r[isamp] or 0), int(r[iinst] or 0), cur, int(r[0]), r[1].strip()[:80], stalls))
tot, toti = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
print("total samples", tot, "total warp-instructions", toti)
for d in sorted(data, key=lambda d: -d[0])[:top]:
    st = sorted(d[5].items(), key=lambda kv: -kv[1])[:2]
    print(f"{100 * d[0] / tot:5.1f}% samp {100 * d[1] / toti:5.1f}% inst  {d[2]}:{d[3]:<4} {d[4]:<80} {st}")
