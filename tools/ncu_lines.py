#!/usr/bin/env python3
"""Per-source-line stall attribution from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name ... > f.csv`.
usage: tools/ncu_lines.py f.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        d["file"] = cur
        out.append(d)
def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0
key = "Warp Stall Sampling (All Samples)"
tot = sum(f(d[key]) for d in out)
insts = sum(f(d["Instructions Executed"]) for d in out)
print("total samples %d, warp instructions %d" % (tot, insts))
out.sort(key=lambda d: -f(d[key]))
for d in out[:top]:
    stalls = {k[6:]: f(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and f(v) > 0}
    st = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
    print("%5.1f%% inst %4.1f%%  %-18s:%-4s %-64s %s" % (100 * f(d[key]) / tot, 100 * f(d["Instructions Executed"]) / insts, d["file"], d["Line No"],
                                                    d["Source"].strip()[:64], " ".join("%s=%d" % (k, v) for k, v in st)))
