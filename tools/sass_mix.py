#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` dump: executed-instruction mix by SASS opcode and stall reasons.
usage: tools/sass_mix.py file.csv [kernel-substring]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else None
op = collections.Counter(); samp = collections.Counter(); stall = collections.Counter()
tot = totS = 0
hdr = None; active = True
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        active = (want is None) or (want in r[1]); continue
    if len(r) > 5 and r[0] == "Address":
        hdr = r; idx = {h: i for i, h in enumerate(hdr)}
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or not active or len(r) < len(hdr):
        continue
    sass = r[idx["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
    o = m.group(2).split(".")[0] if m else "?"
    try:
        n = int(r[idx["Instructions Executed"]] or 0); s = int(r[idx["# Samples"]] or 0)
    except ValueError:
        continue
    op[o] += n; samp[o] += s; tot += n; totS += s
    for c in stall_cols:
        stall[c] += int(r[idx[c]] or 0)
print("total warp instr", tot, "samples", totS)
for o, n in op.most_common(30):
    print(f"{o:10s} {n:12d} {100*n/max(tot,1):5.1f}%  samples {100*samp[o]/max(totS,1):5.1f}%")
print({k: round(100 * v / max(totS, 1), 1) for k, v in stall.most_common(12)})
