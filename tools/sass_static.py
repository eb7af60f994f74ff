#!/usr/bin/env python3
"""Static SASS opcode histogram of one kernel in the built library: tools/sass_static.py <regex> [lib.so]"""
import collections, re, subprocess, sys, glob, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(root, "quantizationawarethzdoe_b200", "csrc", "libthzdoe.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(sys.argv[1])
cur, hist = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1) if pat.search(m.group(1)) else None
        if cur: hist[cur] = collections.Counter()
        continue
    if cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m: hist[cur][m.group(1)] += 1
for k, h in hist.items():
    tot = sum(h.values())
    print(k, "total", tot)
    print("   ", ", ".join("%s %d" % kv for kv in h.most_common(14)))
