#!/usr/bin/env python3
"""Aggregate an ncu --import-source capture by SASS opcode: executed warp-instructions per opcode.
usage: ncu_opcodes.py file.ncu-rep [top]"""
import csv, io, subprocess, sys
from collections import Counter
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = [l for l in out.splitlines(True) if l.startswith('"') and not l.startswith('"Kernel Name"')]
rd = csv.DictReader(io.StringIO("".join(lines)))
cnt = Counter(); stall = Counter(); samples = Counter()
tot = 0
for r in rd:
    src = r.get("Source") or r.get("SASS") or ""
    toks = src.replace("@P0", "").replace("@!P0", "").split()
    toks = [t for t in toks if not t.startswith("@")]
    if not toks: continue
    op = toks[0].split(".")[0] + ("." + ".".join(toks[0].split(".")[1:2]) if "." in toks[0] else "")
    try: n = int(float(r.get("Instructions Executed") or 0))
    except ValueError: n = 0
    cnt[op] += n; tot += n
    try: samples[op] += int(float(r.get("# Samples") or 0))
    except ValueError: pass
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print(f"total warp-instructions executed: {tot}")
for op, n in cnt.most_common(top):
    print(f"{op:24s} {n:16d} {100*n/tot:6.2f}%   stall-samples {100*samples[op]/max(1,sum(samples.values())):6.2f}%")
