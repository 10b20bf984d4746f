#!/usr/bin/env python
"""Aggregate an ncu SASS source page by CUDA source line.

    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<kernel> > sass.csv
    cuobjdump -xelf all libaec.so.0 ; nvdisasm -g -c <cubin> > dis.txt
    python ncu_by_line.py sass.csv dis.txt <mangled-kernel-substring> [top]

ncu's CSV export carries per-SASS-instruction counters but no line numbers;
nvdisasm -g prints the same instructions in the same order with `//## File ...
line N` markers.  The two listings are joined by instruction index.
"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, dis_txt, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ci = {n: hdr.index(n) for n in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples")}
ins = [(r[ci["Source"]].strip(), int(r[ci["Instructions Executed"]] or 0), int(r[ci["Thread Instructions Executed"]] or 0),
        int(r[ci["# Samples"]] or 0)) for r in rows[2:] if len(r) > ci["# Samples"]]

lines = open(dis_txt, errors="ignore").read().split("\n")
# locate the function
start = None
for i, l in enumerate(lines):
    if l.startswith(".text.") and kern in l:
        start = i
        break
    if re.match(r"\s*\.section\s+\.text\.", l) and kern in l:
        start = i
        break
assert start is not None, "kernel not found in disassembly"
cur = ("?", 0)
maps = []
for l in lines[start + 1:]:
    if re.match(r"\s*\.section\s", l) and maps:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        maps.append((cur, m.group(2)))
print("sass instrs: ncu %d, nvdisasm %d" % (len(ins), len(maps)))
n = min(len(ins), len(maps))
agg = defaultdict(lambda: [0, 0, 0, 0])
for k in range(n):
    key = maps[k][0]
    a = agg[key]
    a[0] += ins[k][1]; a[1] += ins[k][2]; a[2] += ins[k][3]; a[3] += 1
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[2] for v in agg.values()) or 1
print("total warp instructions executed: %d, samples %d" % (tot, tots))
print("%-26s %6s %12s %7s %7s %6s" % ("file:line", "sass", "warp-inst", "inst%", "smpl%", "thr/w"))
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-26s %6d %12d %6.2f%% %6.2f%% %6.1f" % ("%s:%d" % key, v[3], v[0], 100.0 * v[0] / tot, 100.0 * v[2] / tots,
                                                 v[1] / v[0] if v[0] else 0))
