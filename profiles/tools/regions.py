#!/usr/bin/env python
"""Sum an ncu_by_line.py listing over named source regions.

    python regions.py by_line.txt units  file:a-b:name ...
`units` = how many warp-level work items the launch had (e.g. tiles*warps) so the
table also shows warp instructions per item."""
import sys
rows = []
for l in open(sys.argv[1]).read().split("\n")[3:]:
    p = l.split()
    if len(p) < 6:
        continue
    f, ln = p[0].rsplit(":", 1)
    rows.append((f, int(ln), int(p[2]), float(p[4].rstrip("%")), float(p[5])))
units = float(sys.argv[2])
tot = sum(r[2] for r in rows)
print("total warp-inst %d = %.0f per unit" % (tot, tot / units))
seen = 0
for spec in sys.argv[3:]:
    f, rng, name = spec.split(":")
    a, b = map(int, rng.split("-"))
    sel = [r for r in rows if r[0] == f and a <= r[1] <= b]
    w = sum(r[2] for r in sel)
    seen += w
    t = sum(r[2] * r[4] / 32 for r in sel)
    s = sum(r[3] for r in sel)
    print("%-22s %6.2f%% %7.0f/unit  lanes %.2f  stall-samples %5.2f%%" % (name, 100.0 * w / tot, w / units, t / w if w else 0, s))
print("%-22s %6.2f%%" % ("(elsewhere)", 100.0 * (tot - seen) / tot))
