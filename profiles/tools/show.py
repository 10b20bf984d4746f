#!/usr/bin/env python
"""Print the encode/decode times of the bench lines a gpurun job left in gpurun_out/ (TAG_bench*.json)."""
import json, sys, os
tag = sys.argv[1]
d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "gpurun_out")
print(open(os.path.join(d, tag + "_pytest.txt")).read()[-400:])
for f in (tag + "_bench.json", tag + "_bench_others.json"):
    try:
        lines = open(os.path.join(d, f)).read().splitlines()
    except OSError:
        continue
    for l in lines:
        if not l.startswith("{"):
            print(l[:300]); continue
        j = json.loads(l); r = j["roofline"]
        print("%-28s enc %.3f ms (%.1f%%)  dec %.3f ms (%.1f%%)  value %.0f  sm %s %s" % (
            j["config"]["workload"][:28], r["encode"]["ms"], 100 * r["encode"]["frac"], r["decode"]["ms"],
            100 * r["decode"]["frac"], j["value"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"]))
