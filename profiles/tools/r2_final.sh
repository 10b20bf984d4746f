# round-2 capture: tests twice, bench both arms, launch list, full ncu capture reduced on the box, other workloads, SZ, un-indexed timings
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
T=${TAG:-r2_final}
for i in 1 2; do timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${T}_pytest_$i.txt; tail -2 gpurun_out/${T}_pytest_$i.txt; done
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
j = json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/${T}_bench_reference.json").read().strip().splitlines()[-1])
print("value", j["value"], "enc ms", j["roofline"]["encode"]["ms"], "dec ms", j["roofline"]["decode"]["ms"], "frac", j["roofline"]["frac"], "traffic", j["roofline"]["traffic"])
for k in ("e2e", "e2e_indexed", "e2e_pageable", "pcie_copy_floor"):
    print(k, j[k]["value"], j[k]["ms_per_step"])
print("reference", r["value"], r["cpu_baseline"]["cores"], "ratio e2e", j["e2e"]["value"] / r["value"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --device-only --steps 3 --warmup 3 > gpurun_out/${T}_launch_bench.log 2>&1
bash profiles/tools/ncu_summarize.sh ${T}_c1 c1 256 > /dev/null 2>&1
rm -f gpurun_out/${T}_bench_others.json
for w in c2:256 c3:256 c4:1024 c5_noise:512 c5_restricted:256; do n=${w%%:*}; m=${w#*:}; timeout 400 python bench.py --device-only --steps 5 --warmup 3 --workload $n --mib $m 2>/dev/null | tail -1 >> gpurun_out/${T}_bench_others.json; done
timeout 500 python profiles/tools/sz_bench.py 64 2>&1 | tail -1 > gpurun_out/${T}_sz.txt; cp gpurun_out/r2_sz_bench.json gpurun_out/${T}_sz_bench.json
timeout 600 python profiles/tools/time_noindex.py c1:256 c2:256 c3:64 c4:256 c5_noise:128 2>&1 | tail -6 > gpurun_out/${T}_noindex.txt; cp gpurun_out/r2_noindex.json gpurun_out/${T}_noindex.json; cat gpurun_out/${T}_noindex.txt | cut -c1-260
