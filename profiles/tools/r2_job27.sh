cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/r2_final_pytest_2.txt; cat gpurun_out/r2_final_pytest_2.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_final_bench.json 2>/dev/null
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2_final_bench.json").read().strip().splitlines()[-1])
print("value", j["value"], "frac", j["roofline"]["frac"], "traffic", j["roofline"]["traffic"])
for k in ("e2e","e2e_indexed","e2e_pageable","pcie_copy_floor"): print(k, j[k]["value"], j[k]["ms_per_step"])
PY
bash profiles/tools/ncu_summarize.sh r2_final_c1 c1 256 > /dev/null 2>&1
timeout 100 python profiles/tools/time_scan.py c1 c2 c3:64 c4:256 c5_noise:128 c5_restricted 2>&1 | tail -1 | tee gpurun_out/r2_final_scan.txt
