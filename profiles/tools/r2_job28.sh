cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_final_bench.json 2>/dev/null
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2_final_bench.json").read().strip().splitlines()[-1])
print("value", j["value"], "frac", j["roofline"]["frac"], "traffic", j["roofline"]["traffic"], "e2e", j["e2e"]["value"], j["e2e"]["ms_per_step"], "pageable", j["e2e_pageable"]["value"])
PY
