timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q --tb=short 2>&1 | tail -12; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2; timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2_o_bench.json; python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_o_bench.json").read())
print("value", j["value"], "traffic", j["roofline"]["traffic"])
for k in ("e2e","e2e_indexed","e2e_pageable","pcie_copy_floor"): print(k, j[k]["value"], j[k]["ms_per_step"])
PY
