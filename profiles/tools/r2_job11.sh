cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_m_bench.json 2> gpurun_out/r2_m_bench.err; python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2_m_bench.json").read().strip().splitlines()[-1])
print("value", j["value"], "enc ms", j["roofline"]["encode"]["ms"], "dec ms", j["roofline"]["decode"]["ms"], "frac", j["roofline"]["frac"])
for k in ("e2e", "e2e_indexed", "e2e_pageable", "pcie_copy_floor"):
    print(k, j[k]["value"], j[k]["ms_per_step"])
PY
tail -3 gpurun_out/r2_m_bench.err
