cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_e_scan_pytest.txt; cat gpurun_out/r2_e_scan_pytest.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_e_pytest.txt; cat gpurun_out/r2_e_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_e_bench.json 2> gpurun_out/r2_e_bench.err; python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2_e_bench.json").read().strip().splitlines()[-1])
print("value", j["value"], "enc ms", j["roofline"]["encode"]["ms"], "dec ms", j["roofline"]["decode"]["ms"])
for k in ("e2e", "e2e_indexed", "e2e_pageable", "pcie_copy_floor"):
    print(k, j[k]["value"], j[k]["ms_per_step"])
PY
tail -3 gpurun_out/r2_e_bench.err
AECB200_NO_COOP=1 timeout 600 python bench.py --steps 10 --warmup 3 --device-only > gpurun_out/r2_e_bench_nocoop.json 2>> gpurun_out/r2_e_bench.err; python -c "
import json
j = json.loads(open('gpurun_out/r2_e_bench_nocoop.json').read().strip().splitlines()[-1])
print('NOCOOP value', j['value'], 'enc ms', j['roofline']['encode']['ms'], 'dec ms', j['roofline']['decode']['ms'])"
for w in 22 23 24 25 26; do AECB200_SCAN_WINDOW_BITS=$((1<<w)) timeout 300 python profiles/tools/time_noindex.py c1:256 2>&1 | tail -1 | cut -c1-400; done
