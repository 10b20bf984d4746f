timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q --tb=short 2>&1 | tail -4
for ts in 0 1; do
  echo "== AECB200_SCAN_TWO_STREAMS=$ts"
  AECB200_SCAN_TWO_STREAMS=$ts timeout 600 python profiles/tools/time_noindex.py c1 c2 c3:64 c4:512 c5_noise:256 c5_restricted 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l[:1] != 'c': print(l.rstrip()); continue
    n, _, j = l.partition(' '); j = json.loads(j)
    print(n, 'scan_ms %.2f' % j['scan_parallel_ms'], 'fast', j['scan_parallel_fast'], '/', j['nrsi'], 'buffer_decode_ms %.2f' % j['buffer_decode_noindex_ms'])
"
done
timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2_q_bench.json; python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_q_bench.json").read())
print("value", j["value"])
for k in ("e2e","e2e_indexed","e2e_pageable","pcie_copy_floor"): print(k, j[k]["value"], j[k]["ms_per_step"])
PY
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
