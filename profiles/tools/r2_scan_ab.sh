timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q --tb=short 2>&1 | tail -8
for sp in 0 1; do
  echo "== AECB200_SCAN_SPARSE=$sp"
  AECB200_SCAN_SPARSE=$sp timeout 600 python profiles/tools/time_noindex.py c1 c2 c3:64 c4:1024 c5_noise:512 c5_restricted 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l[:1] != 'c': print(l.rstrip()); continue
    n, _, j = l.partition(' '); j = json.loads(j)
    print(n, 'scan_ms %.2f' % j['scan_parallel_ms'], 'fast', j['scan_parallel_fast'], '/', j['nrsi'], 'buffer_decode_ms %.2f' % j['buffer_decode_noindex_ms'])
"
  cp gpurun_out/r2_noindex.json gpurun_out/r2_noindex_sparse$sp.json
done
timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2_p_bench.json; python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_p_bench.json").read())
print("value", j["value"])
for k in ("e2e","e2e_indexed","e2e_pageable","pcie_copy_floor"): print(k, j[k]["value"], j[k]["ms_per_step"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:skim -c 200 --csv --log-file gpurun_out/r2_skim_sparse_launches.csv python profiles/tools/time_noindex.py c1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2_skim_sparse_launches.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]; 
    try: v = float(r[-1].replace(",", ""))
    except: continue
    agg.setdefault(name, []).append(v)
for k, v in agg.items(): print(k, len(v), "median us %.1f" % (sorted(v)[len(v)//2] / 1000.0))
PY
