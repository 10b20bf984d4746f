cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q 2>&1 | tail -5
AECB200_BENCH_NO_PIN=1 timeout 600 python bench.py --steps 10 --warmup 3 --device-only > gpurun_out/r2_f_bench_nopin.json 2> gpurun_out/r2_f.err; python -c "
import json
j = json.loads(open('gpurun_out/r2_f_bench_nopin.json').read().strip().splitlines()[-1])
print('NOPIN value', j['value'], 'enc ms', j['roofline']['encode']['ms'], 'dec ms', j['roofline']['decode']['ms'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_f_launches.csv python bench.py --device-only --steps 3 --warmup 3 > gpurun_out/r2_f_launch_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r2_f_launches.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
for r in rows[1:][-24:]:
    print(r[ki][:70], r[vi])
PY
