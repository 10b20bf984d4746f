cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_final_bench.json 2>/dev/null
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2_final_bench.json").read().strip().splitlines()[-1])
print("value", j["value"], "frac", j["roofline"]["frac"], "traffic", j["roofline"]["traffic"], "e2e", j["e2e"]["value"], j["e2e"]["ms_per_step"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:skim -c 200 --csv --log-file gpurun_out/r2_final_skim_launches.csv python profiles/tools/time_noindex.py c1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2_final_skim_launches.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]
    try: v = float(r[-1].replace(",", ""))
    except: continue
    agg.setdefault(name, []).append(v)
for k, v in agg.items(): print(k, len(v), "median us %.1f" % (sorted(v)[len(v)//2] / 1000.0))
PY
