cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"skim|scan|group|decode" -c 300 --csv --log-file gpurun_out/r2_h_skim_launches.csv python profiles/tools/time_noindex.py c1:256 > gpurun_out/r2_h_skim_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r2_h_skim_launches.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except: continue
    k = r[ki][:60]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in agg.items(): print("%-62s n=%4d total=%10.1f us  avg=%8.1f us" % (k, n, t / 1e3, t / 1e3 / n))
PY
