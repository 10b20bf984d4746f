cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
rm -f gpurun_out/r2_i_bench_others.json
for w in c2:256 c3:256 c4:1024 c5_noise:512 c5_restricted:256; do n=${w%%:*}; m=${w#*:}; timeout 400 python bench.py --device-only --steps 5 --warmup 3 --workload $n --mib $m 2>/dev/null | tail -1 >> gpurun_out/r2_i_bench_others.json; done
python - <<'PY'
import json
for l in open("gpurun_out/r2_i_bench_others.json"):
    if l.startswith("{"):
        j = json.loads(l); r = j["roofline"]
        print(j["config"]["workload"][:14], "enc %.4f ms (%.3f) dec %.4f ms (%.3f) ratio %.2f" % (r["encode"]["ms"], r["encode"]["frac"], r["decode"]["ms"], r["decode"]["frac"], j["detail"]["ratio"]))
PY
for w in c2 c5_noise c3; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"aec_(encode|decode_warp)_kernel" -s 6 -c 2 -f -o gpurun_out/prof_r2_$w python bench.py --device-only --steps 1 --warmup 3 --workload $w --mib 128 > gpurun_out/r2_i_ncu_$w.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
