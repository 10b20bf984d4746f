# scratch job runner for gpurun (edited per experiment): TAG=x [NCU=1] [TESTS=0] [FULL=1] bash profiles/tools/gpu_job.sh
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
T=${TAG:-x}
if [ "${TESTS:-1}" = "1" ]; then timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} 2>&1 | tail -25 > gpurun_out/${T}_pytest.txt; fi
if [ "${FULL:-0}" = "1" ]; then
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
else
timeout 300 python bench.py --device-only --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
fi
if [ "${OTHERS:-0}" = "1" ]; then for w in c2 c3 c4 c5_noise c5_restricted; do timeout 200 python bench.py --device-only --steps 5 --warmup 3 --workload $w --mib 128 2>&1 | tail -1 >> gpurun_out/${T}_bench_others.json; done; fi
if [ "${NCU:-0}" = "1" ]; then timeout 600 ncu --set full --clock-control none --import-source on -k regex:"aec_(encode|decode_warp)_kernel" -s 6 -c 2 -f -o gpurun_out/prof_${T} python bench.py --device-only --steps 1 --warmup 3 > gpurun_out/${T}_ncu.log 2>&1; fi
tail -12 gpurun_out/${T}_pytest.txt; tail -3 gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']; print('enc %.4f ms dec %.4f ms value %.0f e2e %.1f'%(r['encode']['ms'], r['decode']['ms'], j['value'], j['e2e']['value']))
"
