cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/f_launches.csv python bench.py --device-only --steps 3 --warmup 3 > gpurun_out/f_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"aec_(encode|decode_warp)_kernel" -s 6 -c 2 -f -o gpurun_out/prof_f python bench.py --device-only --steps 1 --warmup 3 > gpurun_out/f_ncu.log 2>&1
tail -1 gpurun_out/f_bench.json; tail -1 gpurun_out/f_bench_reference.json
