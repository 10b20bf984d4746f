# round 2, job 1: reference check_code_options against our library (time-boxed), baseline bench of the round-1 build
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export LD_LIBRARY_PATH=$GRAFT_REPO_ROOT/libaec_b200/lib:$LD_LIBRARY_PATH
ldd oracle/_ref/check_code_options | grep -E 'libaec' > gpurun_out/r2_cco_ldd.txt
( time timeout 420 oracle/_ref/check_code_options ) > gpurun_out/r2_check_code_options.log 2> gpurun_out/r2_check_code_options.time
echo "exit $?" >> gpurun_out/r2_check_code_options.time
grep -c PASS gpurun_out/r2_check_code_options.log; tail -2 gpurun_out/r2_check_code_options.log; cat gpurun_out/r2_check_code_options.time
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_a_bench.json 2> gpurun_out/r2_a_bench.err; tail -c 600 gpurun_out/r2_a_bench.json
