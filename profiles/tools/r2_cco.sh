# the reference's unchanged check_code_options binary against our libaec (time-boxed)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export LD_LIBRARY_PATH=$GRAFT_REPO_ROOT/libaec_b200/lib:$LD_LIBRARY_PATH
ldd oracle/_ref/check_code_options | grep -E 'libaec' > gpurun_out/r2_cco_ldd.txt
START=$(date +%s.%N)
timeout ${1:-360} stdbuf -oL oracle/_ref/check_code_options > gpurun_out/r2_check_code_options.log 2>&1
RC=$?
END=$(date +%s.%N)
echo "exit $RC after $(echo "$END - $START" | bc -l 2>/dev/null || python -c "print($END - $START)") s" > gpurun_out/r2_check_code_options.time
echo "PASS lines: $(grep -c PASS gpurun_out/r2_check_code_options.log) of 840; FAIL lines: $(grep -c FAIL gpurun_out/r2_check_code_options.log)" >> gpurun_out/r2_check_code_options.time
cat gpurun_out/r2_check_code_options.time; cat gpurun_out/r2_cco_ldd.txt; tail -3 gpurun_out/r2_check_code_options.log
