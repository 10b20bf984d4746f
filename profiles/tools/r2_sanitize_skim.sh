cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for sk in -1 1; do
  AECB200_SCAN_SKIP8=$sk timeout 900 compute-sanitizer --tool memcheck --kernel-regex kns=aec_skim --print-limit 20 python profiles/tools/sanitize.py > gpurun_out/r2_sanitize_skim_memcheck_skip$sk.txt 2>&1
  echo "== memcheck skim kernels, AECB200_SCAN_SKIP8=$sk"; grep -E "ERROR SUMMARY|sanitize workload ok|Error|Invalid" gpurun_out/r2_sanitize_skim_memcheck_skip$sk.txt | sort | uniq -c | head -8
done
AECB200_SCAN_SKIP8=1 timeout 900 compute-sanitizer --tool racecheck --kernel-regex kns=aec_skim --print-limit 20 python profiles/tools/sanitize.py > gpurun_out/r2_sanitize_skim_racecheck.txt 2>&1
echo "== racecheck skim kernels"; grep -E "RACECHECK SUMMARY|sanitize workload ok|hazard" gpurun_out/r2_sanitize_skim_racecheck.txt | sort | uniq -c | head -8
