cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=aec_ --print-limit 20 python profiles/tools/sanitize.py > gpurun_out/r2_sanitize_$tool.txt 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload ok|Error|hazard" gpurun_out/r2_sanitize_$tool.txt | sort | uniq -c | head -8
done
