cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
bash profiles/tools/ncu_summarize.sh r2_final_c1 c1 256 > /dev/null 2>&1
ls -la gpurun_out | grep r2_final_c1 | head
