cd $GRAFT_REPO_ROOT
timeout 120 python profiles/tools/time_scan.py c1 c2 c3:64 2>&1 | tail -1
for v in chunk16k chunk16k_nc8 chunk8k nc8; do
  AECB200_LIB=$GRAFT_REPO_ROOT/libaec_b200/lib/variants/$v/libaec.so.0 timeout 120 python profiles/tools/time_scan.py c1 c2 c3:64 2>&1 | tail -1
done
