cd $GRAFT_REPO_ROOT
timeout 120 python profiles/tools/time_scan.py c1 c4:256 c2 2>&1 | tail -1
for v in chunk16k chunk64k tile4k tile16k dp8 nc8 nc2; do
  AECB200_LIB=$GRAFT_REPO_ROOT/libaec_b200/lib/variants/$v/libaec.so.0 timeout 120 python profiles/tools/time_scan.py c1 c4:256 c2 2>&1 | tail -1
done
for w in 16777216 67108864; do AECB200_SCAN_WINDOW_BITS=$w timeout 120 python profiles/tools/time_scan.py c1 c4:256 c2 2>&1 | tail -1; done
timeout 120 python profiles/tools/time_scan.py c1 c4:256 c2 2>&1 | tail -1
