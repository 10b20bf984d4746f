for sk in -1 0; do
  echo "== AECB200_SCAN_SKIP8=$sk"
  AECB200_SCAN_SKIP8=$sk timeout 600 python profiles/tools/time_noindex.py c1 c4:512 c5_noise:256 c5_restricted 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l[:1] != 'c': print(l.rstrip()); continue
    n, _, j = l.partition(' '); j = json.loads(j)
    print(n, 'scan_ms %.2f' % j['scan_parallel_ms'], 'fast', j['scan_parallel_fast'], '/', j['nrsi'], 'buffer_decode_ms %.2f' % j['buffer_decode_noindex_ms'])
"
done
