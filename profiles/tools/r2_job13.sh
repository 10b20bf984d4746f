timeout 600 python profiles/tools/dbg_fuzz.py > gpurun_out/r2_dbg_fuzz.txt 2>&1; grep -c "<<<" gpurun_out/r2_dbg_fuzz.txt; grep "<<<" gpurun_out/r2_dbg_fuzz.txt | head -20
for al in 0 1; do
  AECB200_SCANNER_ALONE=$al timeout 600 python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2_alone$al.json
  python - <<PY
import json
j=json.loads(open("gpurun_out/r2_alone$al.json").read())
print("alone=$al value", j["value"], "ms", j["ms_per_step"], "frac", j["roofline"]["frac"], {k:v for k,v in j.items() if "enc" in k or "dec" in k})
PY
done
AECB200_SCANNER_ALONE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
