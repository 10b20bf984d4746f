cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python profiles/tools/dbg_enc.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_c_scan_pytest.txt; cat gpurun_out/r2_c_scan_pytest.txt
timeout 600 python profiles/tools/time_noindex.py c1:256 c2:128 c3:64 c4:128 c5_noise:64 2>&1 | tail -8
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_c_pytest.txt; cat gpurun_out/r2_c_pytest.txt
AECB200_SCAN_MODE=2 AECB200_SCAN_WINDOW_BITS=4096 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_c_pytest_forced.txt; cat gpurun_out/r2_c_pytest_forced.txt
