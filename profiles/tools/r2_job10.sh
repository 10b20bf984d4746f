cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_l_pytest.txt; cat gpurun_out/r2_l_pytest.txt
AECB200_SCAN_MODE=2 AECB200_SCAN_WINDOW_BITS=4096 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scan.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_l_pytest_forced.txt; cat gpurun_out/r2_l_pytest_forced.txt
timeout 300 python profiles/tools/time_noindex.py c1:256 2>&1 | tail -1 | cut -c1-300
