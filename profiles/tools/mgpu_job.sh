# N-GPU check + bench: gpurun --gpus N -- 'N=2 bash profiles/tools/mgpu_job.sh'   (STRONG=1: also BASELINE config 5, 8 GiB in all)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${N:-2}
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_mgpu_pytest_$N.txt; cat gpurun_out/r2_mgpu_pytest_$N.txt; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > gpurun_out/r2_mgpu_check_$N.txt 2> gpurun_out/r2_mgpu_check_$N.err; tail -8 gpurun_out/r2_mgpu_check_$N.txt; tail -3 gpurun_out/r2_mgpu_check_$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/r2_mgpu_bench$N.err | tail -1 > gpurun_out/r2_mgpu_bench$N.json
python -c "
import json
j=json.loads(open('gpurun_out/r2_mgpu_bench$N.json').read()); print('n_gpus', j['n_gpus'], 'value', j['value'], 'ms_per_step', j['ms_per_step'], 'stitch_checked', j['stitch_checked'], 'e2e', j['e2e']['value'], 'pcie', j['pcie_copy_floor']['value'], j['detail']['numa'])"
tail -3 gpurun_out/r2_mgpu_bench$N.err
if [ "${STRONG:-0}" = "1" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 --strong --mib 8192 --device-only 2> gpurun_out/r2_mgpu_strong$N.err | tail -1 > gpurun_out/r2_mgpu_strong$N.json
python -c "
import json
j=json.loads(open('gpurun_out/r2_mgpu_strong$N.json').read()); print('STRONG n_gpus', j['n_gpus'], 'value', j['value'], 'ms_per_step', j['ms_per_step'], 'stitch_checked', j['stitch_checked'], j['scaling'])"
fi
