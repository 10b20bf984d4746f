# 2-GPU check + bench: gpurun --gpus 2 -- 'bash profiles/tools/mgpu_job.sh'
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${N:-2}
timeout 500 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --device-only 2> gpurun_out/mgpu_bench$N.err | tail -1 > gpurun_out/mgpu_bench$N.json
python -c "
import json
j=json.loads(open('gpurun_out/mgpu_bench$N.json').read()); print('n_gpus', j['n_gpus'], 'value', j['value'], 'ms_per_step', j['ms_per_step'])"
tail -3 gpurun_out/mgpu_bench$N.err
