"""Where a sharded step spends its time (CUDA events between the enqueues): run under torchrun or alone."""
import os, sys, json
import torch
sys.path.insert(0, ".")
import libaec_b200 as L
from libaec_b200 import datagen
from libaec_b200.parallel import ShardedCodec, shard_range

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
p, _ = datagen.CONFIGS["c1"]
R = p.rsi * p.block_size
total = (256 << 20) // 4 * world
s, c = shard_range(total, R, rank, world)
raw = datagen.generate("c1", c, s)
d_raw = torch.from_numpy(raw).cuda()
d_back = torch.empty(raw.size + 16, dtype=torch.uint8, device="cuda")
sc = ShardedCodec(p, rank, world, local, stream=torch.cuda.current_stream().cuda_stream)
nrsi = (c + R - 1) // R
sc._ensure(raw.size, nrsi)
names = ["encode+summary", "all_gather", "plan", "repair", "place", "decode"]
acc = {n: 0.0 for n in names}
steps = 20
for it in range(steps + 3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    ev[0].record()
    sc.codec.encode_enqueue(p, d_raw, raw.size, sc.local, sc.offsets, d_grp=sc.grp); ev[1].record()
    if dist is not None:
        dist.all_gather_into_tensor(sc._all, sc._info_d); src = sc._all
    else:
        src = sc._info_d
    ev[2].record()
    sc.codec.shard_plan(src, world, rank, sc._plan_d); ev[3].record()
    sc.codec.encode_repair(p, d_raw, raw.size, sc.local); ev[4].record()
    sc.codec.place_planned(sc.local, sc.placed); ev[5].record()
    sc.bits = None
    sc.decode_enqueue(d_back, raw.size); ev[6].record()
    torch.cuda.synchronize()
    if it >= 3:
        for i, n in enumerate(names):
            acc[n] += ev[i].elapsed_time(ev[i + 1]) / steps
if rank == 0:
    print(json.dumps({"world": world, **{k: round(v, 4) for k, v in acc.items()}, "sum": round(sum(acc.values()), 4)}))
if dist is not None:
    dist.destroy_process_group()
