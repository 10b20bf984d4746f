"""Stage timing of ShardedCodec.encode under torchrun (diagnostic)."""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libaec_b200 import datagen
from libaec_b200.parallel import ShardedCodec, shard_range, plan_shards
from libaec_b200.api import Carry

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
p, _ = datagen.CONFIGS["c1"]
n = (256 << 20) // 4
raw = datagen.generate("c1", n, rank * n)
d_raw = torch.from_numpy(raw).cuda()
sc = ShardedCodec(p, rank, world, local, stream=torch.cuda.current_stream().cuda_stream)
for it in range(4):
    sc.encode(d_raw, raw.size)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = T()
    sc.codec.encode_enqueue(p, d_raw, raw.size, sc.local, sc.offsets)
    st, bits, kend = sc.codec.encode_finish()
    t1 = T()
    klo, khi, fc = sc.codec.shard_info()
    mine = torch.tensor([bits, klo, khi], dtype=torch.int64, device="cuda")
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    infos = [tuple(int(x) for x in v.tolist()) for v in allv]
    t2 = T()
    plan = plan_shards(infos)[rank]
    if plan.k_in != 0:
        sc.codec.set_tile_limit(fc + 1)
        sc.codec.encode_enqueue(p, d_raw, raw.size, sc.local, None, Carry(0, plan.k_in, 0))
    t3 = T()
    sc.codec.place_bits(sc.local, bits, sc.placed, plan.bit_offset & 31)
    t4 = T()
    w = sc.placed.view(torch.int32)
    nwords = (((plan.bit_offset & 31) + bits + 31) >> 5)
    edge = torch.stack([w[0], w[max(nwords - 1, 0)]]).to(torch.int64)
    alle = [torch.zeros_like(edge) for _ in range(world)]
    dist.all_gather(alle, edge)
    if rank > 0 and (plan.bit_offset & 31):
        w[0] = w[0] | alle[rank - 1][1].to(torch.int32)
    t5 = T()
    if rank == 0 or it == 2:
        print("rank %d it %d: encode %.3f ms, gather1 %.3f, repair %.3f (k_in %d, fc %d), place %.3f, gather2 %.3f" %
              (rank, it, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, plan.k_in, fc, (t4-t3)*1e3, (t5-t4)*1e3), flush=True)
dist.destroy_process_group()
