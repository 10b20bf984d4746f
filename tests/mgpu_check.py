"""Run under torchrun with N >= 2 GPUs: the N-GPU stream must be byte-identical
to the 1-GPU stream (T3 of SURVEY.md section 4), for several configs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import libaec_b200 as L  # noqa: E402
from libaec_b200 import datagen  # noqa: E402
from libaec_b200.parallel import ShardedCodec, shard_range  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name, mib in (("c1", 64), ("c2", 32), ("c4", 24), ("c5_noise", 16), ("c5_restricted", 8)):
        p, _ = datagen.CONFIGS[name]
        B = p.bytes_per_sample
        R = p.rsi * p.block_size
        total = (mib << 20) // B * world - 13            # short last RSI on the last rank
        s, c = shard_range(total, R, rank, world)
        raw = datagen.generate(name, c, s)
        d_raw = torch.from_numpy(raw).cuda()
        sc = ShardedCodec(p, rank, world, local, stream=torch.cuda.current_stream().cuda_stream)
        plan = sc.encode(d_raw, raw.size)
        owned = sc.owned_bytes().clone()
        # the same protocol with nothing but enqueues (summary, all_gather, plan, repair, placement on the device)
        sc.placed.zero_()
        sc.step_enqueue(d_raw, raw.size)
        plan2 = sc.step_finish()
        same_plan = (plan2.bit_offset, plan2.k_in, plan2.end_bit, plan2.total_bits, plan2.head_or) == \
                    (plan.bit_offset, plan.k_in, plan.end_bit, plan.total_bits, plan.head_or)
        async_ok = same_plan and torch.equal(sc.owned_bytes(), owned)
        # gather every rank's bytes on rank 0
        sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([owned.numel()], dtype=torch.int64, device="cuda"))
        mx = int(max(int(x) for x in sizes))
        buf = torch.zeros(mx, dtype=torch.uint8, device="cuda")
        buf[: owned.numel()] = owned
        allb = [torch.zeros(mx, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(allb, buf)
        # decode needs no exchange
        d_back = torch.empty(raw.size + 16, dtype=torch.uint8, device="cuda")
        st, written = sc.decode(d_back, raw.size)
        good = st == 0 and written == raw.size and torch.equal(d_back[: raw.size], d_raw) and async_ok
        if not async_ok:
            print("rank %d %s: device-side protocol differs from the host protocol (plan %s vs %s)" % (rank, name, plan2, plan), flush=True)
        if rank == 0:
            stitched = torch.cat([allb[r][: int(sizes[r])] for r in range(world)]).cpu().numpy()
            whole = datagen.generate(name, total, 0)
            ref = L.DeviceCodec(device=local)
            d_whole = torch.from_numpy(whole).cuda()
            cap = (L.encode_bound(p, whole.size) + 64 + 3) // 4 * 4
            d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
            ref.encode_enqueue(p, d_whole, whole.size, d_out)
            st, bits, _ = ref.encode_finish()
            single = d_out[: (bits + 7) // 8].cpu().numpy()
            same = stitched.size == single.size and np.array_equal(stitched, single)
            print("%s: %d ranks, %d bytes, k_in per rank via plan, stitched == single-GPU stream: %s" %
                  (name, world, single.size, same), flush=True)
            good = good and same
            ref.close()
        t = torch.tensor([1 if good else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = ok and bool(t.item())
        sc.close()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
