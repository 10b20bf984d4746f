"""The multi-GPU shard protocol's device side on ONE GPU: R logical ranks code their RSI ranges one after
the other, their 32-byte summaries are put side by side the way the all_gather would, and the plan,
repair and placement kernels (no host arithmetic in between) write every rank's words straight into
one buffer for the whole stream.  That buffer must equal the single-coder stream byte for byte
(T3 of SURVEY section 4; tests/mgpu_check.py runs the same with real ranks and NCCL)."""
import numpy as np
import pytest

import libaec_b200 as L
from libaec_b200 import datagen
from libaec_b200.parallel import ShardedCodec, shard_range
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,mib,world", [("c1", 24, 3), ("c1", 8, 8), ("c2", 16, 4), ("c4", 12, 2),
                                            ("c5_noise", 8, 3), ("c5_restricted", 4, 5)])
def test_logical_ranks_stitch_into_the_single_coder_stream(name, mib, world):
    import torch
    p, _ = datagen.CONFIGS[name]
    B = p.bytes_per_sample
    R = p.rsi * p.block_size
    total = (mib << 20) // B - 11                         # short last RSI on the last rank
    whole = datagen.generate(name, total)
    op = po.Params(p.bits_per_sample, p.block_size, p.rsi, p.flags)
    want = (po.ref_encode if po.ref_available() else po.orc_encode)(op, whole)["out"]
    cap = (L.encode_bound(p, whole.size) + 64 + 3) // 4 * 4
    d_stream = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    ranks = []
    for r in range(world):
        s, c = shard_range(total, R, r, world)
        d_raw = torch.from_numpy(whole[s * B:(s + c) * B].copy()).cuda()
        sc = ShardedCodec(p, r, world, 0, stream=torch.cuda.current_stream().cuda_stream)
        sc._ensure(c * B, (c + R - 1) // R)
        sc.codec.encode_enqueue(p, d_raw, c * B, sc.local, sc.offsets, d_grp=sc.grp)   # writes sc._info_d
        ranks.append((sc, d_raw, c * B))
    torch.cuda.synchronize()
    d_all = torch.cat([sc._info_d for sc, _, _ in ranks])           # what all_gather_into_tensor delivers
    k_ins = []
    for r, (sc, d_raw, nb) in enumerate(ranks):
        sc.codec.shard_plan(d_all, world, r, sc._plan_d)
        sc.codec.encode_repair(p, d_raw, nb, sc.local)
        sc.codec.place_planned(sc.local, None, dst_ptr=d_stream.data_ptr(), dst_cap=cap, global_stream=True,
                               last_rank=r == world - 1)
        torch.cuda.synchronize()
        k_ins.append(int(sc._plan_d[0].item()))
    total_bits = int(ranks[0][0]._plan_d[4].item())
    assert (total_bits + 7) // 8 == want.size, (total_bits, want.size)
    got = d_stream[: want.size].cpu().numpy()
    if not np.array_equal(got, want):
        d = np.nonzero(got != want)[0]
        raise AssertionError("%s x%d: %d bytes differ, first at %s, incoming k per rank %s" % (name, world, d.size, d[:6], k_ins))
    for sc, _, _ in ranks:
        sc.close()


def test_device_side_protocol_equals_host_protocol_single_rank():
    import torch
    p, _ = datagen.CONFIGS["c1"]
    raw = datagen.generate("c1", (8 << 20) // 4 - 3)
    d_raw = torch.from_numpy(raw).cuda()
    sc = ShardedCodec(p, 0, 1, 0, stream=torch.cuda.current_stream().cuda_stream)
    plan = sc.encode(d_raw, raw.size)
    owned = sc.owned_bytes().clone()
    sc.placed.zero_()
    sc.step_enqueue(d_raw, raw.size)
    d_back = torch.zeros(raw.size + 16, dtype=torch.uint8, device="cuda")
    sc.decode_enqueue(d_back, raw.size)
    plan2 = sc.step_finish()
    assert (plan2.bit_offset, plan2.k_in, plan2.end_bit, plan2.total_bits) == (plan.bit_offset, plan.k_in, plan.end_bit, plan.total_bits)
    assert torch.equal(sc.owned_bytes(), owned)
    st, written = sc.codec.decode_finish()
    assert st == 0 and written == raw.size and torch.equal(d_back[: raw.size], d_raw)
    sc.close()
