"""Multi-GPU shard stitch on CPU: world_size-2 gloo processes run the host-side
protocol of libaec_b200/parallel.py (exclusive scan of shard bit lengths, k
clamp chain, placement at the global bit phase, boundary-word merge) with the
CPU model harness standing in for the CUDA kernels, and the stitched stream
must be byte-identical to the oracle's stream of the whole input (T3 of
SURVEY.md section 4)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cases import pack_samples, synth_values
from libaec_b200.parallel import place_bits_host, plan_shards, shard_range, tail64_host
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model():
    path = os.path.join(ROOT, "tests", "_build", "libaec_cpumodel.so")
    if not os.path.exists(path):
        from libaec_b200.build import build
        build()
    return C.CDLL(path)


def _model_encode(m, p, raw, seed_k):
    src = np.ascontiguousarray(raw)
    cap = (po.worst_case_bytes(p, src.size) + 64 + 3) // 4 * 4
    out = np.zeros(cap, np.uint8)
    ol, eb, ek = C.c_size_t(0), C.c_uint64(0), C.c_uint32(0)
    rc = m.model_encode(C.c_uint32(p.bits_per_sample), C.c_uint32(p.block_size), C.c_uint32(p.rsi),
                        C.c_uint32(p.flags), C.c_int(0), src.ctypes.data_as(C.c_void_p), C.c_size_t(src.size),
                        out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(ol), None,
                        C.c_uint64(0), C.c_uint32(seed_k), C.c_uint32(0), C.byref(eb), C.byref(ek))
    assert rc == 0
    return out[:ol.value].copy(), eb.value, ek.value


CASES = [
    (po.Params(32, 16, 16, po.AEC_DATA_SIGNED | po.AEC_DATA_PREPROCESS), 3, 37),
    (po.Params(16, 32, 8, po.AEC_DATA_PREPROCESS | po.AEC_DATA_MSB), 5, 11),
    (po.Params(8, 8, 64, po.AEC_DATA_PREPROCESS), 1, 9),
    (po.Params(12, 16, 4, 0), 2, 23),
    (po.Params(24, 64, 3, po.AEC_DATA_3BYTE | po.AEC_DATA_PREPROCESS), 4, 6),
]


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _model()
    ok = True
    for ci, (p, kind, nrsi_total) in enumerate(CASES):
        rng = np.random.default_rng(100 + ci)
        R = p.rsi * p.block_size
        total = nrsi_total * R - 5                      # short last RSI
        vals = synth_values(rng, p.bits_per_sample, total, kind, bool(p.flags & po.AEC_DATA_SIGNED))
        raw = pack_samples(vals, p)
        B = p.bytes_per_sample
        s, c = shard_range(total, R, rank, world)
        shard = raw[s * B:(s + c) * B]
        kmax = {5: 29, 4: 13, 3: 5, 2: 1, 1: 0}[p.id_len]
        # 1. independent shard encode (k seed 0) + the shard's clamp pair + its last 64 bits
        s0, bits, klo = _model_encode(m, p, shard, 0)
        _, _, khi = _model_encode(m, p, shard, kmax)
        t64 = tail64_host(s0, bits)
        # 2. the only exchange
        mine = torch.tensor([bits, klo, khi, t64 - (1 << 64) if t64 >= (1 << 63) else t64], dtype=torch.int64)
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        infos = [(v[0], v[1], v[2], v[3] & 0xFFFFFFFFFFFFFFFF) for v in (t.tolist() for t in allv)]
        plan = plan_shards(infos)[rank]
        # 3. re-code with the true incoming k (the tail never changes: lengths do not depend on k
        #    and only the leading plateau blocks can change their bits)
        stream, bits2, _ = _model_encode(m, p, shard, plan.k_in)
        assert bits2 == bits
        # 4. place at the global phase, first word completed with the predecessor's tail
        placed = place_bits_host(stream, bits, plan.bit_offset, plan.head_or)
        words = placed.view(">u4").astype(np.int64)
        owned = words[: plan.word_hi - plan.word_lo].astype(">u4").view(np.uint8)
        if rank == world - 1:
            owned = owned[: (plan.total_bits + 7) // 8 - plan.word_lo * 4]
        # gather the segments on rank 0 and compare with the whole-stream oracle
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([owned.size], dtype=torch.int64))
        mx = int(max(int(x) for x in sizes))
        buf = torch.zeros(mx, dtype=torch.uint8)
        buf[: owned.size] = torch.from_numpy(owned.copy())
        allb = [torch.zeros(mx, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(allb, buf)
        if rank == 0:
            got = np.concatenate([allb[r][: int(sizes[r])].numpy() for r in range(world)])
            want = po.orc_encode(p, raw)["out"]
            ok = ok and np.array_equal(got, want)
    results[rank] = ok
    dist.destroy_process_group()


def test_two_rank_stitch_equals_single_stream():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, port, results), nprocs=2, join=True)
    assert results[0] is True and results[1] is True


def test_plan_shards_chain():
    plans = plan_shards([(100, 3, 5, 0xFFFFFFFFFFFFFFFF), (64, 0, 29, 0), (7, 9, 9, 0), (50, 2, 4, 0)])
    assert plans[1].head_or == 0xF0000000          # 100 % 32 = 4 tail bits of shard 0
    assert [p.bit_offset for p in plans] == [0, 100, 164, 171]
    assert [p.k_in for p in plans] == [0, 3, 3, 9]
    assert plans[-1].end_bit == plans[-1].total_bits == 221
    assert plans[1].word_lo == 3 and plans[1].word_hi == 5
