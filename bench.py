#!/usr/bin/env python
"""bench.py -- AEC encode & decode throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1] [--mib 256]
    python bench.py --impl reference ...        # the reference's CPU implementation, all host cores

A step = one pass of the hot path over one batch of synthetic input: encode the
batch, decode it back (both directions of the metric).  `value` is raw
(uncompressed) GB/s with buffers resident in HBM; `e2e` is the same work
through the libaec-facing host-pointer C ABI (pinned host buffers, H2D/D2H
inside the timed region).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aec_encode_decode_throughput_raw"
UNIT = "GB/s"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def shard_samples(total_samples: int, rsi_samples: int, rank: int, world: int):
    """Contiguous RSI-aligned shard [start, start+count) of rank."""
    nrsi = (total_samples + rsi_samples - 1) // rsi_samples
    per = (nrsi + world - 1) // world
    s = min(rank * per, nrsi) * rsi_samples
    e = min(min((rank + 1) * per, nrsi) * rsi_samples, total_samples)
    return s, max(e - s, 0)


# ----------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# ----------------------------------------------------------------------------

def cpu_reference_run(name: str, nbytes: int, threads: int, steps: int, warmup: int):
    """Times oracle/_ref (the unmodified reference compiled by oracle/Makefile),
    or the oracle port when that build is absent, one shard per thread."""
    from concurrent.futures import ThreadPoolExecutor
    from libaec_b200 import datagen
    from oracle import pyoracle as po
    p, _ = datagen.CONFIGS[name]
    op = po.Params(p.bits_per_sample, p.block_size, p.rsi, p.flags)
    B = p.bytes_per_sample
    total = nbytes // B
    kind = "reference" if po.ref_available() else "port"
    enc = po.ref_encode if kind == "reference" else po.orc_encode
    dec = po.ref_decode if kind == "reference" else po.orc_decode
    shards = []
    for t in range(threads):
        s, c = shard_samples(total, p.rsi * p.block_size, t, threads)
        if c:
            shards.append(datagen.generate(name, c, s))

    def work(raw):
        t0 = time.perf_counter()
        e = enc(op, raw)
        t1 = time.perf_counter()
        d = dec(op, e["out"], raw.size)
        t2 = time.perf_counter()
        assert e["status"] == 0 and d["status"] == 0 and np.array_equal(d["out"][:64], raw[:64])
        return t1 - t0, t2 - t1, e["out"].size

    times = []
    with ThreadPoolExecutor(max_workers=len(shards)) as ex:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            res = list(ex.map(work, shards))
            wall = time.perf_counter() - t0
            if it >= warmup:
                times.append((wall, max(r[0] for r in res), max(r[1] for r in res)))
    raw_bytes = sum(s.size for s in shards)
    wall = float(np.mean([t[0] for t in times]))
    return {"kind": kind, "cores": len(shards), "raw_bytes": raw_bytes, "ms_per_step": wall * 1e3,
            "value": 2 * raw_bytes / wall / 1e9,
            "encode_gbs": raw_bytes / float(np.mean([t[1] for t in times])) / 1e9,
            "decode_gbs": raw_bytes / float(np.mean([t[2] for t in times])) / 1e9}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    nbytes = min(args.mib, 256) << 20
    r = cpu_reference_run(args.workload, nbytes, threads, max(1, min(args.steps, 3)), min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": max(1, min(args.steps, 3)), "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args), "bytes_per_step": r["raw_bytes"]},
        "encode_gbs": r["encode_gbs"], "decode_gbs": r["decode_gbs"],
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": "%d MiB of the workload, one RSI-aligned shard per host thread, aec_buffer_encode + aec_buffer_decode" % (r["raw_bytes"] >> 20)},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_name(args):
    from libaec_b200 import datagen
    p, desc = datagen.CONFIGS[args.workload]
    return "%s: %s; %d MiB per GPU (n=%d J=%d rsi=%d flags=%d)" % (
        args.workload, desc, args.mib, p.bits_per_sample, p.block_size, p.rsi, p.flags)


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------

def run_ours(args):
    import torch
    import libaec_b200 as L
    from libaec_b200 import datagen

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    p, _ = datagen.CONFIGS[args.workload]
    B = p.bytes_per_sample
    R = p.rsi * p.block_size
    per_gpu = (args.mib << 20) // B
    total = per_gpu * world                      # weak scaling: fixed work per GPU
    start, count = shard_samples(total, R, rank, world)
    raw = datagen.generate(args.workload, count, start)
    nrsi = (count + R - 1) // R

    stream = torch.cuda.current_stream()
    codec = L.DeviceCodec(device=local, stream=stream.cuda_stream)
    d_raw = torch.from_numpy(raw).cuda()
    cap = (L.encode_bound(p, raw.size) + 64 + 3) // 4 * 4
    d_comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_offs = torch.empty(nrsi, dtype=torch.int64, device="cuda")
    d_back = torch.empty(raw.size + 16, dtype=torch.uint8, device="cuda")
    d_grp = torch.zeros(max(codec.group_index_entries(p, raw.size), 1), dtype=torch.int64, device="cuda")

    # one checked pass: byte-exact round trip on this rank's shard
    codec.encode_enqueue(p, d_raw, raw.size, d_comp, d_offs, d_grp=d_grp)
    st, bits, kend = codec.encode_finish()
    assert st == 0
    comp_bytes = (bits + 7) // 8
    codec.decode_enqueue(p, d_comp, comp_bytes, d_offs, nrsi, d_back, raw.size, d_grp=d_grp)
    st, written = codec.decode_finish()
    assert st == 0 and written == raw.size
    assert torch.equal(d_back[:raw.size], d_raw), "round trip differs"
    handover = codec.last_handover

    # multi-GPU: shards are independent RSI ranges; the only exchange is the
    # tiny all-gather of per-shard (bits, k) that places each shard in the
    # single stream (see DESIGN.md "multi-GPU")
    shard_bits = [bits]
    if dist is not None:
        t = torch.tensor([bits, kend], dtype=torch.int64, device="cuda")
        allb = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allb, t)
        shard_bits = [int(x[0].item()) for x in allb]

    sharded = None
    if world > 1:
        # N > 1: the full shard protocol (independent encode, 24-byte all_gather, k repair,
        # placement at the global bit phase, boundary-word exchange) is inside every step
        from libaec_b200.parallel import ShardedCodec
        sharded = ShardedCodec(p, rank, world, local, stream=stream.cuda_stream)

    def step():
        codec.encode_enqueue(p, d_raw, raw.size, d_comp, d_offs, d_grp=d_grp)
        codec.decode_enqueue(p, d_comp, comp_bytes, d_offs, nrsi, d_back, raw.size, d_grp=d_grp)

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps + 1)]
    for _ in range(args.warmup):
        step()
        if sharded is not None:
            sharded.encode(d_raw, raw.size)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = codec.launches
    sh_l0 = sharded.codec.launches if sharded is not None else 0
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    enc_ms, dec_ms = [], []
    t_start.record()
    marks = []
    for i in range(args.steps):
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        if sharded is not None:
            sharded.encode_local(d_raw, raw.size)       # independent shard encode
            b.record()
            sharded.decode_enqueue(d_back, raw.size)    # decode needs no exchange
            sharded.exchange_begin()                    # 32-byte all_gather on a side stream, next to the decode
            sharded.stitch()                            # gathered offsets -> k repair, placement
        else:
            codec.encode_enqueue(p, d_raw, raw.size, d_comp, d_offs, d_grp=d_grp)
            b.record()
            codec.decode_enqueue(p, d_comp, comp_bytes, d_offs, nrsi, d_back, raw.size, d_grp=d_grp)
        c.record()
        marks.append((a, b, c))
    t_end.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    launches = codec.launches - launches0 + (sharded.codec.launches - sh_l0 if sharded is not None else 0)
    # keep the sampler alive long enough to see the load
    if rank == 0:
        t_busy = time.time()
        while time.time() - t_busy < 1.0:
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = t_start.elapsed_time(t_end)
    for a, b, c in marks:
        enc_ms.append(a.elapsed_time(b)); dec_ms.append(b.elapsed_time(c))
    if dist is not None:
        t = torch.tensor([elapsed_ms, float(np.mean(enc_ms)), float(np.mean(dec_ms))], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, enc_mean, dec_mean = (float(x) for x in t.tolist())
        tot = torch.tensor([raw.size, comp_bytes], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        raw_total, comp_total = (int(x) for x in tot.tolist())
    else:
        enc_mean, dec_mean = float(np.mean(enc_ms)), float(np.mean(dec_ms))
        raw_total, comp_total = raw.size, comp_bytes
    ms_per_step = elapsed_ms / args.steps
    value = 2 * raw_total / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: libaec-facing host-pointer ABI, pinned host buffers ----
    h_raw = torch.from_numpy(raw).pin_memory()
    h_comp = torch.empty(cap, dtype=torch.uint8).pin_memory()
    h_back = torch.empty(raw.size + 16, dtype=torch.uint8).pin_memory()
    h_offs = torch.empty(nrsi, dtype=torch.int64).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_step():
        st, n, noff = codec.encode_host(p, h_raw.data_ptr(), raw.size, h_comp.data_ptr(), cap,
                                        h_offs.data_ptr(), nrsi)
        assert st == 0 and n == comp_bytes
        st, m = codec.decode_host(p, h_comp.data_ptr(), n, h_back.data_ptr(), raw.size,
                                  h_offs.data_ptr(), noff)
        assert st == 0 and m == raw.size

    if args.device_only:
        e2e_steps = 0
        e2e_s = float("inf")
    else:
        e2e_step()
        assert np.array_equal(h_back[:raw.size].numpy(), raw), "e2e round trip differs"
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    if e2e_steps:
        e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = 2 * raw_total / e2e_s / 1e9

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    enc_gbs = raw.size / (enc_mean * 1e-3) / 1e9
    dec_gbs = raw.size / (dec_mean * 1e-3) / 1e9
    algo_bytes = raw.size + comp_bytes            # per launch: raw in + compressed out (encode), reverse for decode
    dominant = "aec_encode_kernel" if enc_mean >= dec_mean else "aec_decode_warp_kernel"
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of that kernel from the committed ncu capture
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            tj = json.load(f)[dominant]
        if args.workload == "c1" and args.mib == 256:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except Exception:
        traffic = None
    dom_ms = max(enc_mean, dec_mean)
    achieved = algo_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes,
                "encode": {"ms": enc_mean, "achieved": algo_bytes / (enc_mean * 1e-3) / 1e9,
                           "frac": algo_bytes / (enc_mean * 1e-3) / 1e9 / peak},
                "decode": {"ms": dec_mean, "achieved": algo_bytes / (dec_mean * 1e-3) / 1e9,
                           "frac": algo_bytes / (dec_mean * 1e-3) / 1e9 / peak}}

    cpu = None
    if not args.device_only:
        r = cpu_reference_run(args.workload, min(args.mib, 128) << 20, 1, 1, 0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": 1, "kind": r["kind"],
               "encode_gbs": r["encode_gbs"], "decode_gbs": r["decode_gbs"],
               "sample": "%d MiB of the workload, single thread, aec_buffer_encode + aec_buffer_decode" % (r["raw_bytes"] >> 20)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args), "raw_bytes_per_gpu": raw.size,
                   "compressed_bytes_per_gpu": comp_bytes, "ratio": raw.size / comp_bytes,
                   "l2": "inputs larger than L2 (%.0f MiB raw per step vs 126 MB L2), no explicit flush" % (raw.size / 2**20),
                   "parallelism": "rsi-shards x%d" % world, "shard_bits": shard_bits,
                   "decode_rsis_handed_to_careful_kernel": handover, "rsis_per_gpu": nrsi},
        "encode_gbs": enc_gbs * world, "decode_gbs": dec_gbs * world,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": raw.size + comp_bytes + 8 * nrsi,
                "d2h_bytes_per_step": comp_bytes + raw.size + 8 * nrsi,
                "api": "aecb200_encode_host + aecb200_decode_host (what aec_buffer_encode/decode call), pinned host buffers"},
        "gpu_launches": int(launches),
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1")
    ap.add_argument("--mib", type=int, default=256, help="raw MiB per GPU")
    ap.add_argument("--device-only", action="store_true", help="skip the e2e and CPU-baseline legs (for ncu runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
