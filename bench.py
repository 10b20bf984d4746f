#!/usr/bin/env python
"""bench.py -- AEC encode & decode throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1] [--mib 256]
    python bench.py --impl reference ...        # the reference's CPU implementation, all host cores

A step = one pass of the hot path over one batch of synthetic input: encode the
batch, decode it back (both directions of the metric).  `value` is raw
(uncompressed) GB/s with buffers resident in HBM; `e2e` is the same work through
the reference's own entry points -- aec_buffer_encode then aec_buffer_decode of
include/libaec.h on host buffers, the decode WITHOUT any index (a libaec 0.3.4
caller has none to give: RSI boundaries are discovered on the device), H2D/D2H
inside the timed region.  `e2e` uses pinned host buffers, `e2e_pageable` plain
numpy memory, `e2e_indexed` the offsets extension (aec_encode_enable_offsets /
aec_decode_set_offsets).  One JSON line on stdout (rank 0).
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
    # a step of the reference arm = the same workload and bytes as one GPU's step, whatever N is
    # (bounded sample: the CPU figure does not depend on how many GPUs the other arm uses)
    nbytes = min(args.mib, 256) << 20
    r = cpu_reference_run(args.workload, nbytes, threads, max(1, args.steps), max(0, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": max(1, args.steps), "warmup": max(0, args.warmup), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": bench_config(args),
        "encode_gbs": r["encode_gbs"], "decode_gbs": r["decode_gbs"],
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": "%d MiB of the workload, one RSI-aligned shard per host thread, aec_buffer_encode + aec_buffer_decode" % (r["raw_bytes"] >> 20)},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def bench_config(args):
    """The same dict in both arms (the driver compares them)."""
    return {"workload": workload_name(args), "raw_bytes_per_gpu_step": args.mib << 20,
            "l2": "inputs larger than L2 (%d MiB raw per step vs 126 MB L2), no explicit flush" % args.mib}


def csrc_fingerprint():
    """sha256 over the kernel sources: ties an ncu capture under profiles/ to the build being benchmarked."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "libaec_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".h", ".c")):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode()); h.update(f.read())
    return h.hexdigest()[:16]


def pin_to_gpu_numa(local: int):
    """Run this rank on the CPUs next to its GPU (pinned buffers are then allocated on that node)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        base = "/sys/bus/pci/devices/" + bus
        node = int(open(base + "/numa_node").read())
        cpus = []
        for part in open(base + "/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:                      # no NUMA information: leave the affinity alone
        return {"numa_node": None, "error": type(e).__name__}


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
    numa = pin_to_gpu_numa(local) if not os.environ.get("AECB200_BENCH_NO_PIN") else {"numa_node": None}   # before any pinned allocation
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    p, _ = datagen.CONFIGS[args.workload]
    B = p.bytes_per_sample
    R = p.rsi * p.block_size
    if args.strong:
        total = (args.mib << 20) // B            # strong scaling: --mib is the WHOLE job, split over the ranks
    else:
        total = ((args.mib << 20) // B) * world  # weak scaling: fixed work per GPU
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

    for _ in range(args.warmup):
        step()
        if sharded is not None:
            sharded.step_enqueue(d_raw, raw.size)
            sharded.decode_enqueue(d_back, raw.size)
            sharded.step_finish()
    torch.cuda.synchronize()
    stitch_checked = None
    if sharded is not None:
        # The bytes this rank owns of the ONE stream must be what a single coder writes there: code the
        # shard again, seeded with the stream state the plan says precedes it (bit phase, k, the
        # predecessor's bits of the shared word), and compare.
        sharded.step_enqueue(d_raw, raw.size)
        plan = sharded.step_finish()
        owned = sharded.owned_bytes()
        d_chk = torch.zeros(cap + 8, dtype=torch.uint8, device="cuda")
        codec.encode_enqueue(p, d_raw, raw.size, d_chk, None,
                             carry=L.Carry(plan.bit_offset & 31, plan.k_in, plan.head_or))
        st, _, _ = codec.encode_finish()
        ok = st == 0 and bool(torch.equal(d_chk[: owned.numel()], owned))
        t = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        stitch_checked = bool(t.item())
        assert stitch_checked, "a rank's bytes of the stitched stream differ from the single-coder stream"
        del d_chk
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
            # shard encode, 32-byte all_gather (NCCL, from device memory), plan kernel, k repair and the
            # placement at the global bit phase: enqueues only, no host round trip inside the step
            sharded.step_enqueue(d_raw, raw.size)
            b.record()
            sharded.decode_enqueue(d_back, raw.size)    # decode needs no exchange; runs next to the placement
            sharded.join()
        else:
            codec.encode_enqueue(p, d_raw, raw.size, d_comp, d_offs, d_grp=d_grp)
            b.record()
            codec.decode_enqueue(p, d_comp, comp_bytes, d_offs, nrsi, d_back, raw.size, d_grp=d_grp)
        c.record()
        marks.append((a, b, c))
    t_end.record()
    torch.cuda.synchronize()
    if sharded is not None:
        sharded.step_finish()
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

    # ---- e2e: the reference's own entry points (include/libaec.h) on host buffers ----
    import ctypes as C
    lib = L.load_library()

    def stream_for(src_ptr, src_len, dst_ptr, dst_len):
        s = L.AecStream()
        s.bits_per_sample, s.block_size, s.rsi, s.flags = p.bits_per_sample, p.block_size, p.rsi, p.flags
        s.next_in, s.avail_in, s.next_out, s.avail_out = src_ptr, src_len, dst_ptr, dst_len
        return s

    def plugin_round_trip(raw_ptr, comp_ptr, back_ptr, offs=None):
        """aec_buffer_encode + aec_buffer_decode (offs: the offsets extension instead)."""
        s = stream_for(raw_ptr, raw.size, comp_ptr, cap)
        if offs is None:
            assert lib.aec_buffer_encode(C.byref(s)) == 0
        else:
            assert lib.aec_encode_init(C.byref(s)) == 0
            lib.aec_encode_enable_offsets(C.byref(s))
            assert lib.aec_encode(C.byref(s), C.c_int(L.AEC_FLUSH)) == 0
            n = C.c_size_t(0)
            lib.aec_encode_count_offsets(C.byref(s), C.byref(n))
            assert n.value == nrsi
            lib.aec_encode_get_offsets(C.byref(s), C.c_void_p(offs), C.c_size_t(nrsi))
            assert lib.aec_encode_end(C.byref(s)) == 0
        n_comp = s.total_out
        assert n_comp == comp_bytes
        d = stream_for(comp_ptr, n_comp, back_ptr, raw.size)
        if offs is None:
            assert lib.aec_buffer_decode(C.byref(d)) == 0
        else:
            assert lib.aec_decode_init(C.byref(d)) == 0
            lib.aec_decode_set_offsets(C.byref(d), C.c_void_p(offs), C.c_size_t(nrsi))
            assert lib.aec_decode(C.byref(d), C.c_int(L.AEC_FLUSH)) == 0
            lib.aec_decode_end(C.byref(d))
        assert d.total_out == raw.size

    def time_e2e(fn, steps):
        fn()                                     # untimed first call (allocations inside the library)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / steps
        if dist is not None:
            t = torch.tensor([sec], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return 2 * raw_total / sec / 1e9, sec

    e2e = e2e_pageable = e2e_indexed = pcie = None
    if not args.device_only:
        e2e_steps = max(1, min(args.steps, 5))
        h_raw = torch.from_numpy(raw).pin_memory()
        h_comp = torch.empty(cap, dtype=torch.uint8).pin_memory()
        h_back = torch.empty(raw.size + 16, dtype=torch.uint8).pin_memory()
        h_offs = torch.empty(nrsi, dtype=torch.int64).pin_memory()
        v, sec = time_e2e(lambda: plugin_round_trip(h_raw.data_ptr(), h_comp.data_ptr(), h_back.data_ptr()), e2e_steps)
        assert np.array_equal(h_back[:raw.size].numpy(), raw), "e2e round trip differs"
        e2e = {"value": v, "unit": UNIT, "h2d_bytes_per_step": raw.size + comp_bytes,
               "d2h_bytes_per_step": comp_bytes + raw.size, "ms_per_step": sec * 1e3,
               "api": "aec_buffer_encode + aec_buffer_decode (include/libaec.h, ctypes), pinned host buffers, "
                      "decode without an index (RSI boundaries discovered on the device)"}
        h_back.zero_()
        v, sec = time_e2e(lambda: plugin_round_trip(h_raw.data_ptr(), h_comp.data_ptr(), h_back.data_ptr(),
                                                    offs=h_offs.data_ptr()), e2e_steps)
        assert np.array_equal(h_back[:raw.size].numpy(), raw), "e2e (indexed) round trip differs"
        e2e_indexed = {"value": v, "unit": UNIT, "ms_per_step": sec * 1e3,
                       "api": "aec_encode + aec_encode_get_offsets, aec_decode_set_offsets + aec_decode (offsets extension), pinned"}
        n_comp = np.zeros(cap, np.uint8)
        n_back = np.zeros(raw.size + 16, np.uint8)
        v, sec = time_e2e(lambda: plugin_round_trip(raw.ctypes.data, n_comp.ctypes.data, n_back.ctypes.data), e2e_steps)
        assert np.array_equal(n_back[:raw.size], raw), "e2e (pageable) round trip differs"
        e2e_pageable = {"value": v, "unit": UNIT, "ms_per_step": sec * 1e3,
                        "api": "aec_buffer_encode + aec_buffer_decode, pageable numpy buffers, no index"}
        # what the copies alone cost: the same bytes, same directions, pinned, nothing else
        def copies():
            d_raw.copy_(h_raw, non_blocking=True); h_comp[:comp_bytes].copy_(d_comp[:comp_bytes], non_blocking=True)
            d_comp[:comp_bytes].copy_(h_comp[:comp_bytes], non_blocking=True); h_back[:raw.size].copy_(d_back[:raw.size], non_blocking=True)
            torch.cuda.synchronize()
        v, sec = time_e2e(copies, e2e_steps)
        pcie = {"value": v, "unit": UNIT, "ms_per_step": sec * 1e3,
                "what": "cudaMemcpyAsync of the same bytes and directions back to back (pinned), all ranks at once"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    enc_gbs = raw.size / (enc_mean * 1e-3) / 1e9
    dec_gbs = raw.size / (dec_mean * 1e-3) / 1e9
    algo_bytes = raw.size + comp_bytes            # per launch: raw in + compressed out (encode), reverse for decode
    dominant = "aec_encode_kernel" if enc_mean >= dec_mean else "aec_decode_warp_kernel"
    # dram__bytes_read.sum + dram__bytes_write.sum of that kernel from an `ncu --set full` capture of THIS
    # build (profiles/r2_traffic.json records the fingerprint of the kernel sources it was taken from);
    # null when the sources have changed since
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("csrc") == csrc_fingerprint() and args.workload == tj.get("workload") and args.mib == tj.get("mib"):
            k = tj["kernels"][dominant]
            traffic = k["dram_bytes_read"] + k["dram_bytes_write"]
            traffic_src = tj.get("capture")
    except Exception:
        traffic = None
    dom_ms = max(enc_mean, dec_mean)
    achieved = algo_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
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
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": bench_config(args),
        "detail": {"raw_bytes_per_gpu": raw.size, "compressed_bytes_per_gpu": comp_bytes, "ratio": raw.size / comp_bytes,
                   "parallelism": "rsi-shards x%d" % world, "shard_bits": shard_bits, "stitch_checked": stitch_checked,
                   "decode_rsis_handed_to_careful_kernel": handover, "rsis_per_gpu": nrsi, "numa": numa,
                   "csrc": csrc_fingerprint()},
        "stitch_checked": stitch_checked,
        "encode_gbs": enc_gbs * world, "decode_gbs": dec_gbs * world,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "e2e": e2e, "e2e_indexed": e2e_indexed, "e2e_pageable": e2e_pageable, "pcie_copy_floor": pcie,
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
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: --mib is the size of the whole job (e.g. 8192 for BASELINE config 5), split over the GPUs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
