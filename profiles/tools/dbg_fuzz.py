"""Damaged streams: where do the warp-per-RSI decoder, the lane-per-RSI decoder and the oracle disagree?"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import numpy as np, torch
import libaec_b200 as L
from oracle import pyoracle as po
from test_gpu_scan import _multi_rsi_case, _scan, P

codec = L.DeviceCodec()
rng = np.random.default_rng(2024)
for seed in range(120):
    case = _multi_rsi_case(seed)
    if case is None: continue
    p, raw, count = case
    comp = po.orc_encode(p, raw, pad_rsi_build=bool(p.flags & L.AEC_PAD_RSI))["out"].copy()
    R = p.rsi * p.block_size; nrsi = (count + R - 1) // R
    flips = []
    for _ in range(int(rng.integers(1, 6))):
        i = int(rng.integers(0, comp.size)); b = int(rng.integers(0, 8)); comp[i] ^= np.uint8(1 << b); flips.append((i, b))
    st1, off1, _ = _scan(codec, torch, p, comp, nrsi + 3, 1, 0)
    for window in (2048, 1 << 25): _scan(codec, torch, p, comp, nrsi + 3, 2, window)
    B = p.bytes_per_sample
    a = L.buffer_decode(P(p), comp, count * B)
    codec.set_careful_decode(True); codec.set_scan_mode(1, 0)
    pad = (-comp.size) % 4
    d_in = torch.from_numpy(np.concatenate([comp, np.zeros(pad + 8, np.uint8)])).cuda()
    d_off = torch.from_numpy(off1.astype(np.int64)).cuda() if off1.size else torch.zeros(1, dtype=torch.int64, device="cuda")
    d_out = torch.zeros(count * B + 16, dtype=torch.uint8, device="cuda")
    codec.decode_enqueue(P(p), d_in, comp.size, d_off, off1.size, d_out, count * B)
    stc, written = codec.decode_finish()
    codec.set_careful_decode(False)
    o = po.orc_decode(p, comp, count * B)
    car = d_out[:written].cpu().numpy()
    def first_diff(x, y):
        n = min(x.size, y.size); d = np.nonzero(x[:n] != y[:n])[0]
        return int(d[0]) if d.size else (-1 if x.size == y.size else n)
    fc, fo, co = first_diff(a["out"], car), first_diff(a["out"], o["out"]), first_diff(car, o["out"])
    tag = "" if (fc == -1 and fo == -1 and a["status"] == stc == o["status"]) else "  <<<"
    print(f"seed {seed} n={p.bits_per_sample} J={p.block_size} rsi={p.rsi} flags={p.flags:#x} count={count} flips={flips} offs={off1.size}/{nrsi} "
          f"st fast/careful/oracle={a['status']}/{stc}/{o['status']} sizes={a['out'].size}/{written}/{o['out'].size} "
          f"firstdiff f-c={fc} f-o={fo} c-o={co} (sample {fc//B if fc>=0 else -1}, rsi {fc//B//R if fc>=0 else -1}, blk {(fc//B%R)//p.block_size if fc>=0 else -1}){tag}")
