import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import libaec_b200 as L
from cases import random_case
from libaec_b200 import datagen
from oracle import pyoracle as po
def P(p): return L.Params(p.bits_per_sample, p.block_size, p.rsi, p.flags)
codec = L.DeviceCodec(encode_padding=True)
bad = 0
for rep in range(3):
  for seed in range(250):
    p, raw = random_case(seed, allow_pad=True)
    want = po.orc_encode(p, raw, pad_rsi_build=True)
    src = np.ascontiguousarray(raw)
    cap = L.encode_bound(P(p), src.size) + 16
    out = np.zeros(cap, np.uint8)
    st, n, _ = codec.encode_host(P(p), src.ctypes.data, src.size, out.ctypes.data, cap)
    if st != want["status"] or not np.array_equal(out[:n], want["out"]):
        d = np.nonzero(out[:min(n, want["out"].size)] != want["out"][:min(n, want["out"].size)])[0]
        print("MISMATCH rep", rep, "seed", seed, p, "n", n, want["out"].size, "first diff byte", d[:5], "ndiff", d.size, flush=True)
        bad += 1
print("pad cases bad:", bad)
p, _ = datagen.CONFIGS["c1"]
raw = datagen.generate("c1", (32 << 20) // 4)
want = po.orc_encode(po.Params(p.bits_per_sample, p.block_size, p.rsi, p.flags), raw)
for rep in range(3):
    enc = L.buffer_encode(p, raw)
    same = np.array_equal(enc["out"], want["out"])
    if not same:
        m = min(enc["out"].size, want["out"].size)
        d = np.nonzero(enc["out"][:m] != want["out"][:m])[0]
        print("c1 32MiB mismatch: sizes", enc["out"].size, want["out"].size, "first diffs", d[:8], "ndiff", d.size)
    else:
        print("c1 32MiB ok")
