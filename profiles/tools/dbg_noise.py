import sys, hashlib
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import libaec_b200 as L
from libaec_b200 import datagen
from oracle import pyoracle as po
for name, mib in (("c5_noise", 32), ("c1", 64)):
    p, _ = datagen.CONFIGS[name]
    op = po.Params(p.bits_per_sample, p.block_size, p.rsi, p.flags)
    ns = (mib << 20) // p.bytes_per_sample - 5
    raw = datagen.generate(name, ns)
    want = po.ref_encode(op, raw)["out"]
    codec = L.DeviceCodec()
    d_raw = torch.from_numpy(raw).cuda()
    cap = (L.encode_bound(p, raw.size) + 64 + 3) // 4 * 4
    d_comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    bad_dev = bad_host = 0
    for rep in range(12):
        d_comp.fill_(0xA5); torch.cuda.synchronize()
        codec.encode_enqueue(p, d_raw, raw.size, d_comp)
        st, bits, _ = codec.encode_finish()
        got = d_comp[: want.size].cpu().numpy()
        if st != 0 or (bits + 7) // 8 != want.size or not np.array_equal(got, want):
            d = np.nonzero(got != want)[0]
            print(name, "DEVICE rep", rep, "st", st, "bytes", (bits + 7) // 8, want.size, "ndiff", d.size, "first", d[:6], flush=True)
            bad_dev += 1
        enc = L.buffer_encode(p, raw)
        if enc["status"] != 0 or enc["out"].size != want.size or not np.array_equal(enc["out"], want):
            m = min(enc["out"].size, want.size)
            d = np.nonzero(enc["out"][:m] != want[:m])[0]
            print(name, "HOST rep", rep, "st", enc["status"], "size", enc["out"].size, want.size, "ndiff", d.size, "first", d[:6], "last", d[-3:], flush=True)
            bad_host += 1
    print(name, "bad device", bad_dev, "bad host", bad_host, flush=True)
