import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import libaec_b200 as L
from libaec_b200 import datagen
from oracle import pyoracle as po
name, mib = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
p, _ = datagen.CONFIGS[name]
op = po.Params(p.bits_per_sample, p.block_size, p.rsi, p.flags)
ns = (mib << 20) // p.bytes_per_sample - 5
raw = datagen.generate(name, ns)
want = po.ref_encode(op, raw)["out"]
d_want = torch.from_numpy(want).cuda()
codec = L.DeviceCodec()
d_raw = torch.from_numpy(raw).cuda()
cap = (L.encode_bound(p, raw.size) + 64 + 3) // 4 * 4
d_comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
bad_dev = bad_host = 0
for rep in range(reps):
    d_comp.fill_(0xA5); torch.cuda.synchronize()
    codec.encode_enqueue(p, d_raw, raw.size, d_comp)
    st, bits, _ = codec.encode_finish()
    ok = st == 0 and (bits + 7) // 8 == want.size and bool(torch.equal(d_comp[: want.size], d_want))
    if not ok:
        got = d_comp[: want.size].cpu().numpy()
        d = np.nonzero(got != want)[0]
        print(name, "DEVICE rep", rep, "st", st, "bytes", (bits + 7) // 8, want.size, "ndiff", d.size, "first", d[:8], "last", d[-3:], flush=True)
        if d.size:
            i = int(d[0]); print("   got", got[i - 4:i + 12], "want", want[i - 4:i + 12], flush=True)
        bad_dev += 1
    if rep % 4 == 0:
        enc = L.buffer_encode(p, raw)
        if enc["status"] != 0 or enc["out"].size != want.size or not np.array_equal(enc["out"], want):
            m = min(enc["out"].size, want.size)
            d = np.nonzero(enc["out"][:m] != want[:m])[0]
            print(name, "HOST rep", rep, "st", enc["status"], "size", enc["out"].size, want.size, "ndiff", d.size, "first", d[:8], "last", d[-3:], flush=True)
            if d.size:
                i = int(d[0]); print("   got", enc["out"][i - 4:i + 12], "want", want[i - 4:i + 12], flush=True)
            bad_host += 1
print(name, mib, "reps", reps, "bad device", bad_dev, "bad host", bad_host, flush=True)
