"""Small encode/decode/scan/shard/SZ calls for compute-sanitizer (memcheck, racecheck, synccheck)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import libaec_b200 as L
from libaec_b200 import datagen
from libaec_b200.parallel import ShardedCodec
from oracle import pyoracle as po
for name, n in (("c1", 1 << 17), ("c2", 1 << 17), ("c3", 1 << 17), ("c4", 100_000), ("c5_noise", 1 << 15), ("c5_restricted", 1 << 16)):
    p, _ = datagen.CONFIGS[name]
    raw = datagen.generate(name, n - 3)
    enc = L.buffer_encode(p, raw, want_offsets=True)
    want = po.orc_encode(po.Params(p.bits_per_sample, p.block_size, p.rsi, p.flags), raw)
    assert enc["status"] == 0 and np.array_equal(enc["out"], want["out"]), name
    dec = L.buffer_decode(p, enc["out"], raw.size)                      # boundary discovery + decode
    assert dec["status"] == 0 and np.array_equal(dec["out"], raw), name
    dec = L.buffer_decode(p, enc["out"], raw.size, offsets=enc["offsets"])
    assert dec["status"] == 0 and np.array_equal(dec["out"], raw), name
    codec = L.DeviceCodec()
    codec.set_scan_mode(2, 4096)
    d_in = torch.from_numpy(np.concatenate([enc["out"], np.zeros(16, np.uint8)])).cuda()
    d_off = torch.zeros(enc["offsets"].size + 2, dtype=torch.int64, device="cuda")
    st, found = codec.scan_offsets(p, d_in, enc["out"].size, d_off, enc["offsets"].size)
    assert st == 0 and found == enc["offsets"].size, name
    codec.close()
p, _ = datagen.CONFIGS["c1"]
raw = datagen.generate("c1", 1 << 17)
sc = ShardedCodec(p, 0, 1, 0, stream=torch.cuda.current_stream().cuda_stream)
sc.step_enqueue(torch.from_numpy(raw).cuda(), raw.size); sc.step_finish(); sc.close()
src = np.cumsum(np.random.default_rng(0).integers(-2, 3, size=80_000)).astype(np.uint8)
z = L.sz_compress(src, src.size * 2 + 1000, 16 | 32 | 128, 64, 8, 1000)
b = L.sz_decompress(z["out"], src.size, 16 | 32 | 128, 64, 8, 1000)
assert np.array_equal(b["out"], src)
print("sanitize workload ok")
