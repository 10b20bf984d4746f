"""Boundary discovery alone (resident stream), best of 5: python time_scan.py c1 c4:256 ...  (AECB200_LIB picks a variant library)"""
import sys, time, os
import numpy as np
import torch
sys.path.insert(0, ".")
import libaec_b200 as L
from libaec_b200 import datagen

out = []
for spec in sys.argv[1:] or ["c1"]:
    name, _, mib = spec.partition(":")
    mib = int(mib or 256)
    p, _ = datagen.CONFIGS[name]
    B = p.bytes_per_sample
    raw = datagen.generate(name, (mib << 20) // B)
    comp = np.ascontiguousarray(L.buffer_encode(p, raw)["out"])
    R = p.rsi * p.block_size
    nrsi = (raw.size // B + R - 1) // R
    codec = L.DeviceCodec()
    d_in = torch.from_numpy(np.concatenate([comp, np.zeros(16 - comp.size % 4, np.uint8)])).cuda()
    d_off = torch.zeros(nrsi + 1, dtype=torch.int64, device="cuda")
    codec.set_scan_mode(2, 0)
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st, found = codec.scan_offsets(p, d_in, comp.size, d_off, nrsi)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    assert st == 0 and found == nrsi, (st, found, nrsi)
    out.append("%s %.2f" % (name, best * 1e3))
    codec.close()
print(os.environ.get("AECB200_LIB", "default").split("/")[-2] if os.environ.get("AECB200_LIB") else "default", "|", os.environ.get("AECB200_SCAN_WINDOW_BITS", "-"), "|", "  ".join(out), flush=True)
