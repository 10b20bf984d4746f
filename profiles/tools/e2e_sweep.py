#!/usr/bin/env python
"""Host-pointer round trip (aecb200_encode_host + aecb200_decode_host, pinned buffers) of the c1
workload for several pipeline piece sizes.  Usage: python profiles/tools/e2e_sweep.py [MiB ...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import libaec_b200 as L  # noqa: E402
from libaec_b200 import datagen  # noqa: E402

p, _ = datagen.CONFIGS["c1"]
nbytes = 256 << 20
raw = datagen.generate("c1", nbytes // p.bytes_per_sample)
R = p.rsi * p.block_size
nrsi = (raw.size // p.bytes_per_sample + R - 1) // R
cap = L.encode_bound(p, raw.size) + 16
h_raw = torch.from_numpy(raw).pin_memory()
h_comp = torch.empty(cap, dtype=torch.uint8).pin_memory()
h_back = torch.empty(raw.size + 16, dtype=torch.uint8).pin_memory()
h_offs = torch.empty(nrsi, dtype=torch.int64).pin_memory()
codec = L.DeviceCodec(0)
for mib in [float(x) for x in sys.argv[1:]] or [0, 4, 8, 16, 32]:
    codec.set_pipeline_piece(int(mib * (1 << 20)))
    te = td = 0.0
    for it in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st, n, noff = codec.encode_host(p, h_raw.data_ptr(), raw.size, h_comp.data_ptr(), cap, h_offs.data_ptr(), nrsi)
        t1 = time.perf_counter()
        st2, m = codec.decode_host(p, h_comp.data_ptr(), n, h_back.data_ptr(), raw.size, h_offs.data_ptr(), noff)
        t2 = time.perf_counter()
        assert st == 0 and st2 == 0 and m == raw.size
        if it:
            te += t1 - t0; td += t2 - t1
    assert np.array_equal(h_back[:raw.size].numpy(), raw)
    te /= 5; td /= 5
    print("piece %5.1f MiB: encode %.2f ms, decode %.2f ms, round trip %.1f GB/s (raw bytes x2)" %
          (mib, te * 1e3, td * 1e3, 2 * raw.size / (te + td) / 1e9), flush=True)
