"""N-GPU stream == 1-GPU stream on real GPUs (skipped with fewer than 2)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_stream_equals_single_gpu_stream():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_contexts_on_two_devices_in_one_process():
    """One process, two GPUs: kernels that need more than 48 KiB of dynamic shared memory (J = 64 encoder,
    warp-per-RSI decoder) run on both devices (the opt-in is per device), libaec.h streams follow the
    caller's current device, and no call leaves the current device switched."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    sys.path.insert(0, ROOT)
    import libaec_b200 as L
    from oracle import pyoracle as po
    p = L.Params(32, 64, 128, L.AEC_DATA_PREPROCESS)
    rng = np.random.default_rng(1)
    raw = np.cumsum(rng.integers(-50, 51, size=1 << 20)).astype("<u4").view(np.uint8)
    want = po.orc_encode(po.Params(32, 64, 128, po.AEC_DATA_PREPROCESS), raw, want_offsets=True)
    for dev in (0, 1, 0):
        torch.cuda.set_device(1 - dev)                    # the caller sits on the OTHER device
        codec = L.DeviceCodec(device=dev)
        with torch.cuda.device(dev):
            d_in = torch.from_numpy(raw).cuda()
            cap = (L.encode_bound(p, raw.size) + 64 + 3) // 4 * 4
            d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
            d_offs = torch.empty(want["offsets"].size, dtype=torch.int64, device="cuda")
            d_back = torch.zeros(raw.size + 16, dtype=torch.uint8, device="cuda")
        assert codec.encode_enqueue(p, d_in, raw.size, d_out, d_offs) == 0
        st, bits, _ = codec.encode_finish()
        assert st == 0 and np.array_equal(d_out[: (bits + 7) // 8].cpu().numpy(), want["out"]), dev
        assert codec.decode_enqueue(p, d_out, (bits + 7) // 8, d_offs, d_offs.numel(), d_back, raw.size) == 0
        st, written = codec.decode_finish()
        assert st == 0 and written == raw.size and np.array_equal(d_back[: raw.size].cpu().numpy(), raw), dev
        assert torch.cuda.current_device() == 1 - dev     # untouched by the calls
        codec.close()
    # libaec.h streams run on the device that is current at init (the pool is keyed by device)
    for dev in (1, 0):
        torch.cuda.set_device(dev)
        enc = L.buffer_encode(p, raw)
        assert enc["status"] == 0 and np.array_equal(enc["out"], want["out"]), dev
        dec = L.buffer_decode(p, enc["out"], raw.size)
        assert dec["status"] == 0 and np.array_equal(dec["out"], raw), dev
        assert torch.cuda.current_device() == dev
