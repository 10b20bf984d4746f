"""Un-indexed decode timing: the device scan alone (resident stream) and aec_buffer_decode through
host pointers (pageable numpy buffers), per workload."""
import sys, time, json
import numpy as np
import torch
sys.path.insert(0, ".")
import libaec_b200 as L
from libaec_b200 import datagen

def main():
    names = sys.argv[1:] or ["c1"]
    out = {}
    for spec in names:
        name, _, mib = spec.partition(":")
        mib = int(mib or 256)
        p, _ = datagen.CONFIGS[name]
        B = p.bytes_per_sample
        raw = datagen.generate(name, (mib << 20) // B)
        enc = L.buffer_encode(p, raw)
        comp = np.ascontiguousarray(enc["out"])
        R = p.rsi * p.block_size
        nrsi = (raw.size // B + R - 1) // R
        codec = L.DeviceCodec()
        d_in = torch.from_numpy(np.concatenate([comp, np.zeros(16 - comp.size % 4, np.uint8)])).cuda()
        d_off = torch.zeros(nrsi + 1, dtype=torch.int64, device="cuda")
        res = {"raw_mib": mib, "comp_bytes": int(comp.size), "nrsi": nrsi}
        for mode, label in ((2, "scan_parallel_ms"), (1, "scan_serial_ms")):
            if mode == 1 and mib > 64:
                continue
            codec.set_scan_mode(mode, 0)
            best = 1e9
            for _ in range(3):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                st, found = codec.scan_offsets(p, d_in, comp.size, d_off, nrsi)
                torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
            assert st == 0 and found == nrsi, (st, found, nrsi)
            res[label] = best * 1e3
            res[label.replace("_ms", "_fast")] = codec.last_scan_fast
        back = np.zeros(raw.size, np.uint8)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            dec = L.buffer_decode(p, comp, raw.size)
            best = min(best, time.perf_counter() - t0)
        assert dec["status"] == 0 and np.array_equal(dec["out"], raw)
        res["buffer_decode_noindex_ms"] = best * 1e3
        res["buffer_decode_noindex_raw_gbs"] = raw.size / best / 1e9
        out[name] = res
        print(name, json.dumps(res), flush=True)
    json.dump(out, open("gpurun_out/r2_noindex.json", "w"), indent=1)

main()
