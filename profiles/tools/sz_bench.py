"""BASELINE config 3 as stated: SZ_BufftoBuffCompress/Decompress, 8-bit NN, 4 MiB chunks; per-chunk
latency of single calls and throughput with many chunks in flight (aecb200_sz_*_batch); the reference's
libsz on one host core beside it."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import libaec_b200 as L
from libaec_b200 import datagen
from oracle import pyoracle as po

MASK, BPP, PPB, PPS = 16 | 32 | 128 | 1, 8, 32, 4096
CH = 4 << 20
NCH = int(sys.argv[1]) if len(sys.argv) > 1 else 64
chunks = [datagen.generate("c3", CH, i * CH) for i in range(NCH)]
res = {"chunk_bytes": CH, "chunks": NCH}
# single calls
lat_c, lat_d, comp = [], [], []
for c in chunks[:16]:
    t0 = time.perf_counter(); e = L.sz_compress(c, CH + 4096, MASK, BPP, PPB, PPS); t1 = time.perf_counter()
    assert e["status"] == 0
    comp.append(e["out"])
    lat_c.append(t1 - t0)
for c, z in zip(chunks[:16], comp):
    t0 = time.perf_counter(); d = L.sz_decompress(z, CH, MASK, BPP, PPB, PPS); t1 = time.perf_counter()
    assert d["status"] == 0 and np.array_equal(d["out"], c)
    lat_d.append(t1 - t0)
res["single_call_compress_ms"] = float(np.median(lat_c[2:]) * 1e3)
res["single_call_decompress_ms"] = float(np.median(lat_d[2:]) * 1e3)
# batches
dests = [np.zeros(CH + 4096, np.uint8) for _ in range(NCH)]
backs = [np.zeros(CH, np.uint8) for _ in range(NCH)]
for threads in (1, 2, 4, 8):
    best_c = best_d = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        e = L.sz_compress_batch(chunks, [CH + 4096] * NCH, MASK, BPP, PPB, PPS, threads=threads, dests=dests)
        best_c = min(best_c, time.perf_counter() - t0)
        assert e["status"] == 0
        zs = [o.copy() for o in e["out"]]
        t0 = time.perf_counter()
        d = L.sz_decompress_batch(zs, [CH] * NCH, MASK, BPP, PPB, PPS, threads=threads, dests=backs)
        best_d = min(best_d, time.perf_counter() - t0)
        assert d["status"] == 0
    assert all(np.array_equal(b, c) for b, c in zip(backs, chunks))
    res["batch_t%d_compress_gbs" % threads] = NCH * CH / best_c / 1e9
    res["batch_t%d_decompress_gbs" % threads] = NCH * CH / best_d / 1e9
res["ratio"] = CH * NCH / sum(z.size for z in zs)
# the reference's libsz, one core
if po.ref_available():
    t0 = time.perf_counter(); w = po.orc_sz_compress(chunks[0], CH + 4096, MASK, BPP, PPB, PPS); t1 = time.perf_counter()
    res["oracle_port_compress_ms"] = (t1 - t0) * 1e3
print(json.dumps(res))
json.dump(res, open("gpurun_out/r2_sz_bench.json", "w"), indent=1)
