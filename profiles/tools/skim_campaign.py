"""CPU campaign for the boundary discovery (tests/_build/libaec_cpumodel.so = the host+device code of aec_skim_core.cuh):
whole, truncated and bit-flipped streams, sparse candidates and dense tables, long jumps on and off, small and large
windows, all against the one-thread scan; group index from the tables against the skimmed one.
    python profiles/tools/skim_campaign.py FIRST_SEED LAST_SEED"""
import sys, ctypes as C, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import pyoracle as po
from oracle.pyoracle import AEC_DATA_SIGNED, AEC_PAD_RSI
from cases import pack_samples, random_params, synth_values
from test_skim_model import scan, scan_grp, sparse_stats
m = C.CDLL(os.path.join(ROOT, "tests", "_build", "libaec_cpumodel.so"))
t0 = time.time(); done = 0; switched = 0; slow_tot = 0
lo, hi = int(sys.argv[1]), int(sys.argv[2])
for seed in range(lo, hi):
    rng = np.random.default_rng(123_000 + seed)
    p = random_params(rng, allow_pad=bool(seed & 1))
    rsi = int(rng.integers(9, 300)) if seed % 5 else int(rng.integers(9, 40))
    p = type(p)(p.bits_per_sample, p.block_size, rsi, p.flags)
    R = p.rsi * p.block_size
    nrsi = int(rng.integers(12, 40))
    if nrsi * R > 300_000: nrsi = max(6, 300_000 // R)
    count = nrsi * R - int(rng.integers(0, R))
    kind = int(rng.integers(0, 6))
    vals = synth_values(rng, p.bits_per_sample, count, kind, bool(p.flags & AEC_DATA_SIGNED))
    if seed % 4 == 0:     # zero stretches of random length at random places
        v2 = vals.copy(); z = vals.min()
        for _ in range(int(rng.integers(1, 12))):
            a = int(rng.integers(0, count)); b = min(count, a + int(rng.integers(1, 6 * R)))
            v2[a:b] = z
        vals = v2
    raw = np.ascontiguousarray(pack_samples(vals, p))
    enc = po.orc_encode(p, raw, want_offsets=True, pad_rsi_build=bool(p.flags & AEC_PAD_RSI))
    comp = enc["out"].copy()
    mode = seed % 3
    if mode == 1:   # truncated
        comp = comp[: int(rng.integers(1, comp.size + 1))]
    elif mode == 2: # damaged
        for _ in range(int(rng.integers(1, 5))):
            i = int(rng.integers(0, comp.size)); comp[i] ^= np.uint8(1 << int(rng.integers(0, 8)))
    o1, e1, _, end1 = scan(m, p, comp, nrsi + 3, 0, 1)
    for sk in (0, 1):
        m.model_set_skip8(sk)
        for window in (max(1024, (comp.size * 8 // int(rng.integers(2, 9))) // 128 * 128), 1 << 25):
            m.model_set_sparse(1)
            o2, e2, fast, end2 = scan(m, p, comp, nrsi + 3, window, 0)
            slow, dw, marked, pos = sparse_stats(m)
            assert e1 == e2 and np.array_equal(o1, o2), ("SPARSE", seed, p, kind, mode, sk, window)
            slow_tot += slow; switched += dw > 0
            m.model_set_sparse(0)
            o3, e3, fast3, _ = scan(m, p, comp, nrsi + 3, window, 0)
            assert e1 == e3 and np.array_equal(o1, o3), ("DENSE", seed, p, kind, mode, sk, window)
            assert fast >= fast3 - 1 or mode == 2, ("FAST", seed, p, fast, fast3)
    m.model_set_skip8(0); m.model_set_sparse(1)
    if mode == 0 and o1.size >= 2:
        n = int(o1.size) - 1
        gs, ref, _ = scan_grp(m, p, comp, n, 1 << 25)
        G = (p.rsi + 31) // 32; used = (p.rsi + G - 1) // G
        assert np.array_equal(gs[:, :used], ref[:, :used]), ("GRP", seed, p)
    done += 1
print("seeds", lo, hi, "done", done, "dense switches", switched, "slow total", slow_tot, "%.0f s" % (time.time() - t0))
