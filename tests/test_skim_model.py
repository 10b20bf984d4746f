"""RSI boundary discovery without a GPU: the host+device code of csrc/aec_skim_core.cuh (per-position CDS
entries by rank/select, pointer doubling, RSI lengths by greedy descent, the walk) run on the CPU by
tests/_build/libaec_cpumodel.so, arranged in windows and tiles like aec_skim.cu arranges it, against the
one-thread scan and the offsets the oracle's encoder records."""
import ctypes as C
import os

import numpy as np
import pytest

from cases import pack_samples, random_case, random_params, synth_values
from oracle import pyoracle as po
from oracle.pyoracle import AEC_DATA_PREPROCESS, AEC_DATA_SIGNED, AEC_PAD_RSI

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model():
    from libaec_b200.build import build
    build()
    return C.CDLL(os.path.join(ROOT, "tests", "_build", "libaec_cpumodel.so"))


def scan(m, p, comp, max_rsi, window, serial, start_bit=0):
    src = np.ascontiguousarray(comp)
    offs = np.zeros(max(max_rsi, 1), np.uint64)
    found, flags, fast, end = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    rc = m.model_scan_offsets(C.c_uint32(p.bits_per_sample), C.c_uint32(p.block_size), C.c_uint32(p.rsi),
                              C.c_uint32(p.flags), src.ctypes.data_as(C.c_void_p), C.c_size_t(src.size),
                              C.c_uint64(start_bit), offs.ctypes.data_as(C.c_void_p), C.c_uint64(max_rsi),
                              C.c_uint64(window), C.c_int(serial), C.byref(found), C.byref(flags), C.byref(fast),
                              C.byref(end))
    assert rc == 0
    return offs[:found.value].copy(), flags.value & 2, fast.value, end.value


def multi_rsi_case(seed, max_total=24_000):
    rng = np.random.default_rng(77_000 + seed)
    p = random_params(rng, allow_pad=bool(seed & 1))
    R = p.rsi * p.block_size
    if 6 * R > max_total:
        return None
    count = int(rng.integers(3, 6)) * R + int(rng.integers(0, R))
    vals = synth_values(rng, p.bits_per_sample, count, int(rng.integers(0, 6)), bool(p.flags & AEC_DATA_SIGNED))
    return p, np.ascontiguousarray(pack_samples(vals, p)), count


def test_tables_match_serial_scan_and_encoder_offsets(model):
    done = 0
    for seed in range(120):
        case = multi_rsi_case(seed)
        if case is None:
            continue
        p, raw, count = case
        enc = po.orc_encode(p, raw, want_offsets=True, pad_rsi_build=bool(p.flags & AEC_PAD_RSI))
        assert enc["status"] == 0
        comp = enc["out"]
        R = p.rsi * p.block_size
        nrsi = (count + R - 1) // R
        rng = np.random.default_rng(seed)
        for cut in (comp.size, int(rng.integers(1, comp.size + 1))):
            c = comp[:cut]
            o1, e1, _, end1 = scan(model, p, c, nrsi + 3, 0, 1)
            for window in (1024, 1 << 25):
                o2, e2, fast, end2 = scan(model, p, c, nrsi + 3, window, 0)
                assert e1 == e2 and np.array_equal(o1, o2), (seed, p, cut, window)
                if cut == comp.size and window == 1 << 25 and nrsi > 2:
                    assert fast >= nrsi - 2, (seed, p, fast, nrsi)
            if cut == comp.size:
                k = min(o1.size, enc["offsets"].size)
                assert k >= nrsi - 1 and np.array_equal(o1[:k], enc["offsets"][:k]), (seed, p)
            o3, _, _, end3 = scan(model, p, c, 2, 2048, 0)
            assert np.array_equal(o3, o1[:2]), (seed, p, cut)
            if o1.size > 2:
                # stopping early leaves the position of the next RSI behind (what a streaming decode resumes from)
                assert end3 == o1[2] or (p.flags & AEC_PAD_RSI and (end3 + 7) // 8 * 8 == o1[2]), (seed, p)
        done += 1
    assert done > 40


def test_tables_small_random_cases(model):
    """every n, flag set and block size of the sweep, streams of less than one RSI included"""
    for seed in range(150):
        p, raw = random_case(seed, allow_pad=True, max_samples=2000)
        enc = po.orc_encode(p, raw, pad_rsi_build=bool(p.flags & AEC_PAD_RSI))
        comp = enc["out"]
        if comp.size == 0:
            continue
        o1, e1, _, _ = scan(model, p, comp, 8, 0, 1)
        o2, e2, _, _ = scan(model, p, comp, 8, 2048, 0)
        assert e1 == e2 and np.array_equal(o1, o2), (seed, p)


def test_tables_resume_inside_the_stream(model):
    """a scan may start at any RSI (streaming decode resumes at the RSI of the next undelivered sample)"""
    for seed in range(40):
        case = multi_rsi_case(seed)
        if case is None:
            continue
        p, raw, count = case
        enc = po.orc_encode(p, raw, want_offsets=True, pad_rsi_build=bool(p.flags & AEC_PAD_RSI))
        offs = enc["offsets"]
        if offs.size < 3:
            continue
        o1, _, _, _ = scan(model, p, enc["out"], 100, 0, 1, start_bit=int(offs[2]))
        o2, _, _, _ = scan(model, p, enc["out"], 100, 1024, 0, start_bit=int(offs[2]))
        assert np.array_equal(o1, o2) and o1[0] == offs[2], (seed, p)


def scan_grp(m, p, comp, max_rsi, window):
    src = np.ascontiguousarray(comp)
    offs = np.zeros(max(max_rsi, 1), np.uint64)
    grp = np.zeros(max(max_rsi, 1) * 32, np.uint64)
    ref = np.zeros(max(max_rsi, 1) * 32, np.uint64)
    found, flags, fast, end = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    rc = m.model_scan_offsets_grp(C.c_uint32(p.bits_per_sample), C.c_uint32(p.block_size), C.c_uint32(p.rsi),
                                  C.c_uint32(p.flags), src.ctypes.data_as(C.c_void_p), C.c_size_t(src.size),
                                  C.c_uint64(0), offs.ctypes.data_as(C.c_void_p), C.c_uint64(max_rsi),
                                  C.c_uint64(window), C.c_int(0), C.byref(found), C.byref(flags), C.byref(fast),
                                  C.byref(end), grp.ctypes.data_as(C.c_void_p), ref.ctypes.data_as(C.c_void_p))
    assert rc == 0
    n = found.value
    return grp[: n * 32].reshape(n, 32), ref[: n * 32].reshape(n, 32), fast.value


def test_group_index_from_the_tables_equals_the_skimmed_one(model):
    """The group index the warp-per-RSI decoder reads (where every lane's blocks start, how many of them
    still belong to an earlier zero run), written from the chain tables, against the one a skim of every
    RSI gives: zero-run heavy data, run-of-zero-segment codes, short last RSIs, padded streams."""
    done = 0
    for seed in range(120):
        case = multi_rsi_case(seed)
        if case is None:
            continue
        p, raw, count = case
        enc = po.orc_encode(p, raw, pad_rsi_build=bool(p.flags & AEC_PAD_RSI))
        R = p.rsi * p.block_size
        nrsi = (count + R - 1) // R
        G = (p.rsi + 31) // 32
        used = (p.rsi + G - 1) // G                         # lanes that own blocks
        for window in (4096, 1 << 25):
            grp, ref, fast = scan_grp(model, p, enc["out"], nrsi, window)
            assert grp.shape[0] == nrsi
            assert np.array_equal(grp[:, :used], ref[:, :used]), (seed, p, window)
            if window == 1 << 25 and nrsi > 2:
                assert fast >= nrsi - 2
        done += 1
    assert done > 40


def test_eight_rsi_jumps_of_the_walk(model):
    """Streams of many short RSIs: the walk jumps eight RSIs per look-up (RSI lengths doubled three times) and
    the offsets in between are filled in afterwards; same offsets, same group index."""
    model.model_set_skip8(1)
    try:
        done = 0
        for seed in range(120):
            case = multi_rsi_case(seed, max_total=40_000)
            if case is None:
                continue
            p, raw, count = case
            if p.rsi > 16:
                continue                                    # many RSIs per stream wanted here
            vals = np.tile(raw, 6)[: raw.size * 6 // p.bytes_per_sample * p.bytes_per_sample]
            enc = po.orc_encode(p, vals, want_offsets=True, pad_rsi_build=bool(p.flags & AEC_PAD_RSI))
            R = p.rsi * p.block_size
            nrsi = (vals.size // p.bytes_per_sample + R - 1) // R
            rng = np.random.default_rng(seed)
            for cut in (enc["out"].size, int(rng.integers(1, enc["out"].size + 1))):
                c = enc["out"][:cut]
                o1, e1, _, end1 = scan(model, p, c, nrsi + 3, 0, 1)
                for window in (2048, 1 << 25):
                    o2, e2, fast, end2 = scan(model, p, c, nrsi + 3, window, 0)
                    assert e1 == e2 and np.array_equal(o1, o2), (seed, p, cut, window)
                o3, _, _, _ = scan(model, p, c, 11, 1 << 25, 0)
                assert np.array_equal(o3, o1[:11]), (seed, p, cut)
            grp, ref, fast = scan_grp(model, p, enc["out"], nrsi, 1 << 25)
            G = (p.rsi + 31) // 32
            used = (p.rsi + G - 1) // G
            assert np.array_equal(grp[:, :used], ref[:, :used]), (seed, p)
            assert nrsi < 20 or fast >= nrsi - 9, (seed, p, fast, nrsi)
            done += 1
        assert done > 15
    finally:
        model.model_set_skip8(0)


def sparse_stats(m):
    v = [C.c_uint64(0) for _ in range(4)]
    m.model_sparse_stats(*[C.byref(x) for x in v])
    return tuple(x.value for x in v)                     # slow, dense windows, marked, positions


@pytest.mark.parametrize("skip8", [0, 1])
def test_sparse_candidates_give_the_same_offsets_as_dense_tables(model, skip8):
    """RSI lengths at marked chain ends only (SK_CAND): same offsets, same group index as with a length for every
    position; few positions are marked; zero-heavy streams, whose RSI starts follow no plain chain, make the walk
    work the lengths out itself and then ask for dense tables -- with the same result."""
    model.model_set_skip8(skip8)
    try:
        done = switched = 0
        for seed in range(60):
            rng = np.random.default_rng(91_000 + seed)
            p = random_params(rng, allow_pad=bool(seed & 1))
            rsi = int(rng.integers(9, 130))
            p = type(p)(p.bits_per_sample, p.block_size, rsi, p.flags)
            R = p.rsi * p.block_size
            nrsi = int(rng.integers(24, 60))
            if nrsi * R > 400_000:
                nrsi = max(20, 400_000 // R)
            count = nrsi * R - int(rng.integers(0, R))
            kind = int(rng.integers(0, 6)) if seed % 3 else 4          # every third stream: 90 % zeros
            vals = synth_values(rng, p.bits_per_sample, count, kind, bool(p.flags & AEC_DATA_SIGNED))
            if kind == 4 and seed % 2:                                 # whole zero segments between the data
                v2 = vals.reshape(-1)[: count // R * R].copy()
                zero = vals.min()
                for r0 in range(0, v2.size, R * 2):
                    v2[r0 + R // 3: r0 + R] = zero
                vals = np.concatenate([v2, vals[v2.size:]])
            raw = np.ascontiguousarray(pack_samples(vals, p))
            enc = po.orc_encode(p, raw, want_offsets=True, pad_rsi_build=bool(p.flags & AEC_PAD_RSI))
            assert enc["status"] == 0
            comp = enc["out"]
            window = max(4096, (comp.size * 8 // 5) // 128 * 128)      # about five windows
            o1, e1, _, _ = scan(model, p, comp, nrsi + 3, 0, 1)
            k = min(o1.size, enc["offsets"].size)
            assert np.array_equal(o1[:k], enc["offsets"][:k])
            model.model_set_sparse(0)
            od, ed, fd, _ = scan(model, p, comp, nrsi + 3, window, 0)
            gd, ref, _ = scan_grp(model, p, comp, nrsi, window)
            model.model_set_sparse(1)
            os_, es, fs, _ = scan(model, p, comp, nrsi + 3, window, 0)
            slow, dense_windows, marked, positions = sparse_stats(model)
            gs, _, _ = scan_grp(model, p, comp, nrsi, window)
            assert e1 == ed == es and np.array_equal(o1, od) and np.array_equal(o1, os_), (seed, p, kind)
            G = (p.rsi + 31) // 32
            used = (p.rsi + G - 1) // G
            assert np.array_equal(gs[:, :used], ref[:, :used]) and np.array_equal(gd[:, :used], ref[:, :used]), (seed, p)
            assert fs >= fd - 1, (seed, p, fs, fd)                     # nothing falls back to the CDS-by-CDS skim that did not before
            assert marked <= positions                                 # how few depends on the data: fixed-length CDSs (uncompressed
            #                                                            blocks) shift every position alike and their chains never merge
            done += 1
        assert done == 60
        # RSIs that begin with 70 zero blocks and end with 30 data blocks: fewer than 2^top plain CDSs lead up to every RSI
        # start, nothing marks it; the walk works the lengths out itself and asks for dense tables after the first window
        p = po.Params(8, 8, 100, AEC_DATA_PREPROCESS)
        R, nrsi = p.rsi * p.block_size, 400
        rng = np.random.default_rng(5)
        v = (100 + np.cumsum(rng.integers(-2, 3, size=nrsi * R))).reshape(nrsi, R)
        v[:, : 70 * p.block_size] = v[:, 70 * p.block_size: 70 * p.block_size + 1]
        raw = np.ascontiguousarray(pack_samples((v.reshape(-1) & 255).astype(np.uint64), p))
        comp = po.orc_encode(p, raw)["out"]
        o1, e1, _, _ = scan(model, p, comp, nrsi + 3, 0, 1)
        o2, e2, fast, _ = scan(model, p, comp, nrsi + 3, (comp.size * 8 // 4) // 128 * 128, 0)
        slow, dense_windows, marked, positions = sparse_stats(model)
        assert e1 == e2 and np.array_equal(o1, o2) and fast >= nrsi - 1
        assert slow > 16 and dense_windows >= 2, (slow, dense_windows)
    finally:
        model.model_set_skip8(0)
        model.model_set_sparse(1)
