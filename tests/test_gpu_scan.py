"""RSI boundary discovery for streams without an index (aec_skim.cu): the parallel tables
(per-bit-position CDS lengths, pointer doubling, one look-up per RSI) against the one-thread scan and
against the offsets the oracle's encoder records, on whole, truncated and padded streams; then the
un-indexed decode of the README workload through aec_buffer_decode."""
import numpy as np
import pytest

import libaec_b200 as L
from cases import pack_samples, random_case, random_params, synth_values
from libaec_b200 import datagen
from oracle import pyoracle as po
from oracle.pyoracle import AEC_DATA_SIGNED

pytestmark = pytest.mark.gpu


def P(p):
    return L.Params(p.bits_per_sample, p.block_size, p.rsi, p.flags)


def _scan(codec, torch, p, comp, max_rsi, mode, window):
    codec.set_scan_mode(mode, window)
    pad = (-comp.size) % 4
    d_in = torch.from_numpy(np.concatenate([comp, np.zeros(pad + 8, np.uint8)])).cuda()
    d_off = torch.zeros(max(max_rsi, 1), dtype=torch.int64, device="cuda")
    st, found = codec.scan_offsets(P(p), d_in, comp.size, d_off, max_rsi)
    return st, d_off[:found].cpu().numpy().astype(np.uint64), codec.last_scan_fast


def _multi_rsi_case(seed):
    rng = np.random.default_rng(77_000 + seed)
    p = random_params(rng, allow_pad=bool(seed & 1))
    R = p.rsi * p.block_size
    if 12 * R > 200_000:
        return None
    count = int(rng.integers(5, 12)) * R + int(rng.integers(0, R))
    vals = synth_values(rng, p.bits_per_sample, count, int(rng.integers(0, 6)), bool(p.flags & AEC_DATA_SIGNED))
    return p, np.ascontiguousarray(pack_samples(vals, p)), count


def test_parallel_scan_matches_serial_scan_and_encoder_offsets():
    import torch
    codec = L.DeviceCodec()
    done = 0
    for seed in range(300):
        case = _multi_rsi_case(seed)
        if case is None:
            continue
        p, raw, count = case
        pad_build = bool(p.flags & L.AEC_PAD_RSI)
        enc = po.orc_encode(p, raw, want_offsets=True, pad_rsi_build=pad_build)
        assert enc["status"] == 0
        comp = enc["out"]
        R = p.rsi * p.block_size
        nrsi = (count + R - 1) // R
        rng = np.random.default_rng(seed)
        for cut in (comp.size, int(rng.integers(1, comp.size + 1))):
            c = np.ascontiguousarray(comp[:cut])
            st1, off1, _ = _scan(codec, torch, p, c, nrsi + 3, 1, 0)
            for window in (1024, 1 << 25):
                st2, off2, fast = _scan(codec, torch, p, c, nrsi + 3, 2, window)
                assert st2 == st1, (seed, p, cut, window)
                assert np.array_equal(off2, off1), (seed, p, cut, window)
                if cut == comp.size and window == 1 << 25 and nrsi > 2:
                    # whole RSIs come from the tables; only the last (short or padded) one may not
                    assert fast >= nrsi - 2, (seed, p, fast, nrsi)
            if cut == comp.size:
                k = min(off1.size, enc["offsets"].size)
                assert k >= nrsi - 1 and np.array_equal(off1[:k], enc["offsets"][:k]), (seed, p)
            # fewer RSIs asked for than the stream holds
            st3, off3, _ = _scan(codec, torch, p, c, 2, 2, 1024)
            assert np.array_equal(off3, off1[:2]), (seed, p, cut)
        done += 1
    assert done > 150
    codec.close()


def test_parallel_scan_small_random_cases():
    """every n, flag set and block size of the sweep, including streams of less than one RSI"""
    import torch
    codec = L.DeviceCodec()
    for seed in range(400):
        p, raw = random_case(seed, allow_pad=True)
        pad_build = bool(p.flags & L.AEC_PAD_RSI)
        enc = po.orc_encode(p, raw, pad_rsi_build=pad_build)
        comp = np.ascontiguousarray(enc["out"])
        if comp.size == 0:
            continue
        st1, off1, _ = _scan(codec, torch, p, comp, 8, 1, 0)
        st2, off2, _ = _scan(codec, torch, p, comp, 8, 2, 2048)
        assert st1 == st2 and np.array_equal(off1, off2), (seed, p)
    codec.close()


@pytest.mark.parametrize("name,mib", [("c1", 32), ("c2", 16), ("c3", 8), ("c4", 24), ("c5_noise", 8), ("c5_restricted", 4)])
def test_unindexed_buffer_decode_of_the_configs(name, mib):
    """aec_buffer_decode (no index: what every libaec.h caller does) == the input, and nearly every
    RSI boundary comes from the tables."""
    import torch
    p, _ = datagen.CONFIGS[name]
    B = p.bytes_per_sample
    raw = datagen.generate(name, (mib << 20) // B)
    enc = L.buffer_encode(p, raw)
    assert enc["status"] == 0
    dec = L.buffer_decode(p, enc["out"], raw.size)
    assert dec["status"] == 0 and dec["out"].size == raw.size
    assert np.array_equal(dec["out"], raw)
    # the group index the discovery wrote from its tables is the one the warp-per-RSI kernel accepts: hardly any
    # RSI goes to the careful kernel
    import ctypes as C
    lib = L.load_library()
    ctx = C.c_void_p(lib.aecb200_pool_get())
    handed = lib.aecb200_ctx_last_handover(ctx)
    lib.aecb200_pool_put(ctx)
    assert handed <= 2, (name, handed)
    codec = L.DeviceCodec()
    R = p.rsi * p.block_size
    nrsi = (raw.size // B + R - 1) // R
    st, off, fast = _scan(codec, torch, p, np.ascontiguousarray(enc["out"]), nrsi, 2, 0)
    assert st == 0 and off.size == nrsi and fast >= nrsi - 2, (name, off.size, nrsi, fast)
    codec.close()


def test_decode_range_matches_slices_of_the_full_decode():
    """aec_decode_range (random access through the RSI index) against slices of the oracle's decode."""
    rng = np.random.default_rng(5)
    done = 0
    for seed in range(120):
        case = _multi_rsi_case(seed)
        if case is None:
            continue
        p, raw, count = case
        pad_build = bool(p.flags & L.AEC_PAD_RSI)
        enc = po.orc_encode(p, raw, want_offsets=True, pad_rsi_build=pad_build)
        B = p.bytes_per_sample
        full = po.orc_decode(p, enc["out"], count * B)
        assert full["status"] == 0
        for _ in range(4):
            s0 = int(rng.integers(0, count))
            s1 = int(rng.integers(s0, count + 1))
            got = L.decode_range(P(p), enc["out"], enc["offsets"], s0 * B, (s1 - s0) * B)
            assert got["status"] == 0, (seed, p, s0, s1)
            assert np.array_equal(got["out"], full["out"][s0 * B: s1 * B]), (seed, p, s0, s1)
        # a range that ends beyond the data
        got = L.decode_range(P(p), enc["out"], enc["offsets"], (count - 1) * B, 2 * B * p.rsi * p.block_size)
        assert got["status"] == L.AEC_DATA_ERROR or got["out"].size >= B, (seed, p)
        assert L.decode_range(P(p), enc["out"], enc["offsets"], 1 if B > 1 else 0, B)["status"] == (L.AEC_CONF_ERROR if B > 1 else 0)
        done += 1
    assert done > 60


def test_decoder_reports_the_offsets_it_discovered():
    """aec_decode_enable_offsets / aec_decode_get_offsets: the index a decode without an index leaves
    behind equals the encoder's, and decoding a range with it works."""
    for name, mib in (("c1", 8), ("c2", 8), ("c4", 6)):
        p, _ = datagen.CONFIGS[name]
        B = p.bytes_per_sample
        raw = datagen.generate(name, (mib << 20) // B - 7)
        enc = L.buffer_encode(p, raw, want_offsets=True)
        dec = L.buffer_decode_discover(p, enc["out"], raw.size)
        assert dec["status"] == 0 and np.array_equal(dec["out"], raw)
        assert np.array_equal(dec["offsets"], enc["offsets"]), name
        R = p.rsi * p.block_size * B
        got = L.decode_range(p, enc["out"], dec["offsets"], 3 * R + 5 * B, 2 * R)
        assert got["status"] == 0 and np.array_equal(got["out"], raw[3 * R + 5 * B: 5 * R + 5 * B])


def test_damaged_streams_parallel_equals_serial_and_nothing_crashes():
    """Bit flips anywhere in the stream.  What the reference delivers for damaged input is not defined by the
    reference itself: its second-extension decoder indexes a 91-entry table with an unchecked fundamental
    sequence (decode.c:589-600) and its inverse predictor runs on values beyond n bits (decode.c:96-131), so
    sample values after the damage are not compared with anything.  What must hold: the tables tell the same
    story as the one-thread scan (same offsets, same status), both decoders finish with a libaec status and
    deliver the same number of bytes when both call the stream good, and no call fails in CUDA."""
    import torch
    codec = L.DeviceCodec()
    rng = np.random.default_rng(2024)
    done = 0
    for seed in range(120):
        case = _multi_rsi_case(seed)
        if case is None:
            continue
        p, raw, count = case
        pad_build = bool(p.flags & L.AEC_PAD_RSI)
        comp = po.orc_encode(p, raw, pad_rsi_build=pad_build)["out"].copy()
        R = p.rsi * p.block_size
        nrsi = (count + R - 1) // R
        nflip = int(rng.integers(1, 6))
        for _ in range(nflip):
            i = int(rng.integers(0, comp.size))
            comp[i] ^= np.uint8(1 << int(rng.integers(0, 8)))
        st1, off1, _ = _scan(codec, torch, p, comp, nrsi + 3, 1, 0)
        for window in (2048, 1 << 25):
            st2, off2, _ = _scan(codec, torch, p, comp, nrsi + 3, 2, window)
            assert st2 == st1 and np.array_equal(off2, off1), (seed, p, window)
        B = p.bytes_per_sample
        a = L.buffer_decode(P(p), comp, count * B)                      # tables + warp-per-RSI decoder
        assert a["status"] in (L.AEC_OK, L.AEC_DATA_ERROR, L.AEC_MEM_ERROR), (seed, p, a["status"])
        codec.set_careful_decode(True)
        codec.set_scan_mode(1, 0)
        pad = (-comp.size) % 4
        d_in = torch.from_numpy(np.concatenate([comp, np.zeros(pad + 8, np.uint8)])).cuda()
        d_off = torch.from_numpy(off1.astype(np.int64)).cuda() if off1.size else torch.zeros(1, dtype=torch.int64, device="cuda")
        d_out = torch.zeros(count * B + 16, dtype=torch.uint8, device="cuda")
        assert codec.decode_enqueue(P(p), d_in, comp.size, d_off, off1.size, d_out, count * B) == 0
        stc, written = codec.decode_finish()
        codec.set_careful_decode(False)
        assert stc in (L.AEC_OK, L.AEC_DATA_ERROR, L.AEC_MEM_ERROR), (seed, p, stc)
        if a["status"] == L.AEC_OK and stc == 0:
            assert a["out"].size == written, (seed, p)
        done += 1
    assert done > 60
    codec.close()
