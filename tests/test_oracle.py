"""CPU tests that pin the oracle (oracle/aec_oracle.c).

Three anchors, strongest first:
  1. differential against the compiled, unmodified reference (oracle/_ref),
     skipped where that build is absent;
  2. the reference's own golden vector data/typical.rz, both directions;
  3. committed fixtures generated from the reference (tests/golden/).
plus the synthetic per-option buffers of the reference's check_code_options.c
with its first-ID assertion (tests/check_code_options.c:25-31).
"""
import hashlib
import os

import numpy as np
import pytest

from cases import random_case, reference_test_patterns
from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not po.ref_available(), reason="compiled reference not present")

TYPICAL = po.Params(16, 64, 256, po.AEC_DATA_MSB | po.AEC_DATA_PREPROCESS)
RZ_SHA = "16a7f994a672c74daaf42e1e990f96e6001fe7fe0698b637a025686ca03f6068"
RAW_SHA = "e6e1bf684916d765320bc1064c20c5801202e3a6c595c9caca0928e1fe6df896"


def test_typical_rz_both_directions():
    rz = np.fromfile(os.path.join(GOLD, "typical.rz"), dtype=np.uint8)
    assert hashlib.sha256(rz.tobytes()).hexdigest() == RZ_SHA
    dec = po.orc_decode(TYPICAL, rz, 1 << 20)
    assert dec["status"] == 0 and dec["out"].size == 1 << 20
    assert hashlib.sha256(dec["out"].tobytes()).hexdigest() == RAW_SHA
    enc = po.orc_encode(TYPICAL, dec["out"])
    assert enc["status"] == 0
    assert np.array_equal(enc["out"], rz)


def test_committed_reference_vectors(golden):
    meta = golden["meta"]
    for seed, row in enumerate(meta):
        n, J, rsi, flags, pad_build, est, tin, s0, s1, s2, d0, d1, d2 = (int(x) for x in row)
        p = po.Params(n, J, rsi, flags)
        raw = golden[f"raw{seed}"]
        enc = po.orc_encode(p, raw, pad_rsi_build=bool(pad_build))
        assert enc["status"] == est, (seed, p)
        assert enc["total_in"] == tin, (seed, p)
        assert np.array_equal(enc["out"], golden[f"enc{seed}"]), (seed, p)
        for j, (size, dst) in enumerate(((s0, d0), (s1, d1), (s2, d2))):
            dec = po.orc_decode(p, golden[f"enc{seed}"], size)
            assert dec["status"] == dst, (seed, p, size)
            if dst == 0:
                assert np.array_equal(dec["out"], golden[f"dec{seed}_{j}"]), (seed, p, size)


@needs_ref
@pytest.mark.parametrize("pad", [False, True])
def test_differential_against_reference(pad):
    for seed in range(1500):
        p, raw = random_case(seed, allow_pad=pad)
        a = po.orc_encode(p, raw, pad_rsi_build=pad)
        b = po.ref_encode(p, raw, pad_rsi_build=pad)
        assert a["status"] == b["status"], (seed, p)
        assert a["total_in"] == b["total_in"], (seed, p)
        assert np.array_equal(a["out"], b["out"]), (seed, p)
        B = p.bytes_per_sample
        ns = len(raw) // B
        for size in (ns * B, (ns // 2) * B, ns * B + 40 * B + 1):
            da = po.orc_decode(p, b["out"], size)
            db = po.ref_decode(p, b["out"], size)
            assert da["status"] == db["status"], (seed, p, size)
            if db["status"] == 0:
                assert np.array_equal(da["out"], db["out"]), (seed, p, size)


def test_reference_option_patterns_first_id():
    """check_code_options.c restated: round trip + first CDS id."""
    flagsets = [0, po.AEC_DATA_PREPROCESS, po.AEC_DATA_PREPROCESS | po.AEC_DATA_SIGNED,
                po.AEC_DATA_PREPROCESS | po.AEC_DATA_MSB,
                po.AEC_DATA_PREPROCESS | po.AEC_DATA_MSB | po.AEC_DATA_SIGNED]
    for flags in flagsets:
        for n in (8, 16, 24, 32):
            f = flags | (po.AEC_DATA_3BYTE if n == 24 else 0)
            for J in (8, 16, 32, 64):
                for rsi in (1, 2, 5, 48):
                    p = po.Params(n, J, rsi, f)
                    if 3072 // (J * p.bytes_per_sample) < rsi:
                        continue
                    for name, want_id, idbits, raw in reference_test_patterns(p, 3072):
                        enc = po.orc_encode(p, raw)
                        assert enc["status"] == 0
                        assert enc["out"][0] >> (8 - idbits) == want_id, (name, p)
                        dec = po.orc_decode(p, enc["out"], raw.size)
                        assert dec["status"] == 0
                        if not (f & po.AEC_DATA_SIGNED):
                            assert np.array_equal(dec["out"], raw), (name, p)


def test_edge_cases():
    p = po.Params(32, 16, 128, po.AEC_DATA_SIGNED | po.AEC_DATA_PREPROCESS)
    empty = po.orc_encode(p, np.zeros(0, np.uint8))
    assert empty["status"] == 0 and empty["out"].tolist() == [0]       # encode.c:686-695
    assert po.orc_encode(po.Params(0, 16, 128, 0), b"")["status"] == po.AEC_CONF_ERROR
    assert po.orc_encode(po.Params(33, 16, 128, 0), b"")["status"] == po.AEC_CONF_ERROR
    assert po.orc_encode(po.Params(8, 12, 128, 0), b"")["status"] == po.AEC_CONF_ERROR
    assert po.orc_encode(po.Params(8, 12, 128, po.AEC_NOT_ENFORCE), b"")["status"] == 0
    assert po.orc_encode(po.Params(8, 13, 128, po.AEC_NOT_ENFORCE), b"")["status"] == po.AEC_CONF_ERROR
    assert po.orc_encode(po.Params(8, 16, 4097, 0), b"")["status"] == po.AEC_CONF_ERROR
    assert po.orc_encode(po.Params(5, 16, 16, po.AEC_RESTRICTED), b"")["status"] == po.AEC_CONF_ERROR
    raw = np.arange(4096, dtype=np.uint32).view(np.uint8)
    full = po.orc_encode(p, raw)
    short = po.orc_encode(p, raw, out_cap=full["out"].size - 1)
    assert short["status"] == po.AEC_STREAM_ERROR                          # encode.c:944-945
    assert short["out"].size == full["out"].size - 1


@needs_ref
def test_sz_shim_against_reference():
    rng = np.random.default_rng(7)
    for bpp, ppb, pps, mask in [(8, 32, 4096, 16 | 32 | 128 | 1), (8, 16, 1000, 32 | 128),
                                (16, 8, 200, 16 | 32), (32, 8, 1024, 16 | 32 | 128),
                                (64, 8, 1024, 16 | 32 | 128), (12, 10, 77, 32), (24, 32, 300, 16)]:
        nbytes = int(rng.integers(3000, 20000))
        px = 4 if bpp > 16 else (2 if bpp > 8 else 1)
        if bpp in (32, 64):
            src = np.cumsum(rng.integers(-2, 3, size=nbytes)).astype(np.uint8)
            src = src[: (nbytes // (bpp // 8)) * (bpp // 8)]
        else:
            vals = (np.cumsum(rng.integers(-3, 4, size=nbytes // px)) + (1 << (bpp - 1))) & ((1 << bpp) - 1)
            dt = np.dtype({1: np.uint8, 2: np.uint16, 4: np.uint32}[px])
            if mask & 16:
                dt = dt.newbyteorder(">")
            src = vals.astype(dt).view(np.uint8)
        a = po.orc_sz_compress(src, src.size * 2 + 1000, mask, bpp, ppb, pps)
        b = po.ref_sz(True, src, src.size * 2 + 1000, mask, bpp, ppb, pps)
        assert a["status"] == b["status"] == 0
        assert np.array_equal(a["out"], b["out"]), (bpp, ppb, pps)
        da = po.orc_sz_decompress(b["out"], src.size, mask, bpp, ppb, pps)
        db = po.ref_sz(False, b["out"], src.size, mask, bpp, ppb, pps)
        assert da["status"] == db["status"] == 0
        assert np.array_equal(da["out"], db["out"]), (bpp, ppb, pps)
        assert np.array_equal(db["out"], src)
