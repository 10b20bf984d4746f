"""The host+device block code of libaec_b200/csrc/aec_core.cuh and
aec_decode_core.cuh (the same functions the CUDA kernels call), arranged on
the CPU the way the kernels arrange them (tests/_build/libaec_cpumodel.so,
built from csrc/cpu_model.cpp), against the oracle.  This is how the block
logic, tile geometry, scan monoids and boundary-word handling are checked on
machines without a GPU; the GPU parity tests proper are test_gpu_parity.py."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from cases import random_case
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model():
    path = os.path.join(ROOT, "tests", "_build", "libaec_cpumodel.so")
    if not os.path.exists(path):
        from libaec_b200.build import build
        build()
    return C.CDLL(path)


def model_encode(m, p, raw, pad=False, seed=(0, 0, 0)):
    src = np.ascontiguousarray(raw)
    cap = (po.worst_case_bytes(p, src.size) + 64 + 3) // 4 * 4
    out = np.zeros(cap, np.uint8)
    ol, eb, ek = C.c_size_t(0), C.c_uint64(0), C.c_uint32(0)
    R = max(1, p.rsi * p.block_size)
    nrsi = (src.size // p.bytes_per_sample + R - 1) // R
    offs = np.zeros(max(nrsi, 1), np.uint64)
    rc = m.model_encode(C.c_uint32(p.bits_per_sample), C.c_uint32(p.block_size), C.c_uint32(p.rsi),
                        C.c_uint32(p.flags), C.c_int(int(pad)), src.ctypes.data_as(C.c_void_p),
                        C.c_size_t(src.size), out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(ol),
                        offs.ctypes.data_as(C.c_void_p), C.c_uint64(seed[0]), C.c_uint32(seed[1]),
                        C.c_uint32(seed[2]), C.byref(eb), C.byref(ek))
    return rc, out[:ol.value].copy(), offs[:nrsi], eb.value, ek.value


def model_decode(m, p, comp, osz, offs=None):
    src = np.ascontiguousarray(comp)
    out = np.zeros(max(osz, 1) + 8, np.uint8)
    ol = C.c_size_t(0)
    rc = m.model_decode(C.c_uint32(p.bits_per_sample), C.c_uint32(p.block_size), C.c_uint32(p.rsi),
                        C.c_uint32(p.flags), src.ctypes.data_as(C.c_void_p), C.c_size_t(src.size),
                        out.ctypes.data_as(C.c_void_p), C.c_size_t(osz), C.byref(ol),
                        offs.ctypes.data_as(C.c_void_p) if offs is not None else None,
                        C.c_size_t(0 if offs is None else len(offs)))
    return rc, out[:ol.value].copy()


@pytest.mark.parametrize("pad", [False, True])
def test_model_matches_oracle(model, pad):
    for seed in range(1200):
        p, raw = random_case(seed, allow_pad=pad)
        if len(raw) // p.bytes_per_sample == 0:
            continue
        ref = po.orc_encode(p, raw, pad_rsi_build=pad, want_offsets=True)
        rc, out, offs, _, _ = model_encode(model, p, raw, pad)
        assert rc == 0
        assert np.array_equal(out, ref["out"]), (seed, p)
        assert np.array_equal(offs, ref["offsets"]), (seed, p)
        B = p.bytes_per_sample
        ns = len(raw) // B
        for size in (ns * B, (ns // 2) * B, ns * B + 40 * B):
            want = po.orc_decode(p, ref["out"], size)
            rc, got = model_decode(model, p, ref["out"], size)
            assert np.array_equal(got, want["out"]), (seed, p, size)
            if size <= ns * B:
                rc, got = model_decode(model, p, ref["out"], size, ref["offsets"])
                assert np.array_equal(got, want["out"]), (seed, p, size)


def test_model_streaming_carry(model):
    """Coding a stream in two launches with the (bits, k, word) carry gives the
    same bytes as one launch: the contract AEC_NO_FLUSH streaming and the
    multi-GPU shard stitch rely on."""
    for seed in range(300):
        p, raw = random_case(5000 + seed)
        rb = p.rsi * p.block_size * p.bytes_per_sample
        nr = len(raw) // rb
        if nr < 2:
            continue
        whole = po.orc_encode(p, raw)
        cut = (nr // 2) * rb
        rc, a, _, eb, ek = model_encode(model, p, raw[:cut])
        word = 0
        if eb % 32:
            w = a[(eb // 32) * 4:(eb // 32) * 4 + 4].tolist() + [0, 0, 0]
            word = (w[0] << 24) | (w[1] << 16) | (w[2] << 8) | w[3]
        # second launch continues at bit eb of the same buffer
        src = np.ascontiguousarray(raw[cut:])
        cap = (po.worst_case_bytes(p, len(raw)) + 64 + 3) // 4 * 4
        out = np.zeros(cap, np.uint8)
        out[:a.size] = a
        ol, eb2, ek2 = C.c_size_t(0), C.c_uint64(0), C.c_uint32(0)
        rc = model.model_encode(C.c_uint32(p.bits_per_sample), C.c_uint32(p.block_size), C.c_uint32(p.rsi),
                                C.c_uint32(p.flags), C.c_int(0), src.ctypes.data_as(C.c_void_p),
                                C.c_size_t(src.size), out.ctypes.data_as(C.c_void_p), C.c_size_t(cap),
                                C.byref(ol), None, C.c_uint64(eb), C.c_uint32(ek), C.c_uint32(word),
                                C.byref(eb2), C.byref(ek2))
        assert rc == 0
        assert np.array_equal(out[:ol.value], whole["out"]), (seed, p)


def test_model_golden_typical(model):
    rz = np.fromfile(os.path.join(ROOT, "tests", "golden", "typical.rz"), dtype=np.uint8)
    p = po.Params(16, 64, 256, po.AEC_DATA_MSB | po.AEC_DATA_PREPROCESS)
    rc, dec = model_decode(model, p, rz, 1 << 20)
    assert rc == 0 and hashlib.sha256(dec.tobytes()).hexdigest().startswith("e6e1bf68")
    rc, enc, _, _, _ = model_encode(model, p, dec)
    assert rc == 0 and np.array_equal(enc, rz)
