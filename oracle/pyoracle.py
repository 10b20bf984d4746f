"""ctypes access to the CPU oracle and to the compiled reference.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline legs -- never from libaec_b200/.

  * ``orc_*``  -> oracle/liboracle.so   (our restatement, oracle/aec_oracle.c)
  * ``ref_*``  -> oracle/_ref/libaec_ref.so, libaec_ref_pad.so, libsz_ref.so
                  (the unmodified reference compiled by oracle/Makefile from
                  /root/reference; present only after ``make -C oracle`` ran in
                  a container that has the reference, then shipped prebuilt)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

AEC_DATA_SIGNED = 1
AEC_DATA_3BYTE = 2
AEC_DATA_MSB = 4
AEC_DATA_PREPROCESS = 8
AEC_RESTRICTED = 16
AEC_PAD_RSI = 32
AEC_NOT_ENFORCE = 64

AEC_OK = 0
AEC_CONF_ERROR = -1
AEC_STREAM_ERROR = -2
AEC_DATA_ERROR = -3
AEC_MEM_ERROR = -4


@dataclass(frozen=True)
class Params:
    bits_per_sample: int
    block_size: int
    rsi: int
    flags: int

    @property
    def bytes_per_sample(self) -> int:
        n = self.bits_per_sample
        if n > 16:
            return 3 if (n <= 24 and self.flags & AEC_DATA_3BYTE) else 4
        return 2 if n > 8 else 1

    @property
    def id_len(self) -> int:
        n = self.bits_per_sample
        if n > 16:
            return 5
        if n > 8:
            return 4
        if self.flags & AEC_RESTRICTED:
            return 1 if n <= 2 else 2
        return 3


class _OrcParams(C.Structure):
    _fields_ = [("bits_per_sample", C.c_uint32), ("block_size", C.c_uint32),
                ("rsi", C.c_uint32), ("flags", C.c_uint32)]


class _OrcTrace(C.Structure):
    _fields_ = [("option", C.c_uint8), ("k", C.c_uint8), ("klo", C.c_uint8),
                ("khi", C.c_uint8), ("cds_bits", C.c_uint32)]


class AecStream(C.Structure):
    """struct aec_stream of the reference ABI (src/libaec.h:67-97)."""
    _fields_ = [("next_in", C.c_void_p), ("avail_in", C.c_size_t), ("total_in", C.c_size_t),
                ("next_out", C.c_void_p), ("avail_out", C.c_size_t), ("total_out", C.c_size_t),
                ("bits_per_sample", C.c_uint), ("block_size", C.c_uint), ("rsi", C.c_uint),
                ("flags", C.c_uint), ("state", C.c_void_p)]


class SZCom(C.Structure):
    _fields_ = [("options_mask", C.c_int), ("bits_per_pixel", C.c_int),
                ("pixels_per_block", C.c_int), ("pixels_per_scanline", C.c_int)]


_orc = None
_refs: dict[str, C.CDLL] = {}


def build_oracle() -> None:
    """Compile liboracle.so (and the reference, when its sources are present)."""
    subprocess.run(["make", "-C", HERE, "-s"], check=True)


def _lib() -> C.CDLL:
    global _orc
    if _orc is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        _orc = C.CDLL(path)
        _orc.orc_encode.restype = C.c_int
        _orc.orc_decode.restype = C.c_int
        _orc.orc_sz_compress.restype = C.c_int
        _orc.orc_sz_decompress.restype = C.c_int
    return _orc


def ref_path(name: str) -> str:
    return os.path.join(HERE, "_ref", name)


def ref_available(name: str = "libaec_ref.so") -> bool:
    return os.path.exists(ref_path(name))


def _ref(name: str) -> C.CDLL:
    if name not in _refs:
        lib = C.CDLL(ref_path(name))
        for fn in ("aec_buffer_encode", "aec_buffer_decode", "aec_encode_init", "aec_encode",
                   "aec_encode_end", "aec_decode_init", "aec_decode", "aec_decode_end"):
            getattr(lib, fn).restype = C.c_int
        _refs[name] = lib
    return _refs[name]


def _as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def worst_case_bytes(p: Params, nbytes: int) -> int:
    """Upper bound on the compressed size (every CDS is at most
    id_len + 1 + n + J*n bits)."""
    B = p.bytes_per_sample
    nsamp = nbytes // B
    J = max(p.block_size, 1)
    nblocks = (nsamp + J - 1) // J + max(p.rsi, 1)
    bits = nblocks * (p.id_len + 1 + (J + 1) * p.bits_per_sample)
    return bits // 8 + nblocks // max(p.rsi, 1) + 64


# --------------------------------------------------------------------------
# our restatement
# --------------------------------------------------------------------------

def orc_encode(p: Params, data, *, pad_rsi_build: bool = False, out_cap: int | None = None,
               want_offsets: bool = False, want_trace: bool = False):
    """Returns dict(status, out, total_in, offsets, trace)."""
    lib = _lib()
    src = _as_u8(data)
    cap = worst_case_bytes(p, src.size) if out_cap is None else out_cap
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    op = _OrcParams(p.bits_per_sample, p.block_size, p.rsi, p.flags)
    out_len = C.c_size_t(0)
    consumed = C.c_size_t(0)
    B = p.bytes_per_sample
    R = max(p.rsi * p.block_size, 1)
    nrsi = (src.size // B + R - 1) // R
    offs = np.zeros(max(nrsi, 1), dtype=np.uint64) if want_offsets else None
    noff = C.c_size_t(0)
    ntr = C.c_size_t(0)
    nblk_cap = nrsi * max(p.rsi, 1)
    tr = (_OrcTrace * max(nblk_cap, 1))() if want_trace else None
    st = lib.orc_encode(C.byref(op), C.c_int(1 if pad_rsi_build else 0),
                        src.ctypes.data_as(C.c_void_p), C.c_size_t(src.size),
                        out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(out_len),
                        C.byref(consumed),
                        offs.ctypes.data_as(C.c_void_p) if want_offsets else None,
                        C.c_size_t(nrsi if want_offsets else 0), C.byref(noff),
                        tr if want_trace else None, C.c_size_t(nblk_cap if want_trace else 0),
                        C.byref(ntr))
    res = {"status": st, "out": out[:out_len.value].copy(), "total_in": consumed.value}
    if want_offsets:
        res["offsets"] = offs[:noff.value].copy()
    if want_trace:
        res["trace"] = [(t.option, t.k, t.klo, t.khi, t.cds_bits) for t in tr[:ntr.value]]
    return res


def orc_decode(p: Params, comp, out_size: int):
    lib = _lib()
    src = _as_u8(comp)
    out = np.zeros(max(out_size, 1), dtype=np.uint8)
    op = _OrcParams(p.bits_per_sample, p.block_size, p.rsi, p.flags)
    out_len = C.c_size_t(0)
    st = lib.orc_decode(C.byref(op), src.ctypes.data_as(C.c_void_p), C.c_size_t(src.size),
                        out.ctypes.data_as(C.c_void_p), C.c_size_t(out_size), C.byref(out_len))
    return {"status": st, "out": out[:out_len.value].copy()}


def _sz(fn, dest_cap, src, mask, bpp, ppb, pps):
    s = _as_u8(src)
    dest = np.zeros(max(dest_cap, 1), dtype=np.uint8)
    dl = C.c_size_t(dest_cap)
    st = fn(dest.ctypes.data_as(C.c_void_p), C.byref(dl), s.ctypes.data_as(C.c_void_p),
            C.c_size_t(s.size), C.c_int(mask), C.c_int(bpp), C.c_int(ppb), C.c_int(pps))
    return {"status": st, "out": dest[:dl.value].copy()}


def orc_sz_compress(src, dest_cap, mask, bpp, ppb, pps):
    return _sz(_lib().orc_sz_compress, dest_cap, src, mask, bpp, ppb, pps)


def orc_sz_decompress(src, dest_cap, mask, bpp, ppb, pps):
    return _sz(_lib().orc_sz_decompress, dest_cap, src, mask, bpp, ppb, pps)


# --------------------------------------------------------------------------
# the compiled reference
# --------------------------------------------------------------------------

def _stream(p: Params, src: np.ndarray, out: np.ndarray, out_cap: int) -> AecStream:
    s = AecStream()
    s.next_in = src.ctypes.data
    s.avail_in = src.size
    s.next_out = out.ctypes.data
    s.avail_out = out_cap
    s.bits_per_sample = p.bits_per_sample
    s.block_size = p.block_size
    s.rsi = p.rsi
    s.flags = p.flags
    return s


def ref_encode(p: Params, data, *, pad_rsi_build: bool = False, out_cap: int | None = None,
               lib: C.CDLL | None = None):
    lib = lib or _ref("libaec_ref_pad.so" if pad_rsi_build else "libaec_ref.so")
    src = _as_u8(data)
    cap = worst_case_bytes(p, src.size) if out_cap is None else out_cap
    out = np.zeros(max(cap, 1) + 8, dtype=np.uint8)
    s = _stream(p, src, out, cap)
    st = lib.aec_buffer_encode(C.byref(s))
    return {"status": st, "out": out[:s.total_out].copy(), "total_in": s.total_in,
            "avail_in": s.avail_in}


def ref_decode(p: Params, comp, out_size: int, *, lib: C.CDLL | None = None):
    lib = lib or _ref("libaec_ref.so")
    # the reference's fast path may read a few bytes past the consumed
    # position (decode.c:222-286); give it slack so valgrind-clean
    src = _as_u8(comp)
    padded = np.zeros(src.size + 16, dtype=np.uint8)
    padded[:src.size] = src
    out = np.zeros(max(out_size, 1) + 8, dtype=np.uint8)
    s = _stream(p, padded, out, out_size)
    s.avail_in = src.size
    st = lib.aec_buffer_decode(C.byref(s))
    n = s.total_out if st == AEC_OK else 0
    return {"status": st, "out": out[:n].copy(), "total_in": s.total_in}


def ref_sz(compress: bool, src, dest_cap, mask, bpp, ppb, pps, *, lib: C.CDLL | None = None):
    lib = lib or C.CDLL(ref_path("libsz_ref.so"))
    fn = lib.SZ_BufftoBuffCompress if compress else lib.SZ_BufftoBuffDecompress
    fn.restype = C.c_int
    s = _as_u8(src)
    padded = np.zeros(s.size + 16, dtype=np.uint8)
    padded[:s.size] = s
    dest = np.zeros(max(dest_cap, 1) + 8, dtype=np.uint8)
    dl = C.c_size_t(dest_cap)
    prm = SZCom(mask, bpp, ppb, pps)
    st = fn(dest.ctypes.data_as(C.c_void_p), C.byref(dl), padded.ctypes.data_as(C.c_void_p),
            C.c_size_t(s.size), C.byref(prm))
    return {"status": st, "out": dest[:dl.value].copy()}
