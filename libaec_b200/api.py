"""ctypes mirror of include/libaec.h, include/szlib.h and include/aec_b200.h.

Names, argument meaning and error codes follow the reference's C API
(/root/reference/src/libaec.h:67-166, src/szlib.h:6-43) so the parity tests
read like the reference's own tests (tests/check_aec.c).  Everything here calls
into the native library; nothing is computed in Python and there is no CPU
fallback: if the library (or a CUDA device) is missing the calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

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
AECB200_CUDA_ERROR = -100

AEC_NO_FLUSH = 0
AEC_FLUSH = 1

_PKG = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.path.join(_PKG, "lib")


@dataclass(frozen=True)
class Params:
    """The four coding parameters of struct aec_stream (libaec.h:84-97)."""
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
    def rsi_bytes(self) -> int:
        return self.rsi * self.block_size * self.bytes_per_sample


class AecStream(C.Structure):
    """struct aec_stream (include/libaec.h)."""
    _fields_ = [("next_in", C.c_void_p), ("avail_in", C.c_size_t), ("total_in", C.c_size_t),
                ("next_out", C.c_void_p), ("avail_out", C.c_size_t), ("total_out", C.c_size_t),
                ("bits_per_sample", C.c_uint), ("block_size", C.c_uint), ("rsi", C.c_uint),
                ("flags", C.c_uint), ("state", C.c_void_p)]


class _Params(C.Structure):
    _fields_ = [("bits_per_sample", C.c_uint32), ("block_size", C.c_uint32),
                ("rsi", C.c_uint32), ("flags", C.c_uint32)]


class Carry(C.Structure):
    """aecb200_carry (include/aec_b200.h)."""
    _fields_ = [("bits", C.c_uint64), ("k", C.c_uint32), ("word", C.c_uint32)]


class SZCom(C.Structure):
    _fields_ = [("options_mask", C.c_int), ("bits_per_pixel", C.c_int),
                ("pixels_per_block", C.c_int), ("pixels_per_scanline", C.c_int)]


_lib = None
_libsz = None

LIBAEC_SYMBOLS = ["aec_encode_init", "aec_encode", "aec_encode_end", "aec_decode_init", "aec_decode",
                  "aec_decode_end", "aec_buffer_encode", "aec_buffer_decode",
                  "aec_encode_enable_offsets", "aec_encode_count_offsets", "aec_encode_get_offsets",
                  "aec_decode_set_offsets", "aec_decode_enable_offsets", "aec_decode_count_offsets",
                  "aec_decode_get_offsets", "aec_decode_range"]
DEVICE_SYMBOLS = ["aecb200_device_count", "aecb200_current_device", "aecb200_ctx_device", "aecb200_ctx_create", "aecb200_ctx_destroy",
                  "aecb200_ctx_set_stream", "aecb200_last_error", "aecb200_ctx_set_encode_padding",
                  "aecb200_ctx_launches", "aecb200_encode_bound", "aecb200_encode_device",
                  "aecb200_encode_finish", "aecb200_decode_device", "aecb200_decode_finish",
                  "aecb200_scan_offsets_device", "aecb200_encode_host", "aecb200_encode_host_piece",
                  "aecb200_decode_host", "aecb200_decode_host_resume", "aecb200_ctx_set_shard_mode",
                  "aecb200_encode_shard_info", "aecb200_ctx_set_tile_limit", "aecb200_place_bits_device",
                  "aecb200_encode_device_indexed", "aecb200_decode_device_indexed", "aecb200_group_index_entries",
                  "aecb200_ctx_set_careful_decode", "aecb200_ctx_last_handover",
                  "aecb200_ctx_set_pipeline_piece", "aecb200_ctx_set_scan_mode", "aecb200_ctx_last_scan_fast",
                  "aecb200_ctx_found_offsets", "aecb200_ctx_set_shard_out", "aecb200_shard_plan_device",
                  "aecb200_encode_repair_device", "aecb200_place_bits_planned", "aecb200_set_device",
                  "aecb200_sz_compress_host", "aecb200_sz_decompress_host", "aecb200_sz_compress_batch",
                  "aecb200_sz_decompress_batch", "aecb200_pool_get", "aecb200_pool_put",
                  "aecb200_ctx_accumulate_next", "aecb200_ctx_accumulated_uploads",
                  "aecb200_ctx_stage_input", "aecb200_ctx_staged_uploads"]
SZ_SYMBOLS = ["SZ_BufftoBuffCompress", "SZ_BufftoBuffDecompress", "SZ_encoder_enabled", "SZ_Compress"]


def load_library() -> C.CDLL:
    """Load libaec.so.0; raises when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        path = os.environ.get("AECB200_LIB") or os.path.join(LIBDIR, "libaec.so.0")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -m libaec_b200.build` (needs nvcc)")
        lib = C.CDLL(path)
        for name in LIBAEC_SYMBOLS + DEVICE_SYMBOLS:
            getattr(lib, name).restype = C.c_int
        lib.aecb200_last_error.restype = C.c_char_p
        lib.aecb200_ctx_launches.restype = C.c_uint64
        lib.aecb200_encode_bound.restype = C.c_size_t
        lib.aecb200_ctx_destroy.restype = None
        lib.aecb200_ctx_set_encode_padding.restype = None
        lib.aecb200_ctx_set_shard_mode.restype = None
        lib.aecb200_ctx_set_careful_decode.restype = None
        lib.aecb200_group_index_entries.restype = C.c_size_t
        lib.aecb200_ctx_last_handover.restype = C.c_uint64
        lib.aecb200_ctx_set_tile_limit.restype = None
        lib.aecb200_ctx_set_pipeline_piece.restype = None
        lib.aecb200_ctx_set_scan_mode.restype = None
        lib.aecb200_ctx_last_scan_fast.restype = C.c_uint64
        lib.aecb200_ctx_found_offsets.restype = C.c_size_t
        lib.aecb200_ctx_set_shard_out.restype = None
        lib.aecb200_pool_get.restype = C.c_void_p
        lib.aecb200_pool_put.restype = None
        lib.aecb200_ctx_accumulate_next.restype = None
        lib.aecb200_ctx_accumulated_uploads.restype = C.c_uint64
        lib.aecb200_ctx_staged_uploads.restype = C.c_uint64
        _lib = lib
    return _lib


def load_sz_library() -> C.CDLL:
    global _libsz
    if _libsz is None:
        load_library()
        path = os.path.join(LIBDIR, "libsz.so.2")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -m libaec_b200.build`")
        lib = C.CDLL(path)
        for name in SZ_SYMBOLS[:3]:
            getattr(lib, name).restype = C.c_int
        _libsz = lib
    return _libsz


def _u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def encode_bound(p: Params, nbytes: int) -> int:
    prm = _Params(p.bits_per_sample, p.block_size, p.rsi, p.flags)
    return int(load_library().aecb200_encode_bound(C.byref(prm), C.c_size_t(nbytes)))


def _stream(p: Params) -> AecStream:
    s = AecStream()
    s.bits_per_sample = p.bits_per_sample
    s.block_size = p.block_size
    s.rsi = p.rsi
    s.flags = p.flags
    return s


# --------------------------------------------------------------------------
# whole-buffer calls == aec_buffer_encode / aec_buffer_decode
# --------------------------------------------------------------------------

def buffer_encode(p: Params, data, out_cap: int | None = None, want_offsets: bool = False):
    """aec_buffer_encode on host buffers. Returns dict(status, out, total_in[, offsets])."""
    lib = load_library()
    src = _u8(data)
    cap = encode_bound(p, src.size) if out_cap is None else out_cap
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    s = _stream(p)
    s.next_in = src.ctypes.data
    s.avail_in = src.size
    s.next_out = out.ctypes.data
    s.avail_out = cap
    if not want_offsets:
        st = lib.aec_buffer_encode(C.byref(s))
        return {"status": st, "out": out[:s.total_out].copy(), "total_in": s.total_in}
    st = lib.aec_encode_init(C.byref(s))
    if st != AEC_OK:
        return {"status": st, "out": out[:0], "total_in": 0, "offsets": np.zeros(0, np.uint64)}
    lib.aec_encode_enable_offsets(C.byref(s))
    st = lib.aec_encode(C.byref(s), C.c_int(AEC_FLUSH))
    n = C.c_size_t(0)
    lib.aec_encode_count_offsets(C.byref(s), C.byref(n))
    offs = np.zeros(max(n.value, 1), dtype=np.uint64)
    lib.aec_encode_get_offsets(C.byref(s), offs.ctypes.data_as(C.c_void_p), C.c_size_t(n.value))
    st2 = lib.aec_encode_end(C.byref(s))
    return {"status": st if st != AEC_OK else st2, "out": out[:s.total_out].copy(),
            "total_in": s.total_in, "offsets": offs[:n.value]}


def buffer_decode(p: Params, comp, out_size: int, offsets=None):
    """aec_buffer_decode on host buffers (optionally with an RSI offset index)."""
    lib = load_library()
    src = _u8(comp)
    out = np.zeros(max(out_size, 1), dtype=np.uint8)
    s = _stream(p)
    s.next_in = src.ctypes.data
    s.avail_in = src.size
    s.next_out = out.ctypes.data
    s.avail_out = out_size
    if offsets is None:
        st = lib.aec_buffer_decode(C.byref(s))
    else:
        offs = np.ascontiguousarray(offsets, dtype=np.uint64)
        st = lib.aec_decode_init(C.byref(s))
        if st == AEC_OK:
            lib.aec_decode_set_offsets(C.byref(s), offs.ctypes.data_as(C.c_void_p), C.c_size_t(offs.size))
            st = lib.aec_decode(C.byref(s), C.c_int(AEC_FLUSH))
            lib.aec_decode_end(C.byref(s))
    return {"status": st, "out": out[:s.total_out].copy(), "total_in": s.total_in}


def decode_range(p: Params, comp, offsets, pos: int, size: int):
    """aec_decode_range: `size` bytes of samples from byte `pos` of the uncompressed data."""
    lib = load_library()
    src = _u8(comp)
    out = np.zeros(max(size, 1), dtype=np.uint8)
    offs = np.ascontiguousarray(offsets, dtype=np.uint64)
    s = _stream(p)
    s.next_in = src.ctypes.data
    s.avail_in = src.size
    s.next_out = out.ctypes.data
    s.avail_out = size
    st = lib.aec_decode_init(C.byref(s))
    if st != AEC_OK:
        return {"status": st, "out": out[:0]}
    st = lib.aec_decode_range(C.byref(s), offs.ctypes.data_as(C.c_void_p), C.c_size_t(offs.size),
                              C.c_size_t(pos), C.c_size_t(size))
    lib.aec_decode_end(C.byref(s))
    return {"status": st, "out": out[:s.total_out].copy()}


def buffer_decode_discover(p: Params, comp, out_size: int):
    """aec_decode on a stream without an index, returning the RSI offsets the decoder discovered."""
    lib = load_library()
    src = _u8(comp)
    out = np.zeros(max(out_size, 1), dtype=np.uint8)
    s = _stream(p)
    s.next_in = src.ctypes.data
    s.avail_in = src.size
    s.next_out = out.ctypes.data
    s.avail_out = out_size
    st = lib.aec_decode_init(C.byref(s))
    if st != AEC_OK:
        return {"status": st, "out": out[:0], "offsets": np.zeros(0, np.uint64)}
    lib.aec_decode_enable_offsets(C.byref(s))
    st = lib.aec_decode(C.byref(s), C.c_int(AEC_FLUSH))
    n = C.c_size_t(0)
    lib.aec_decode_count_offsets(C.byref(s), C.byref(n))
    offs = np.zeros(max(n.value, 1), dtype=np.uint64)
    lib.aec_decode_get_offsets(C.byref(s), offs.ctypes.data_as(C.c_void_p), C.c_size_t(n.value))
    lib.aec_decode_end(C.byref(s))
    return {"status": st, "out": out[:s.total_out].copy(), "offsets": offs[:n.value]}


# --------------------------------------------------------------------------
# streaming calls == aec_encode / aec_decode with caller-chosen windows
# --------------------------------------------------------------------------

class _Streamer:
    def __init__(self, p: Params, init, step, end):
        self.lib = load_library()
        self.p = p
        self.s = _stream(p)
        self._step = step
        self._end = end
        self.status = init(C.byref(self.s))
        self.open = self.status == AEC_OK

    def run(self, data, in_chunk: int, out_chunk: int, out_cap: int, flush_at_end: bool = True):
        """Feed `data` in windows of in_chunk bytes, collect output in windows of
        out_chunk bytes (zlib style, like src/aec.c:191-224). Returns the bytes."""
        src = _u8(data)
        out = np.zeros(max(out_cap, 1), dtype=np.uint8)
        s = self.s
        pos = 0
        produced = 0
        while pos < src.size or s.avail_in:
            if s.avail_in == 0:
                n = min(in_chunk, src.size - pos)
                s.next_in = src.ctypes.data + pos
                s.avail_in = n
                pos += n
            while True:
                room = min(out_chunk, out_cap - produced)
                s.next_out = out.ctypes.data + produced
                s.avail_out = room
                before_in = s.avail_in
                st = self._step(C.byref(s), C.c_int(AEC_NO_FLUSH))
                if st != AEC_OK:
                    self.status = st
                    return out[:produced]
                got = room - s.avail_out
                produced += got
                if s.avail_in == 0 and got < room:
                    break
                if got == 0 and s.avail_in == before_in:
                    break
            if s.avail_in and produced >= out_cap:
                break
        if flush_at_end:
            while True:
                room = min(out_chunk, out_cap - produced)
                s.next_out = out.ctypes.data + produced
                s.avail_out = room
                st = self._step(C.byref(s), C.c_int(AEC_FLUSH))
                if st != AEC_OK:
                    self.status = st
                    break
                got = room - s.avail_out
                produced += got
                if got < room or room == 0:
                    break
        return out[:produced]

    def close(self) -> int:
        if self.open:
            self.open = False
            return self._end(C.byref(self.s))
        return AEC_OK


class Encoder(_Streamer):
    def __init__(self, p: Params):
        lib = load_library()
        super().__init__(p, lib.aec_encode_init, lib.aec_encode, lib.aec_encode_end)


class Decoder(_Streamer):
    def __init__(self, p: Params):
        lib = load_library()
        super().__init__(p, lib.aec_decode_init, lib.aec_decode, lib.aec_decode_end)


# --------------------------------------------------------------------------
# SZIP shim
# --------------------------------------------------------------------------

def _sz(fn, src, dest_cap, mask, bpp, ppb, pps):
    s = _u8(src)
    dest = np.zeros(max(dest_cap, 1), dtype=np.uint8)
    dl = C.c_size_t(dest_cap)
    prm = SZCom(mask, bpp, ppb, pps)
    st = fn(dest.ctypes.data_as(C.c_void_p), C.byref(dl), s.ctypes.data_as(C.c_void_p),
            C.c_size_t(s.size), C.byref(prm))
    return {"status": st, "out": dest[:dl.value].copy()}


def sz_compress(src, dest_cap, options_mask, bits_per_pixel, pixels_per_block, pixels_per_scanline):
    return _sz(load_sz_library().SZ_BufftoBuffCompress, src, dest_cap, options_mask, bits_per_pixel,
               pixels_per_block, pixels_per_scanline)


def sz_decompress(src, dest_cap, options_mask, bits_per_pixel, pixels_per_block, pixels_per_scanline):
    return _sz(load_sz_library().SZ_BufftoBuffDecompress, src, dest_cap, options_mask, bits_per_pixel,
               pixels_per_block, pixels_per_scanline)


def _sz_batch(decompress: bool, chunks, dest_caps, options_mask, bits_per_pixel, pixels_per_block,
              pixels_per_scanline, threads: int = 0, dests=None):
    """aecb200_sz_compress_batch / aecb200_sz_decompress_batch: many chunks in flight."""
    lib = load_library()
    n = len(chunks)
    srcs = [_u8(c) for c in chunks]
    if dests is None:
        dests = [np.zeros(max(int(cap), 1), dtype=np.uint8) for cap in dest_caps]
    src_p = (C.c_void_p * n)(*[s.ctypes.data for s in srcs])
    src_l = (C.c_size_t * n)(*[s.size for s in srcs])
    dst_p = (C.c_void_p * n)(*[d.ctypes.data for d in dests])
    dst_l = (C.c_size_t * n)(*[int(cap) for cap in dest_caps])
    status = (C.c_int * n)()
    fn = lib.aecb200_sz_decompress_batch if decompress else lib.aecb200_sz_compress_batch
    rc = fn(C.c_int(n), dst_p, dst_l, src_p, src_l, C.c_int(options_mask), C.c_int(bits_per_pixel),
            C.c_int(pixels_per_block), C.c_int(pixels_per_scanline), status, C.c_int(threads))
    return {"status": rc, "statuses": list(status), "out": [d[:dst_l[i]] for i, d in enumerate(dests)]}


def sz_compress_batch(chunks, dest_caps, options_mask, bits_per_pixel, pixels_per_block, pixels_per_scanline,
                      threads: int = 0, dests=None):
    return _sz_batch(False, chunks, dest_caps, options_mask, bits_per_pixel, pixels_per_block, pixels_per_scanline,
                     threads, dests)


def sz_decompress_batch(chunks, dest_caps, options_mask, bits_per_pixel, pixels_per_block, pixels_per_scanline,
                        threads: int = 0, dests=None):
    return _sz_batch(True, chunks, dest_caps, options_mask, bits_per_pixel, pixels_per_block, pixels_per_scanline,
                     threads, dests)


# --------------------------------------------------------------------------
# device-resident path (aec_b200.h): torch tensors in HBM
# --------------------------------------------------------------------------

class DeviceCodec:
    """One aecb200 context bound to a CUDA device and (optionally) a torch stream."""

    def __init__(self, device: int = -1, stream=None, encode_padding: bool = False):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        st = self.lib.aecb200_ctx_create(C.byref(self.ctx), C.c_int(device))
        if st != AEC_OK:
            raise RuntimeError(f"aecb200_ctx_create failed ({st}): no usable CUDA device, and there is no CPU fallback")
        if stream is None:
            # Callers hand over torch tensors: unless told otherwise run on torch's current stream of the
            # context's device, so that the codec's work is ordered after the fills / copies torch has queued
            # on those tensors (a private stream would race with them).
            import sys
            torch = sys.modules.get("torch")
            if torch is not None and torch.cuda.is_available():
                dev = int(self.lib.aecb200_ctx_device(self.ctx))
                stream = torch.cuda.current_stream(dev).cuda_stream
        if stream is not None:
            self.lib.aecb200_ctx_set_stream(self.ctx, C.c_void_p(int(stream)))
        if encode_padding:
            self.lib.aecb200_ctx_set_encode_padding(self.ctx, C.c_int(1))

    def close(self):
        if self.ctx:
            self.lib.aecb200_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def error(self) -> str:
        return (self.lib.aecb200_last_error(self.ctx) or b"").decode()

    @property
    def launches(self) -> int:
        return int(self.lib.aecb200_ctx_launches(self.ctx))

    def _check(self, st, what):
        if st == AECB200_CUDA_ERROR:
            raise RuntimeError(f"{what}: {self.error()}")
        return st

    # tensors are torch uint8/any-dtype CUDA tensors; only data_ptr()/nbytes are used
    def encode_enqueue(self, p: Params, d_in, in_bytes: int, d_out, d_offsets=None, carry: Carry | None = None,
                       d_grp=None):
        prm = _Params(p.bits_per_sample, p.block_size, p.rsi, p.flags)
        st = self.lib.aecb200_encode_device_indexed(
            self.ctx, C.byref(prm), C.c_void_p(d_in.data_ptr()), C.c_size_t(in_bytes),
            C.c_void_p(d_out.data_ptr()), C.c_size_t(d_out.numel() * d_out.element_size()),
            C.byref(carry) if carry is not None else None,
            C.c_void_p(d_offsets.data_ptr()) if d_offsets is not None else None,
            C.c_void_p(d_grp.data_ptr()) if d_grp is not None else None)
        return self._check(st, "aecb200_encode_device_indexed")

    def group_index_entries(self, p: Params, in_bytes: int) -> int:
        prm = _Params(p.bits_per_sample, p.block_size, p.rsi, p.flags)
        return int(self.lib.aecb200_group_index_entries(C.byref(prm), C.c_size_t(in_bytes)))

    @property
    def last_handover(self) -> int:
        return int(self.lib.aecb200_ctx_last_handover(self.ctx))

    def set_careful_decode(self, on: bool = True):
        self.lib.aecb200_ctx_set_careful_decode(self.ctx, C.c_int(int(on)))

    def set_scan_mode(self, mode: int, window_bits: int = 0):
        """RSI boundary discovery: 0 auto, 1 one-thread scan, 2 parallel tables; window in stream bits."""
        self.lib.aecb200_ctx_set_scan_mode(self.ctx, C.c_int(mode), C.c_uint64(window_bits))

    @property
    def last_scan_fast(self) -> int:
        return int(self.lib.aecb200_ctx_last_scan_fast(self.ctx))

    def set_pipeline_piece(self, raw_bytes: int):
        """Piece size of the host-pointer pipeline (0 = one piece)."""
        self.lib.aecb200_ctx_set_pipeline_piece(self.ctx, C.c_size_t(raw_bytes))

    def encode_finish(self):
        end = Carry()
        st = self._check(self.lib.aecb200_encode_finish(self.ctx, C.byref(end)), "aecb200_encode_finish")
        return st, int(end.bits), int(end.k)

    # ---- multi-GPU shard helpers (aec_b200.h "multi-GPU shards") ----
    def set_shard_mode(self, on: bool = True):
        self.lib.aecb200_ctx_set_shard_mode(self.ctx, C.c_int(int(on)))

    def shard_info(self):
        lo, hi, fc, tail = C.c_uint32(0), C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
        self.lib.aecb200_encode_shard_info(self.ctx, C.byref(lo), C.byref(hi), C.byref(fc), C.byref(tail))
        return int(lo.value), int(hi.value), int(fc.value), int(tail.value)

    def set_shard_out(self, d_info):
        """Device tensor (4 x int64) that receives (bits, klo, khi, tail64) of every shard-mode encode."""
        self.lib.aecb200_ctx_set_shard_out(self.ctx, C.c_void_p(d_info.data_ptr()) if d_info is not None else None)

    def shard_plan(self, d_all, world: int, rank: int, d_plan=None):
        st = self.lib.aecb200_shard_plan_device(self.ctx, C.c_void_p(d_all.data_ptr()), C.c_int(world), C.c_int(rank),
                                                C.c_void_p(d_plan.data_ptr()) if d_plan is not None else None)
        return self._check(st, "aecb200_shard_plan_device")

    def encode_repair(self, p: Params, d_in, in_bytes: int, d_out):
        prm = _Params(p.bits_per_sample, p.block_size, p.rsi, p.flags)
        st = self.lib.aecb200_encode_repair_device(self.ctx, C.byref(prm), C.c_void_p(d_in.data_ptr()), C.c_size_t(in_bytes),
                                                   C.c_void_p(d_out.data_ptr()), C.c_size_t(d_out.numel() * d_out.element_size()))
        return self._check(st, "aecb200_encode_repair_device")

    def place_planned(self, d_src, d_dst, dst_ptr: int | None = None, dst_cap: int | None = None,
                      global_stream: bool = False, last_rank: bool = False):
        ptr = dst_ptr if dst_ptr is not None else d_dst.data_ptr()
        cap = dst_cap if dst_cap is not None else d_dst.numel() * d_dst.element_size()
        st = self.lib.aecb200_place_bits_planned(self.ctx, C.c_void_p(d_src.data_ptr()), C.c_void_p(ptr), C.c_size_t(cap),
                                                 C.c_int(int(global_stream)), C.c_int(int(last_rank)))
        return self._check(st, "aecb200_place_bits_planned")

    def set_tile_limit(self, ntiles: int):
        self.lib.aecb200_ctx_set_tile_limit(self.ctx, C.c_uint64(ntiles))

    def place_bits(self, d_src, nbits: int, d_dst, dst_bit: int, head_or: int = 0):
        st = self.lib.aecb200_place_bits_device(
            self.ctx, C.c_void_p(d_src.data_ptr()), C.c_uint64(nbits), C.c_void_p(d_dst.data_ptr()),
            C.c_size_t(d_dst.numel() * d_dst.element_size()), C.c_uint64(dst_bit), C.c_uint32(head_or))
        return self._check(st, "aecb200_place_bits_device")

    def decode_enqueue(self, p: Params, d_in, in_bytes: int, d_offsets, nrsi: int, d_out, out_bytes: int,
                       d_grp=None):
        prm = _Params(p.bits_per_sample, p.block_size, p.rsi, p.flags)
        st = self.lib.aecb200_decode_device_indexed(
            self.ctx, C.byref(prm), C.c_void_p(d_in.data_ptr()), C.c_size_t(in_bytes),
            C.c_void_p(d_offsets.data_ptr()), C.c_size_t(nrsi),
            C.c_void_p(d_grp.data_ptr()) if d_grp is not None else None,
            C.c_void_p(d_out.data_ptr()), C.c_size_t(out_bytes))
        return self._check(st, "aecb200_decode_device_indexed")

    def decode_finish(self):
        n = C.c_size_t(0)
        st = self._check(self.lib.aecb200_decode_finish(self.ctx, C.byref(n)), "aecb200_decode_finish")
        return st, int(n.value)

    def scan_offsets(self, p: Params, d_in, in_bytes: int, d_offsets, max_rsi: int, start_bit: int = 0):
        prm = _Params(p.bits_per_sample, p.block_size, p.rsi, p.flags)
        found = C.c_size_t(0)
        st = self.lib.aecb200_scan_offsets_device(
            self.ctx, C.byref(prm), C.c_void_p(d_in.data_ptr()), C.c_size_t(in_bytes), C.c_uint64(start_bit),
            C.c_void_p(d_offsets.data_ptr()), C.c_size_t(max_rsi), C.byref(found))
        return self._check(st, "aecb200_scan_offsets_device"), int(found.value)

    # host-pointer calls through the device ABI (pinned or pageable numpy / torch CPU memory)
    def encode_host(self, p: Params, src_ptr: int, in_bytes: int, dst_ptr: int, out_cap: int,
                    offsets_ptr: int = 0, offsets_cap: int = 0):
        prm = _Params(p.bits_per_sample, p.block_size, p.rsi, p.flags)
        out_len = C.c_size_t(0)
        consumed = C.c_size_t(0)
        noff = C.c_size_t(0)
        st = self.lib.aecb200_encode_host(
            self.ctx, C.byref(prm), C.c_void_p(src_ptr), C.c_size_t(in_bytes), C.c_void_p(dst_ptr),
            C.c_size_t(out_cap), C.byref(out_len), C.byref(consumed),
            C.c_void_p(offsets_ptr) if offsets_ptr else None, C.c_size_t(offsets_cap), C.byref(noff))
        return self._check(st, "aecb200_encode_host"), int(out_len.value), int(noff.value)

    def decode_host(self, p: Params, src_ptr: int, in_bytes: int, dst_ptr: int, out_cap: int,
                    offsets_ptr: int = 0, n_offsets: int = 0):
        prm = _Params(p.bits_per_sample, p.block_size, p.rsi, p.flags)
        out_len = C.c_size_t(0)
        st = self.lib.aecb200_decode_host(
            self.ctx, C.byref(prm), C.c_void_p(src_ptr), C.c_size_t(in_bytes),
            C.c_void_p(offsets_ptr) if offsets_ptr else None, C.c_size_t(n_offsets),
            C.c_void_p(dst_ptr), C.c_size_t(out_cap), C.byref(out_len))
        return self._check(st, "aecb200_decode_host"), int(out_len.value)
