"""Seeded parameter/data generators shared by the CPU and GPU parity tests.

The sweep follows the coverage of the reference's own tests
(/root/reference/tests/check_code_options.c:201-283: bits 8/16/24/32, every
block size, every rsi, five flag orderings) and adds what they leave out
(SURVEY section 4): every n in 1..32, LSB and no-preprocessing with large
buffers, AEC_RESTRICTED, AEC_NOT_ENFORCE block sizes, short last RSIs,
partial-sample tails.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.pyoracle import (AEC_DATA_3BYTE, AEC_DATA_MSB, AEC_DATA_PREPROCESS,  # noqa: E402
                             AEC_DATA_SIGNED, AEC_NOT_ENFORCE, AEC_PAD_RSI, AEC_RESTRICTED,
                             Params)

STD_J = (8, 16, 32, 64)
ODD_J = (2, 4, 6, 10, 12, 14, 20, 24, 30, 34, 48, 62)
RSIS = (1, 2, 3, 7, 16, 63, 64, 65, 128, 129, 130, 200, 256, 300)


def pack_samples(vals: np.ndarray, p: Params) -> np.ndarray:
    """n-bit patterns (uint64 array) -> storage bytes in the layout p asks for
    (/root/reference/src/encode_accessors.c:61-143 read back the same)."""
    B = p.bytes_per_sample
    v = vals.astype(np.uint64)
    out = np.zeros((v.size, B), dtype=np.uint8)
    for i in range(B):
        byte = ((v >> np.uint64(8 * i)) & np.uint64(0xFF)).astype(np.uint8)
        if p.flags & AEC_DATA_MSB:
            out[:, B - 1 - i] = byte
        else:
            out[:, i] = byte
    return out.reshape(-1)


def unpack_samples(raw: np.ndarray, p: Params) -> np.ndarray:
    B = p.bytes_per_sample
    b = np.asarray(raw, dtype=np.uint8)[: (len(raw) // B) * B].reshape(-1, B).astype(np.uint64)
    v = np.zeros(b.shape[0], dtype=np.uint64)
    for i in range(B):
        col = b[:, B - 1 - i] if (p.flags & AEC_DATA_MSB) else b[:, i]
        v |= col << np.uint64(8 * i)
    return v


def synth_values(rng: np.random.Generator, n: int, count: int, kind: int, signed: bool) -> np.ndarray:
    """`count` n-bit patterns from one of six distributions."""
    mask = (1 << n) - 1
    lo, hi = (-(1 << (n - 1)), (1 << (n - 1)) - 1) if signed else (0, mask)
    span = hi - lo
    if count == 0:
        return np.zeros(0, dtype=np.uint64)
    if kind == 0:      # uniform noise
        x = rng.integers(lo, hi + 1, size=count, dtype=np.int64)
    elif kind == 1:    # +-3 random walk
        x = lo + span // 2 + np.cumsum(rng.integers(-3, 4, size=count))
    elif kind == 2:    # flat with sparse jumps
        steps = np.where(rng.random(count) < 0.02, rng.integers(-(span // 4) - 1, span // 4 + 2, size=count), 0)
        x = lo + span // 3 + np.cumsum(steps)
    elif kind == 3:    # gaussian walk, amplitude scaled to the range
        sigma = max(1.0, min(span / 64.0, 2000.0))
        x = lo + span // 2 + np.cumsum(np.rint(rng.normal(0, sigma, size=count)).astype(np.int64))
    elif kind == 4:    # 90 % zeros (values sit on the lower bound)
        x = np.where(rng.random(count) < 0.9, lo, lo + rng.integers(0, min(span, 7) + 1, size=count))
    else:              # slow drift with small noise
        x = lo + span // 2 + (np.arange(count) // 7) % max(span // 8, 1) + rng.integers(0, 2, size=count)
    x = np.clip(x, lo, hi).astype(np.int64)
    return (x & mask).astype(np.uint64)


def random_params(rng: np.random.Generator, *, allow_pad: bool = False) -> Params:
    n = int(rng.integers(1, 33))
    flags = 0
    if rng.random() < 0.5 and n > 1:
        flags |= AEC_DATA_SIGNED
    if rng.random() < 0.5:
        flags |= AEC_DATA_MSB
    if rng.random() < 0.75:
        flags |= AEC_DATA_PREPROCESS
    if 17 <= n <= 24 and rng.random() < 0.6:
        flags |= AEC_DATA_3BYTE
    if n <= 4 and rng.random() < 0.5:
        flags |= AEC_RESTRICTED
    if allow_pad and rng.random() < 0.5:
        flags |= AEC_PAD_RSI
    if rng.random() < 0.25:
        flags |= AEC_NOT_ENFORCE
        J = int(rng.choice(ODD_J))
    else:
        J = int(rng.choice(STD_J))
    rsi = int(rng.choice(RSIS))
    return Params(n, J, rsi, flags)


def random_case(seed: int, *, allow_pad: bool = False, max_samples: int = 6000):
    """(Params, raw bytes) for one seeded case."""
    rng = np.random.default_rng(seed)
    p = random_params(rng, allow_pad=allow_pad)
    R = p.rsi * p.block_size
    mode = int(rng.integers(0, 6))
    if mode == 0:
        count = int(rng.integers(1, 4)) * R                # whole RSIs
    elif mode == 1:
        count = int(rng.integers(0, 3)) * R + int(rng.integers(1, R + 1))   # short last RSI
    elif mode == 2:
        count = int(rng.integers(1, max(2, p.block_size)))  # less than a block
    elif mode == 3:
        count = 1
    else:
        count = int(rng.integers(1, max_samples))
    count = min(count, max_samples)
    kind = int(rng.integers(0, 6))
    vals = synth_values(rng, p.bits_per_sample, count, kind, bool(p.flags & AEC_DATA_SIGNED))
    raw = pack_samples(vals, p)
    if rng.random() < 0.15 and p.bytes_per_sample > 1:
        # trailing bytes that do not complete a sample stay unconsumed
        raw = np.concatenate([raw, rng.integers(0, 256, size=int(rng.integers(1, p.bytes_per_sample)), dtype=np.uint8)])
    return p, raw


def reference_test_patterns(p: Params, nbytes: int):
    """The synthetic buffers of check_code_options.c (zero, SE, uncompressed, FS,
    split k) restated: yields (name, expected_first_id, id_bits, raw bytes)."""
    n = p.bits_per_sample
    B = p.bytes_per_sample
    signed = bool(p.flags & AEC_DATA_SIGNED)
    pp = bool(p.flags & AEC_DATA_PREPROCESS)
    xmin = -(1 << (n - 1)) if signed else 0
    xmax = (1 << (n - 1)) - 1 if signed else (1 << n) - 1
    mask = (1 << n) - 1
    idl = p.id_len
    count = nbytes // B

    def tile(pattern):
        reps = (count + len(pattern) - 1) // len(pattern)
        v = np.array([x & mask for x in pattern] * reps, dtype=np.uint64)[:count]
        return pack_samples(v, p)

    if pp:   # check_code_options.c:38-52: memset 0x55 -> constant samples
        const = int.from_bytes(bytes([0x55] * B), "little") & mask
        yield "zero", 0, idl + 1, tile([const])
        yield "se", 1, idl + 1, tile([xmax - 1] * 4 + [xmax] * 4)
        yield "fs", 1, idl, tile([xmin + 2, xmin, xmin, xmin])
    else:
        yield "zero", 0, idl + 1, tile([0])
        yield "se", 1, idl + 1, tile([0, 0, 0, 0, 1, 0, 0, 2])
        yield "fs", 1, idl, tile([0, 0, 0, 4])
    yield "uncomp", (1 << idl) - 1, idl, tile([xmax, xmin])
    for k in range(1, n - 2):
        if pp:
            pat = [xmin + (1 << (k - 1)) - 1, xmin, xmin + (1 << (k + 1)) - 1, xmin]
        else:
            pat = [0, (1 << k) - 1, 0, (1 << (k + 2)) - 1]
        yield f"split{k}", k + 1, idl, tile(pat)
