"""Deterministic synthetic inputs of the BASELINE.json configs.

Integer-only and counter-based: element i depends only on (seed, i), so any
shard of any size can be generated independently on any rank and the CPU
baseline and the GPU path see identical bytes.  `sm64` is the splitmix64
finaliser of seed ^ i, tri(i, P) = |(i mod 2P) - P|, seeds 0xAEC0000 + config.
"""
from __future__ import annotations

import numpy as np

from .api import (AEC_DATA_3BYTE, AEC_DATA_MSB, AEC_DATA_PREPROCESS, AEC_DATA_SIGNED,
                  AEC_RESTRICTED, Params)

_U = np.uint64


def sm64(i: np.ndarray, seed: int) -> np.ndarray:
    z = i.astype(np.uint64) ^ _U(seed)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> _U(30))) * _U(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _U(27))) * _U(0x94D049BB133111EB)
    return z ^ (z >> _U(31))


def tri(i: np.ndarray, P: int) -> np.ndarray:
    return np.abs((i % (2 * P)).astype(np.int64) - P)


def _index(start: int, count: int) -> np.ndarray:
    return np.arange(start, start + count, dtype=np.int64)


# name -> (Params, storage bytes per sample, description)
CONFIGS = {
    "c1": (Params(32, 16, 128, AEC_DATA_SIGNED | AEC_DATA_PREPROCESS), "README example: int32 signed PP, J16, rsi128, smooth field"),
    "c2": (Params(16, 32, 64, AEC_DATA_PREPROCESS), "uint16 low-entropy imagery, J32, rsi64, zero regions"),
    "c3": (Params(8, 32, 128, AEC_DATA_MSB | AEC_DATA_PREPROCESS), "SZIP 8-bit NN chunks, 32 px/block, 4096 px/scanline"),
    "c4": (Params(24, 32, 128, AEC_DATA_3BYTE | AEC_DATA_MSB | AEC_DATA_PREPROCESS), "GRIB2-style 24-bit 3BYTE|MSB, J32, rsi128"),
    "c5_noise": (Params(32, 16, 128, AEC_DATA_PREPROCESS), "uint32 uniform noise (uncompressed option)"),
    "c5_restricted": (Params(4, 16, 128, AEC_DATA_PREPROCESS | AEC_RESTRICTED), "n=4 restricted option set"),
    "c5_restricted2": (Params(2, 16, 128, AEC_DATA_PREPROCESS | AEC_RESTRICTED), "n=2 restricted option set (id_len 1)"),
}
SEEDS = {"c1": 0xAEC0001, "c2": 0xAEC0002, "c3": 0xAEC0003, "c4": 0xAEC0004,
         "c5_noise": 0xAEC0005, "c5_restricted": 0xAEC0006, "c5_restricted2": 0xAEC0007}


def generate(name: str, nsamples: int, start: int = 0) -> np.ndarray:
    """`nsamples` samples of config `name`, starting at global sample index
    `start`, as the raw storage bytes (uint8 array) the coder consumes."""
    seed = SEEDS[name]
    out = []
    step = 1 << 22
    for s0 in range(start, start + nsamples, step):
        cnt = min(step, start + nsamples - s0)
        i = _index(s0, cnt)
        r = sm64(i, seed)
        if name == "c1":
            x = 3 * tri(i, 8192) + tri(i, 1000003) - 500000 + (r & _U(255)).astype(np.int64) - 128
            out.append(x.astype("<i4").view(np.uint8))
        elif name == "c2":
            row, col = i // 4096, i % 4096
            lit = ((row // 64 + col // 512) % 4) == 0
            a = 1000 + tri(i, 300) // 4 + (r & _U(7)).astype(np.int64)
            b = (((r >> _U(8)) & _U(63)) == 0).astype(np.int64)
            out.append(np.where(lit, a, b).astype("<u2").view(np.uint8))
        elif name == "c3":
            x = 128 + tri(i, 97) // 2 + tri(i // 4096, 50) + (r & _U(3)).astype(np.int64)
            out.append(x.astype(np.uint8))
        elif name == "c4":
            x = (8000000 + 5 * tri(i, 2048) + tri(i, 777777) + (r & _U(1023)).astype(np.int64)) & 0xFFFFFF
            b = np.empty((cnt, 3), dtype=np.uint8)
            b[:, 0] = (x >> 16) & 0xFF
            b[:, 1] = (x >> 8) & 0xFF
            b[:, 2] = x & 0xFF
            out.append(b.reshape(-1))
        elif name == "c5_noise":
            out.append((r & _U(0xFFFFFFFF)).astype("<u4").view(np.uint8))
        elif name == "c5_restricted":
            x = (tri(i, 40) // 10 + (r & _U(1)).astype(np.int64)) & 15
            out.append(x.astype(np.uint8))
        elif name == "c5_restricted2":
            x = (tri(i, 40) // 14 + (((r & _U(15)) == 0)).astype(np.int64)) & 3
            out.append(x.astype(np.uint8))
        else:
            raise KeyError(name)
    return np.concatenate(out) if out else np.zeros(0, np.uint8)
