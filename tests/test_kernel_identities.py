"""Arithmetic identities two device-only shortcuts of the encoder rest on, checked in numpy with
the kernel's own constants (the shortcuts live under __CUDA_ARCH__, so the CPU model of
test_cpu_model.py does not run them; on the GPU they are covered by the parity tests):

* aec_core.cuh, aec_analyze_block: the five k-window sums S(kb..kb+4) = sum_i d_i >> (kb+m) from
  per-bit sample counts accumulated as nibbles of a table look-up and ten byte dot products;
* aec_encode.cu, mapper: a block whose value range fits between the block and both ends of [0, M]
  cannot clip, so its mapped values are zig-zag codes of the plain differences
  (/root/reference/src/encode.c:255-269 is the exact mapper).
"""
import numpy as np

E_W = (0x40100401, 0x20080200, 0x10040100, 0x08020000, 0x04010000)   # weights of C0 C2 C4 C6 for S1..S5
O_W = (0x80200802, 0x40100401, 0x20080200, 0x10040100, 0x08020000)   # weights of C1 C3 C5 C7


def dp4a(a: int, b: int) -> int:
    return sum(((a >> (8 * i)) & 0xFF) * ((b >> (8 * i)) & 0xFF) for i in range(4))


def lut_entry(v: int) -> int:
    w = 0
    for j in range(8):
        w |= ((v >> j) & 1) << (4 * j)
    return w


def test_window_sums_from_bit_counts():
    rng = np.random.default_rng(7)
    lut = [lut_entry(v) for v in range(256)]
    done = 0
    while done < 3000:
        J = int(rng.choice((8, 16)))
        kb = int(rng.integers(0, 28))
        top = int(rng.integers(1, 256))
        d = rng.integers(0, (top << kb) // J + 1, size=J, dtype=np.uint64)
        if (int(d.sum()) >> kb) > 255 or int(d.max()) >= 1 << 25:
            continue
        if done % 50 == 0:                       # the extremes: every count at its maximum
            d = np.full(J, (255 // J) << kb, dtype=np.uint64)
        acc = [0, 0]
        for i in range(J):
            acc[i >> 3] += lut[int(d[i]) >> kb]
        assert all(((a >> (4 * j)) & 15) <= 8 for a in acc for j in range(8))      # a nibble holds its count
        E = (acc[0] & 0x0F0F0F0F) + (acc[1] & 0x0F0F0F0F)
        O = ((acc[0] >> 4) & 0x0F0F0F0F) + ((acc[1] >> 4) & 0x0F0F0F0F)
        for m in range(5):
            want = int(sum(int(x) >> (kb + m) for x in d))
            assert dp4a(E, E_W[m]) + dp4a(O, O_W[m]) == want, (J, kb, m, d)
        done += 1


def map_exact(u0: int, u1: int, M: int) -> int:
    D = abs(u1 - u0)
    th = min(u0, M - u0)
    return (2 * D - (1 if u1 < u0 else 0)) if D <= th else th + D


def test_clip_free_blocks_are_zigzag_codes():
    rng = np.random.default_rng(11)
    hit = 0
    for _ in range(6000):
        n = int(rng.integers(2, 33))
        M = (1 << n) - 1
        J = int(rng.choice((8, 16, 32, 64)))
        centre = int(rng.integers(0, M + 1))
        spread = int(rng.integers(0, max(1, M // int(rng.choice((2, 3, 8, 64, 1024)))) + 1))
        u = np.clip(centre + rng.integers(-spread, spread + 1, size=J + 1), 0, M).astype(np.int64)
        umin, umax = int(u.min()), int(u.max())
        rngv = umax - umin
        if not (rngv <= umin and umax <= M and rngv <= M - umax):
            continue                                                     # the kernel takes the exact path
        hit += 1
        for i in range(1, J + 1):
            x = (int(u[i]) - int(u[i - 1])) & 0xFFFFFFFF                  # 32-bit wrap-around difference
            xs = x - (1 << 32) if x >> 31 else x
            zig = ((x << 1) & 0xFFFFFFFF) ^ (0xFFFFFFFF if xs < 0 else 0)
            assert zig == map_exact(int(u[i - 1]), int(u[i]), M), (n, u[i - 1], u[i])
        if rngv < (1 << 23):                                             # "small": every code below 2^24
            assert all(map_exact(int(u[i - 1]), int(u[i]), M) < (1 << 24) for i in range(1, J + 1))
    assert hit > 1000


def test_se_cost_wrap_is_unreachable():
    """encode.c:428-429 adds the second-extension cost in uint64 with wrap-around: s*(s+1)/2 overflows when
    a pair sum s reaches 2^32.  The length is compared with the uncompressed length after EVERY pair
    (encode.c:430-431), so a wrapped total could only win if ONE pair's term came out small after the wrap.
    The smallest term (s*(s+1) mod 2^64)/2 + d1 + 1 over all s >= 2^32 is 2^31 + 2, far above the largest
    uncompressed length (64 * 32 bits): replicating the u64 arithmetic (aec_core.cuh) is enough, no stream
    can ever show the difference."""
    import math
    smallest = None
    for m in range(1, 5):                                   # s*(s+1) crosses m * 2^64 near sqrt(m * 2^64)
        root = math.isqrt(m << 64)
        for s in range(max(root - 3, 1 << 32), min(root + 4, (1 << 33) - 1)):
            term = ((s * (s + 1)) % (1 << 64)) // 2 + max(0, s - ((1 << 32) - 1)) + 1
            smallest = term if smallest is None else min(smallest, term)
    # between two crossings the remainder grows by about 2s per step, so the minima sit at the crossings
    assert smallest == (1 << 31) + 2
    assert smallest > 64 * 32
