/*
 * aec_decode.cu -- RSI-parallel CCSDS 121.0-B-2 decoder for sm_100a.
 *
 * Replaces the reference's decoder state machine
 * (/root/reference/src/decode.c:402-677 m_id .. m_uncomp and the FLUSH/put_*
 * post-processing at :67-189) for streams whose RSI start offsets are known
 * (from our encoder's index, or from aec_scan_offsets below):
 *
 *   lane  = one RSI.  It walks its RSI block by block with a position-addressed
 *           bit reader (aec_decode_core.cuh), undoes the predictor (a true
 *           recurrence, so it stays inside the lane) and leaves the block's
 *           samples in a padded shared-memory row;
 *   warp  = 32 consecutive RSIs in lock-step over the block index; after every
 *           block the warp stores its 32 rows cooperatively with 32-bit words
 *           in the sample layout the caller asked for (decode.c:144-189).
 *
 * aec_scan_offsets is the slow path for foreign streams that carry no index:
 * one thread skims CDS after CDS (ids, FS terminators by popcount) and records
 * where every RSI starts.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "aec_decode_core.cuh"
#include "aec_device.h"

namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr int DEC_WARPS = 4;

/* Four consecutive samples of a row -> the 32-bit words of their storage bytes. */
template <int B>
__device__ __forceinline__ void store_group(uint8_t *out, uint64_t sample_idx, const uint32_t *s, uint32_t msb)
{
    /* sample_idx is a multiple of 4 / B' such that the byte address is 4-byte aligned */
    if (B == 4) {
        uint32_t v = s[0];
        reinterpret_cast<uint32_t *>(out)[sample_idx] = msb ? __byte_perm(v, 0, 0x0123) : v;
    } else if (B == 2) {
        uint32_t a = s[0] & 0xFFFFu, b = s[1] & 0xFFFFu;
        uint32_t v = msb ? (__byte_perm(a, b, 0x4501)) : (a | (b << 16));
        reinterpret_cast<uint32_t *>(out)[sample_idx >> 1] = v;
    } else if (B == 1) {
        uint32_t v = (s[0] & 0xFFu) | ((s[1] & 0xFFu) << 8) | ((s[2] & 0xFFu) << 16) | (s[3] << 24);
        reinterpret_cast<uint32_t *>(out)[sample_idx >> 2] = v;
    } else {
        uint32_t a = s[0] & 0xFFFFFFu, b = s[1] & 0xFFFFFFu, cc = s[2] & 0xFFFFFFu, d = s[3] & 0xFFFFFFu;
        if (msb) {
            a = __byte_perm(a, 0, 0x4012); b = __byte_perm(b, 0, 0x4012);
            cc = __byte_perm(cc, 0, 0x4012); d = __byte_perm(d, 0, 0x4012);
        }
        uint32_t *o = reinterpret_cast<uint32_t *>(out) + (sample_idx >> 2) * 3;
        o[0] = a | (b << 24);
        o[1] = (b >> 8) | (cc << 16);
        o[2] = (cc >> 16) | (d << 8);
    }
}

template <int JT, int B>
__global__ void __launch_bounds__(DEC_WARPS * 32)
aec_decode_kernel(const AecDecArgs a)
{
    const AecCfg &c = a.cfg;
    const uint32_t J = JT ? (uint32_t)JT : c.J;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t stride = J + 1;                       /* padded row, conflict free */
    extern __shared__ uint32_t rows[];
    uint32_t *wrows = rows + (size_t)warp * 32u * stride;
    uint32_t *row = wrows + (size_t)lane * stride;

    /* with a list: slot i decodes RSI rsi_list[i] (the RSIs the fast kernel handed over) */
    const uint64_t nslots = a.rsi_list ? (uint64_t)*a.rsi_list_count : a.nrsi;
    const uint64_t slot0 = ((uint64_t)blockIdx.x * DEC_WARPS + warp) * 32ull;
    if (slot0 >= nslots) return;
    const bool have = slot0 + lane < nslots;
    const uint64_t r = have ? (a.rsi_list ? (uint64_t)a.rsi_list[slot0 + lane] : slot0 + lane) : 0;
    const bool contiguous = a.rsi_list == nullptr;
    const uint64_t rsi0 = slot0;

    /* how many samples this RSI has to deliver */
    uint64_t limit = 0;
    if (have) {
        uint64_t startS = r * (uint64_t)c.R;
        limit = a.out_samples > startS ? a.out_samples - startS : 0;
        if (limit > c.R) limit = c.R;
    }
    BitRd br;
    br.init(a.in_words, (a.in_bytes + 3) >> 2, a.in_bytes * 8ull);
    RsiDec st; st.pos = have ? a.rsi_offsets[r] : 0; st.zero_left = 0; st.status = DEC_OK;
    uint32_t u_prev = 0;
    uint32_t delivered = 0;
    bool active = have && limit > 0;

    const uint32_t nblocks = c.rsi;
    for (uint32_t b = 0; b < nblocks; b++) {
        uint32_t cnt = 0;
        if (active) {
            cnt = aec_decode_block<JT>(c, br, st, b, row);
            uint64_t room = limit - delivered;
            if (cnt > room) cnt = (uint32_t)room;
            aec_unmap_row(c, row, cnt, (c.pp && b == 0) ? 1u : 0u, &u_prev);
        }
        __syncwarp();
        /* ---- cooperative store of the 32 rows of this block index ---- */
        const uint32_t fullmask = __ballot_sync(FULL, cnt == J);
        const uint32_t partmask = __ballot_sync(FULL, cnt != J && cnt != 0);
        if (JT != 0 && a.out_aligned && partmask == 0 && contiguous) {
            /* rows are full or empty: flat index over 32*J samples, one 32-bit store per lane-step */
            constexpr int SPG = (B == 4) ? 1 : ((B == 2) ? 2 : 4);       /* samples per 32-bit group */
            constexpr int GPR = (JT ? JT : 4) / SPG;                     /* groups per row */
#pragma unroll 4
            for (int g = (int)lane; g < 32 * GPR; g += 32) {
                int rowi = g / GPR, gi = g % GPR;
                if (!((fullmask >> rowi) & 1u)) continue;
                const uint32_t *src = wrows + (size_t)rowi * stride + gi * SPG;
                uint32_t s[4] = {src[0], SPG > 1 ? src[1] : 0u, SPG > 2 ? src[2] : 0u, SPG > 2 ? src[3] : 0u};
                uint64_t sidx = (rsi0 + rowi) * (uint64_t)c.R + (uint64_t)b * J + (uint64_t)gi * SPG;
                store_group<B>(a.out, sidx, s, c.msb);
            }
        } else if (cnt) {
            /* generic: each lane stores its own row bytewise */
            uint64_t sidx = r * (uint64_t)c.R + (uint64_t)b * J;
            for (uint32_t i = 0; i < cnt; i++)
                aec_store_sample(a.out + (sidx + i) * c.B, row[i], c.B, c.msb);
        }
        __syncwarp();
        delivered += cnt;
        if (active && (cnt < J || delivered >= limit)) active = false;
        if (!__any_sync(FULL, active)) break;
    }
    if (have) {
        if (a.rsi_count) a.rsi_count[r] = delivered;
        if (delivered < limit)
            atomicMax(reinterpret_cast<unsigned long long *>(&a.result[0]),
                      ~(unsigned long long)(r * (uint64_t)c.R + delivered));
        if (st.status == DEC_ERROR)
            atomicOr(reinterpret_cast<unsigned long long *>(&a.result[1]), 1ull);
    }
}

/* ======================================================================== */
/* Fast path: one warp per RSI                                               */
/* ======================================================================== */

constexpr int DW_MAX_WARPS = 12;     /* most warps (= RSIs) per CTA; the host picks the count that fills an SM's shared memory best */

/* Bit reader with a 64-bit left-aligned window refilled one word at a time. */
struct Rd64 {
    const uint32_t *w;
    uint32_t nwords, widx;       /* widx: index of the word held in `nextw` */
    uint32_t nextw;              /* prefetched: loaded one refill ahead so its latency hides behind ~32 bits of decoding */
    uint64_t acc;
    int nb;
    __device__ __forceinline__ uint32_t ldraw(uint32_t i) const { return i < nwords ? __ldg(w + i) : 0u; }
    __device__ __forceinline__ uint32_t ld(uint32_t i) const { return __byte_perm(ldraw(i), 0, 0x0123); }
    __device__ __forceinline__ void init(const uint32_t *base, uint32_t nw, uint64_t bitpos)
    {
        w = base; nwords = nw;
        widx = (uint32_t)(bitpos >> 5);
        uint32_t sh = (uint32_t)(bitpos & 31u);
        acc = (((uint64_t)ld(widx) << 32) | ld(widx + 1)) << sh;
        nb = 64 - (int)sh;
        widx += 2;
        nextw = ldraw(widx);
    }
    __device__ __forceinline__ void refill()
    {
        if (nb <= 32) {
            acc |= (uint64_t)__byte_perm(nextw, 0, 0x0123) << (32 - nb);
            nb += 32;
            widx++;
            nextw = ldraw(widx);      /* consumed at the next refill: the load latency stays off the chain */
        }
    }
    /* bit position of the next unread bit (nextw is not part of the window yet) */
    __device__ __forceinline__ uint64_t pos() const { return (uint64_t)widx * 32ull - (uint64_t)nb; }
    /* n in 1..32 */
    __device__ __forceinline__ uint32_t get(uint32_t n)
    {
        refill();
        uint32_t v = (uint32_t)(acc >> 32) >> (32u - n);
        acc <<= n; nb -= (int)n;
        return v;
    }
    /* unary code; sets *bad when the window runs off the stream */
    __device__ __forceinline__ uint32_t fs(uint32_t *bad)
    {
        uint32_t cnt = 0;
        for (;;) {
            refill();
            uint32_t hi = (uint32_t)(acc >> 32);
            if (hi) {
                uint32_t z = (uint32_t)__clz((int)hi);
                acc <<= (z + 1); nb -= (int)(z + 1);
                return cnt + z;
            }
            cnt += 32; acc <<= 32; nb -= 32;
            if (widx > nwords + 2u) { *bad = 1u; return cnt; }
        }
    }
};

/* One block of the lane's group: fast restatement of aec_decode_block for clean streams (anything
 * unusual sets *bad and the RSI is handed to the careful kernel).
 *
 * The row does not receive the mapped values d themselves but the lane's running sum of the signed
 * steps they stand for when nothing clips (d even: +d/2, d odd: -(d+1)/2): with the lane's start
 * value added, that already is the sample, so clip-free data needs no inverse-predictor pass at all
 * (DESIGN.md 4.2).  d is recoverable from consecutive sums (the step <-> d map is a bijection on 32
 * bits), which is what the exact path does when a clip cannot be ruled out.
 *   pfx        running sum of the lane (wrapping)
 *   pmin/pmax  lowest / highest running sum so far (how far the lane moves from its start value)
 *   big        OR of the lane's d (bounds the largest single step) */
__device__ __forceinline__ uint32_t delta_of(uint32_t dv) { return (dv >> 1) ^ (0u - (dv & 1u)); }   /* +h even, -h odd */

struct LaneAcc {
    uint32_t pfx, big;
    int32_t pmin, pmax;
    __device__ __forceinline__ uint32_t add(uint32_t v)
    {
        pfx += delta_of(v); big |= v;
        pmin = min(pmin, (int32_t)pfx); pmax = max(pmax, (int32_t)pfx);
        return pfx;
    }
};

template <int JT>
__device__ __forceinline__ void warp_decode_block(const AecCfg &c, Rd64 &rd, uint32_t b, uint32_t *row,
                                                  uint32_t &zero_left, uint32_t *bad, LaneAcc &la)
{
    const uint32_t J = JT ? (uint32_t)JT : c.J;
    if (zero_left) {
        zero_left--;
#pragma unroll 4
        for (uint32_t i = 0; i < J; i++) row[i] = la.pfx;
        return;
    }
    const uint32_t ref = (c.pp && b == 0) ? 1u : 0u;
    const uint32_t id = rd.get(c.idl);
    if (id == 0) {
        const uint32_t sel = rd.get(1);
        if (ref) row[0] = rd.get(c.n);
        if (sel == 0) {
            uint32_t zb = rd.fs(bad) + 1u;
            if (zb == 5u) { uint32_t a1 = c.rsi - b, a2 = 64u - (b & 63u); zb = a1 < a2 ? a1 : a2; }
            else if (zb > 5u) zb--;
            if (zb > c.rsi - b) { *bad = 1u; zb = 1; }
            zero_left = zb - 1u;
#pragma unroll 4
            for (uint32_t i = ref; i < J; i++) row[i] = la.pfx;
            return;
        }
        uint32_t i = ref;
        while (i < J) {
            uint32_t m = rd.fs(bad);
            if (m > 90u) { *bad = 1u; m = 0; }
            uint32_t s = (uint32_t)((sqrtf(8.0f * (float)m + 1.0f) - 1.0f) * 0.5f);
            while (s * (s + 1u) / 2u > m) s--;
            while ((s + 1u) * (s + 2u) / 2u <= m) s++;
            uint32_t d1 = m - s * (s + 1u) / 2u;
            if ((i & 1u) == 0) { row[i] = la.add(s - d1); i++; }
            row[i] = la.add(d1); i++;
        }
        return;
    }
    if (id == (1u << c.idl) - 1u) {
        /* J fields of n bits, one after the other: the words they occupy are known now, so they are
         * fetched eight at a time with all loads in flight together.  (Through get() every sample waited
         * for a load issued one refill earlier: with a word consumed per sample the lanes of
         * incompressible data spent 57 % of their stall samples there, profiles/r2_summary.md.) */
        if (c.n == 32u) {
            /* whole words at one bit phase: sample i is a funnel shift of words i and i+1 */
            const uint64_t pos = rd.pos();
            const uint32_t wi = (uint32_t)(pos >> 5), sh = (uint32_t)(pos & 31u);
            uint32_t prevw = rd.ld(wi);
            for (uint32_t i0 = 0; i0 < J; i0 += 8u) {
                uint32_t wv[8];
#pragma unroll
                for (int j = 0; j < 8; j++) wv[j] = rd.ldraw(wi + 1u + i0 + (uint32_t)j);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (i0 + (uint32_t)j < J) {
                        const uint32_t wj = __byte_perm(wv[j], 0, 0x0123);
                        const uint32_t v = __funnelshift_l(wj, prevw, sh);
                        prevw = wj;
                        row[i0 + j] = (i0 + (uint32_t)j >= ref) ? la.add(v) : v;
                    }
                }
            }
            rd.init(rd.w, rd.nwords, pos + 32ull * J);
            return;
        }
        uint32_t i = 0;
        while (i < J) {
            uint32_t wv[8];
            wv[0] = rd.nextw;
#pragma unroll
            for (int j = 1; j < 8; j++) wv[j] = rd.ldraw(rd.widx + (uint32_t)j);
            uint32_t used = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                while (rd.nb >= (int)c.n && i < J) {
                    const uint32_t v = (uint32_t)(rd.acc >> 32) >> (32u - c.n);
                    rd.acc <<= c.n; rd.nb -= (int)c.n;
                    row[i] = (i >= ref) ? la.add(v) : v;
                    i++;
                }
                if (i < J) {                              /* nb < n <= 32: room for the next word */
                    rd.acc |= (uint64_t)__byte_perm(wv[j], 0, 0x0123) << (32 - rd.nb);
                    rd.nb += 32;
                    used = (uint32_t)j + 1u;
                }
            }
            rd.widx += used;
            rd.nextw = rd.ldraw(rd.widx);                 /* the word behind the ones taken (a cache hit) */
        }
        return;
    }
    const uint32_t k = id - 1u;
    if (ref) row[0] = rd.get(c.n);
    /* unary part: after a refill the window holds >= 32 valid bits; the ones in it are the codes
     * that complete inside it (popc), decoded before the 64-bit accumulator is touched again */
    {
        uint32_t i = ref, pend = 0;
        while (i < J) {
            rd.refill();
            uint32_t wnd = (uint32_t)(rd.acc >> 32);
            const uint32_t left = J - i;
            uint32_t n = (uint32_t)__popc(wnd);
            n = n < left ? n : left;
            uint32_t cons = 0;
            if (n) {
                uint32_t z = (uint32_t)__clz((int)wnd);
                row[i] = (pend + z) << k;
                pend = 0;
                wnd = (wnd << z) << 1;
                cons = z + 1u;
                for (uint32_t q = 1; q < n; q++) {
                    z = (uint32_t)__clz((int)wnd);
                    row[i + q] = z << k;
                    wnd = (wnd << z) << 1;
                    cons += z + 1u;
                }
                i += n;
            }
            if (i < J) {                          /* the rest of the window is the start of the next code */
                pend += 32u - cons; cons = 32u;
                if (rd.widx > rd.nwords + 2u) { *bad = 1u; break; }
            }
            rd.acc <<= cons; rd.nb -= (int)cons;
        }
    }
    if (k == 0) {
#pragma unroll 4
        for (uint32_t i = ref; i < J; i++) row[i] = la.add(row[i]);
        return;
    }
    /* binary part: k low bits per sample, fetched four (k <= 8) or two (k <= 16) samples at a time */
    uint32_t i = ref;
    const uint32_t m = (1u << k) - 1u;
    if (k <= 8) {
        for (; i + 4 <= J; i += 4) {
            uint32_t q = rd.get(4u * k);
            uint32_t v0 = row[i] + (q >> (3u * k)), v1 = row[i + 1] + ((q >> (2u * k)) & m);
            uint32_t v2 = row[i + 2] + ((q >> k) & m), v3 = row[i + 3] + (q & m);
            row[i] = la.add(v0); row[i + 1] = la.add(v1); row[i + 2] = la.add(v2); row[i + 3] = la.add(v3);
        }
    } else if (k <= 16) {
        for (; i + 2 <= J; i += 2) {
            uint32_t q = rd.get(2u * k);
            uint32_t v0 = row[i] + (q >> k), v1 = row[i + 1] + (q & m);
            row[i] = la.add(v0); row[i + 1] = la.add(v1);
        }
    }
    for (; i < J; i++) row[i] = la.add(row[i] + rd.get(k));
}

/* First guess of a lane's start value: reference + sum of the deltas before the lane, which is
 * exact when nothing clips.  When samples sit on a range boundary (imagery near zero) the
 * clipped steps make that sum leave [0, M]; the nearest boundary is then the better guess.
 * Only a heuristic for the first walk: the iteration that follows is exact either way. */
__device__ __forceinline__ uint32_t project_start(uint32_t uref, uint32_t off, uint32_t M)
{
    long long v = (long long)uref + (long long)(int32_t)off;
    if (v < 0) v = 0;
    if (v > (long long)M) v = (long long)M;
    return (uint32_t)v;
}

/* exact inverse mapper step on normalised values; sets clip when the clipped branch was taken */
__device__ __forceinline__ uint32_t unmap_step(uint32_t u, uint32_t d, uint32_t M, uint32_t &clip)
{
    uint32_t h = (d >> 1) + (d & 1u);
    uint32_t mu = M - u;
    uint32_t step = (d & 1u) ? (u - h) : (u + h);
    uint32_t cv = (u <= mu) ? d : (M - d);
    bool cl = (u < h) || (mu < h);
    clip |= cl ? 1u : 0u;
    return cl ? cv : step;
}

/* Four (B = 1, 3), two (B = 2) or one (B = 4) consecutive samples of one RSI -> 32-bit stores; s_local is the
 * first sample's index inside the RSI whose output starts at rsi_out (4-byte aligned). */
template <int B>
__device__ __forceinline__ void store_group_at(uint8_t *rsi_out, uint32_t s_local, const uint32_t *s, uint32_t bsel)
{
    uint32_t *o = reinterpret_cast<uint32_t *>(rsi_out);
    if (B == 4) {
        o[s_local] = __byte_perm(s[0], 0, bsel);                   /* bsel: 0x3210 as is, 0x0123 byte-swapped */
    } else if (B == 2) {
        o[s_local >> 1] = __byte_perm(s[0], s[1], bsel);           /* 0x5410 / 0x4501 */
    } else if (B == 1) {
        o[s_local >> 2] = __byte_perm(__byte_perm(s[0], s[1], 0x0040), __byte_perm(s[2], s[3], 0x0040), 0x5410);
    } else {
        uint32_t a = s[0] & 0xFFFFFFu, b = s[1] & 0xFFFFFFu, cc = s[2] & 0xFFFFFFu, d = s[3] & 0xFFFFFFu;
        if (bsel == 0x0123u) {
            a = __byte_perm(a, 0, 0x4012); b = __byte_perm(b, 0, 0x4012);
            cc = __byte_perm(cc, 0, 0x4012); d = __byte_perm(d, 0, 0x4012);
        }
        o += (s_local >> 2) * 3u;
        o[0] = a | (b << 24);
        o[1] = (b >> 8) | (cc << 16);
        o[2] = (cc >> 16) | (d << 8);
    }
}

/* The warp's rows (32 rows of `steps` x 32*SPG samples, one more word between rows) -> the RSI's samples.
 * Row r gets the start value of lane r added (0 after the exact walk), then back to the n-bit pattern and,
 * for signed data, sign extension to the storage width (decode.c:78-84, :131).  STEPS > 0: steps known at
 * compile time (rows are short -- two warp steps on the README workload -- so loop overhead matters). */
template <int B, bool SXT, int STEPS>
__device__ __forceinline__ void store_rows(uint8_t *rsi_out, uint32_t sb, uint32_t o, uint32_t nrows, uint32_t steps_rt,
                                           uint32_t usadd, uint32_t xorv, uint32_t sxsh, uint32_t msb)
{
    constexpr int SPG = (B == 4) ? 1 : ((B == 2) ? 2 : 4);
    const uint32_t bsel = (B == 2) ? (msb ? 0x4501u : 0x5410u) : (msb ? 0x0123u : 0x3210u);
    const uint32_t steps = STEPS ? (uint32_t)STEPS : steps_rt;
    for (uint32_t rowi = 0; rowi < nrows; rowi++) {
        const uint32_t ua = __shfl_sync(FULL, usadd, rowi);
#pragma unroll
        for (uint32_t q = 0; q < steps; q++) {
            uint32_t sv[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < SPG; j++) {
                uint32_t x;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x) : "r"(sb + q * 128u * SPG + 4u * j));
                x = (x + ua) ^ xorv;
                if (SXT) x = (uint32_t)((int32_t)(x << sxsh) >> sxsh);
                sv[j] = x;
            }
            store_group_at<B>(rsi_out, o + q * 32u * SPG, sv, bsel);
        }
        sb += steps * 128u * SPG + 4u;
        o += steps * 32u * SPG;
    }
}

template <int B, bool SXT>
__device__ __forceinline__ void store_rows_any(uint8_t *rsi_out, uint32_t sb, uint32_t o, uint32_t nrows, uint32_t steps,
                                               uint32_t usadd, uint32_t xorv, uint32_t sxsh, uint32_t msb)
{
    if (steps == 2u)      store_rows<B, SXT, 2>(rsi_out, sb, o, nrows, steps, usadd, xorv, sxsh, msb);
    else if (steps == 1u) store_rows<B, SXT, 1>(rsi_out, sb, o, nrows, steps, usadd, xorv, sxsh, msb);
    else if (steps == 4u) store_rows<B, SXT, 4>(rsi_out, sb, o, nrows, steps, usadd, xorv, sxsh, msb);
    else                  store_rows<B, SXT, 0>(rsi_out, sb, o, nrows, steps, usadd, xorv, sxsh, msb);
}

template <int JT, int B>
__global__ void __launch_bounds__(DW_MAX_WARPS * 32)
aec_decode_warp_kernel(const AecDecArgs a)
{
    const AecCfg &c = a.cfg;
    const uint32_t J = JT ? (uint32_t)JT : c.J;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t G = a.grp_G;
    const uint32_t GJ = G * J;                        /* samples per lane */
    const uint32_t stride = GJ | 1u;                  /* odd row stride: conflict free */
    extern __shared__ uint32_t rows[];
    uint32_t *row = rows + ((size_t)warp * 32u + lane) * stride;

    const uint64_t r = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (r >= a.nrsi) return;
    /* samples this RSI has to deliver (the last RSI may be short, the caller may ask for less) */
    const uint64_t startS = r * (uint64_t)c.R;
    const uint64_t want = a.out_samples > startS ? a.out_samples - startS : 0;
    const uint32_t limit = want < c.R ? (uint32_t)want : c.R;
    const uint32_t nblk_rsi = (limit + J - 1) / J;         /* blocks that have to be decoded */
    const bool whole = limit > 0;
    uint32_t bad = 0u;

    const uint32_t b0 = lane * G;
    const uint32_t nblk = b0 < nblk_rsi ? (nblk_rsi - b0 < G ? nblk_rsi - b0 : G) : 0u;
    LaneAcc la; la.pfx = 0; la.big = 0; la.pmin = 0; la.pmax = 0;   /* running sum of my steps, how far they reach */
    uint32_t uref = 0;
    uint64_t endpos = 0, startpos = 0;
    uint32_t lead0 = 0, zl_end = 0;
    if (whole && nblk) {
        const uint64_t e = a.grp_index[r * 32ull + lane];
        uint32_t lead = (uint32_t)(e >> 56);
        lead0 = lead;
        startpos = e & 0x00FFFFFFFFFFFFFFull;
        const uint32_t nwords = (uint32_t)((a.in_bytes + 3) >> 2);
        if (startpos > a.in_bytes * 8ull) { bad = 1u; startpos = 0; }
        Rd64 rd;
        rd.init(a.in_words, nwords, startpos);
        uint32_t zero_left = 0;
        for (uint32_t q = 0; q < nblk; q++) {
            uint32_t *rw = row + q * J;
            if (lead) {
                lead--;
                for (uint32_t i = 0; i < J; i++) rw[i] = la.pfx;
            } else {
                warp_decode_block<JT>(c, rd, b0 + q, rw, zero_left, &bad, la);
            }
        }
        endpos = rd.pos();
        zl_end = zero_left;
        if (endpos > a.in_bytes * 8ull) bad = 1u;
        if (lane == 0 && c.pp) {
            uref = (row[0] ^ (c.sext ? (1u << (c.n - 1)) : 0u)) & c.mask;
            row[0] = 0;                                /* the reference sample itself: start value + 0 */
        }
    }
    const uint32_t sum = la.pfx;
    /* index sanity: the next lane's group must start where mine ended, unless it inherits a zero run */
    {
        unsigned long long nstart = __shfl_down_sync(FULL, (unsigned long long)startpos, 1);
        uint32_t nlead = __shfl_down_sync(FULL, lead0, 1);
        uint32_t nn = __shfl_down_sync(FULL, nblk, 1);
        if (whole && nblk && lane < 31 && nn && nlead == 0 && zl_end == 0 && nstart != endpos) bad = 1u;
    }
    bad = __any_sync(FULL, bad) ? 1u : 0u;

    const uint32_t sflip = c.sext ? (1u << (c.n - 1)) : 0u;
    const uint32_t n_s = nblk * J;
    uint32_t usadd = 0;                                /* what the store pass adds to my row */
    bool exact_walk = !bad;
    if (!bad && c.pp) {
        /* ---- unit-delay predictor undone in parallel.  A lane's start value is the reference
         * sample plus the steps of all lanes before it, provided nothing clipped.  A lane that
         * starts at us, stays inside [us+pmin, us+pmax] and never steps by more than big cannot
         * clip when us+pmin >= big and us+pmax <= M-big (steps below 2^20 and rows of at most a
         * few thousand samples keep the running sums far from wrapping); if that holds for every
         * lane, the start values are exact by induction over the lanes and the rows (running sums)
         * only need their start value added when they are stored. ---- */
        uref = __shfl_sync(FULL, uref, 0);
        uint32_t inc = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t o = __shfl_up_sync(FULL, inc, off);
            if (lane >= (uint32_t)off) inc += o;
        }
        const uint32_t us0 = uref + (inc - sum);
        const long long lo = (long long)us0 + la.pmin, hi = (long long)us0 + la.pmax;
        const bool calm = n_s == 0 ||
                          (la.big < (1u << 21) && us0 <= c.mask && lo >= (long long)la.big &&
                           hi <= (long long)c.mask - (long long)la.big);
        if (__all_sync(FULL, calm)) { usadd = us0; exact_walk = false; }
    }
    if (exact_walk && !c.pp) {
        /* no predictor: the rows have to hold the values themselves; recover them from the running sums */
        uint32_t prev = 0;
        for (uint32_t i = 0; i < n_s; i++) {
            const uint32_t cur = row[i];
            const uint32_t dl = cur - prev;
            prev = cur;
            row[i] = (dl << 1) ^ (uint32_t)((int32_t)dl >> 31);
        }
    }
    if (exact_walk && c.pp) {
        /* ---- some sample may clip: every lane walks its samples with the
         * exact inverse map from a speculated start value.  The first walk is optimistic (start
         * values from the prefix sum of the deltas, exact whenever no sample clips) and writes
         * the normalised samples in place; if a lane did not start where its predecessor ended,
         * the mapped values are recovered (the map is a bijection) and the start values are
         * re-derived from the lanes' results until they agree. ---- */
        uint32_t inc = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t o = __shfl_up_sync(FULL, inc, off);
            if (lane >= (uint32_t)off) inc += o;
        }
        /* speculated value before my first sample; kept inside [0, M] so that every walk stays in the
         * mapper's domain and the mapped values can be recovered exactly after a wrong guess */
        uint32_t us = project_start(uref, inc - sum, c.mask);
        const uint32_t i_first = (lane == 0) ? 1u : 0u;
        if (lane == 0) us = uref;
        uint32_t u = us, clip = 0;
        if (lane == 0 && n_s) row[0] = uref;
        {
            /* the rows still hold running sums: the mapped value is the zigzag of their difference */
            uint32_t pprev = 0;
#pragma unroll 4
            for (uint32_t i = i_first; i < n_s; i++) {
                const uint32_t cur = row[i];
                const uint32_t dl = cur - pprev;
                pprev = cur;
                u = unmap_step(u, (dl << 1) ^ (uint32_t)((int32_t)dl >> 31), c.mask, clip);
                row[i] = u;
            }
        }
        uint32_t uprev = __shfl_up_sync(FULL, u, 1);
        bool ok = (lane == 0) || (n_s == 0) || (uprev == us);
        if (!__all_sync(FULL, ok)) {
            /* recover the mapped values from the (wrongly started) samples */
            uint32_t pv = us;
            for (uint32_t i = i_first; i < n_s; i++) { uint32_t cur = row[i]; row[i] = aec_map_delta(pv, cur, c.mask); pv = cur; }
            for (int iter = 0; iter < 36; iter++) {
                /* Which lanes start where their predecessor ended?  The leftmost lane that does not
                 * is repaired exactly every round (everything before it is final), so the loop ends
                 * after at most 32 rounds; the two update rules below only differ in how boldly they
                 * also move the lanes further right (measured on imagery-like, clipped-walk and noise
                 * data: DESIGN.md 4.2). */
                uprev = __shfl_up_sync(FULL, u, 1);
                const bool cons = (lane == 0) || (n_s == 0) || (uprev == us);
                const uint32_t incons = __ballot_sync(FULL, !cons);
                if (incons == 0) break;
                if (iter == 35) { bad = 1u; break; }
                if (__popc(incons) > 4) {
                    /* many lanes off (noise-like data: nearly every lane clips and forgets its start):
                     * chain the lanes' results, absolute after a lane that clipped, relative otherwise */
                    uint32_t val = clip ? u : (u - us);
                    uint32_t rst = clip;
                    if (lane == 0) { val = u; rst = 1u; }
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        uint32_t ov = __shfl_up_sync(FULL, val, off);
                        uint32_t orst = __shfl_up_sync(FULL, rst, off);
                        if (lane >= (uint32_t)off && !rst) { val += ov; rst = orst; }
                    }
                    uint32_t nus = __shfl_up_sync(FULL, val, 1);
                    if (lane > 0) us = nus & c.mask;
                } else {
                    /* few lanes off (a start value guessed wrong here and there): move a lane to its
                     * predecessor's end only if that predecessor itself started consistently, and carry
                     * the same correction through the clip-free lanes that follow it */
                    const bool cons_m1 = __shfl_up_sync(FULL, (int)cons, 1) != 0 || lane == 0;
                    const bool clip_m1 = __shfl_up_sync(FULL, clip, 1) != 0 && lane != 0;
                    const bool fix = !cons && cons_m1;
                    uint32_t val = fix ? (uprev - us) : 0u;
                    uint32_t stop = (fix || !(cons && !clip_m1)) ? 1u : 0u;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        uint32_t ov = __shfl_up_sync(FULL, val, off);
                        uint32_t ostop = __shfl_up_sync(FULL, stop, off);
                        if (lane >= (uint32_t)off && !stop) { val += ov; stop = ostop; }
                    }
                    us = (us + val) & c.mask;
                }
                u = us; clip = 0;
                for (uint32_t i = i_first; i < n_s; i++) u = unmap_step(u, row[i], c.mask, clip);
            }
            bad = __any_sync(FULL, bad) ? 1u : 0u;
            if (!bad) {
                u = us; clip = 0;
                for (uint32_t i = i_first; i < n_s; i++) { u = unmap_step(u, row[i], c.mask, clip); row[i] = u; }
            }
        }
    }
    if (bad) {
        if (lane == 0) { uint32_t slot = atomicAdd(a.rsi_list_count, 1u); a.rsi_list[slot] = (uint32_t)r; }
        return;
    }
    __syncwarp();
    /* ---- cooperative store of the RSI's R samples (contiguous in the output) ---- */
    uint32_t *wrows = rows + (size_t)warp * 32u * stride;
    const bool sxt = c.sext && c.n < 32;
    if (!whole) return;
    /* row r of the warp belongs to lane r; its samples are row value + usadd of that lane (the start
     * value when the rows hold running sums, 0 after the exact walk), then back to the n-bit pattern
     * and sign-extended to the storage width (decode.c:78-84, :131) */
    const uint32_t xorv = c.pp ? sflip : 0u;
    const uint32_t sxsh = (c.pp && sxt) ? 32u - c.n : 0u;
    constexpr int SPG = (B == 4) ? 1 : ((B == 2) ? 2 : 4);      /* samples per 32-bit store group */
    if (JT != 0 && a.out_aligned && limit == c.R && (GJ % (32u * SPG)) == 0 && (c.rsi % G) == 0) {
        /* every lane owns G whole blocks: GJ / (32*SPG) warp steps per row, the row's lane is warp-uniform */
        const uint32_t nrows = c.R / GJ, steps = GJ / (32u * SPG);
        uint8_t *const rsi_out = a.out + startS * B;                          /* 4-byte aligned: R*B is a multiple of 4 here */
        const uint32_t sb0 = (uint32_t)__cvta_generic_to_shared(wrows) + lane * SPG * 4u;
        if (sxsh) store_rows_any<B, true>(rsi_out, sb0, lane * SPG, nrows, steps, usadd, xorv, sxsh, c.msb);
        else      store_rows_any<B, false>(rsi_out, sb0, lane * SPG, nrows, steps, usadd, xorv, 0u, c.msb);
    } else if (JT != 0 && a.out_aligned && (GJ % 4u) == 0 && limit == c.R) {
        const uint32_t ngroups = c.R / SPG;
        for (uint32_t g0 = 0; g0 < ngroups; g0 += 32) {
            const uint32_t g = g0 + lane;
            const bool ok = g < ngroups;
            const uint32_t s0 = ok ? g * SPG : 0u;
            const uint32_t rowi = s0 / GJ, col = s0 - rowi * GJ;
            const uint32_t ua = __shfl_sync(FULL, usadd, rowi);
            if (!ok) continue;
            const uint32_t *src = wrows + (size_t)rowi * stride + col;
            uint32_t sv[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < SPG; j++) {
                uint32_t x = (src[j] + ua) ^ xorv;
                if (sxsh) x = (uint32_t)((int32_t)(x << sxsh) >> sxsh);
                sv[j] = x;
            }
            store_group<B>(a.out, startS + s0, sv, c.msb);
        }
    } else {
        for (uint32_t sb = 0; sb < limit; sb += 32) {
            const uint32_t s0 = sb + lane;
            const bool ok = s0 < limit;
            const uint32_t rowi = ok ? s0 / GJ : 0u, col = ok ? s0 - rowi * GJ : 0u;
            const uint32_t ua = __shfl_sync(FULL, usadd, rowi);
            if (!ok) continue;
            uint32_t x = (wrows[(size_t)rowi * stride + col] + ua) ^ xorv;
            if (sxsh) x = (uint32_t)((int32_t)(x << sxsh) >> sxsh);
            aec_store_sample(a.out + (startS + s0) * c.B, x, c.B, c.msb);
        }
    }
}

/* Group index of RSIs with known start offsets: lane-per-RSI skim. */
__global__ void aec_build_group_index_kernel(const AecDecArgs a, uint64_t *grp_index, int only_missing)
{
    const AecCfg &c = a.cfg;
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.nrsi) return;
    if (only_missing && grp_index[r * 32ull] != 0xFFFFFFFFFFFFFFFFull) return;   /* SK_GRP_MISSING: see aec_skim_core.cuh */
    const uint32_t G = a.grp_G;
    BitRd br;
    br.init(a.in_words, (a.in_bytes + 3) >> 2, a.in_bytes * 8ull);
    RsiDec st; st.pos = a.rsi_offsets[r]; st.zero_left = 0; st.status = DEC_OK;
    for (uint32_t b = 0; b < c.rsi; b++) {
        if (b % G == 0) grp_index[r * 32ull + b / G] = ((uint64_t)st.zero_left << 56) | (st.pos & 0x00FFFFFFFFFFFFFFull);
        if (!aec_skim_block(c, br, st, b)) {
            /* truncated or corrupt: poison the remaining groups so the fast kernel hands the RSI over */
            for (uint32_t g = b / G + 1; g * G < c.rsi; g++) grp_index[r * 32ull + g] = 0x00FFFFFFFFFFFFFFull;
            break;
        }
    }
}

__global__ void aec_scan_offsets_kernel(const AecCfg c, const uint32_t *in_words, uint64_t in_bytes,
                                        uint64_t start_bit, uint64_t *offsets, uint64_t max_rsi,
                                        uint64_t *result)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    BitRd br;
    br.init(in_words, (in_bytes + 3) >> 2, in_bytes * 8ull);
    RsiDec st; st.pos = start_bit; st.zero_left = 0; st.status = DEC_OK;
    uint64_t found = 0;
    for (uint64_t r = 0; r < max_rsi; r++) {
        if (c.pad) st.pos = (st.pos + 7ull) & ~7ull;
        uint64_t start = st.pos;
        st.zero_left = 0;
        if (start >= br.nbits) break;
        offsets[found++] = start;          /* even a truncated RSI may still deliver leading samples */
        for (uint32_t b = 0; b < c.rsi; b++)
            if (!aec_skim_block(c, br, st, b)) break;
        if (st.status != DEC_OK) break;
    }
    result[0] = found;
    result[1] = (st.status == DEC_ERROR) ? 1ull : 0ull;
    result[2] = st.pos;
}

template <int JT, int B>
cudaError_t launch_dec(const AecDecArgs &a, cudaStream_t st)
{
    auto kern = aec_decode_kernel<JT, B>;
    uint32_t J = JT ? (uint32_t)JT : a.cfg.J;
    uint32_t smem = DEC_WARPS * 32u * (J + 1u) * 4u;
    static AecSmemOptIn optin;
    {
        cudaError_t e = optin.ensure(kern, smem, 48 * 1024);
        if (e != cudaSuccess) return e;
    }
    uint64_t nwarps = (a.nrsi + 31) / 32;
    uint64_t grid = (nwarps + DEC_WARPS - 1) / DEC_WARPS;
    if (grid == 0) return cudaSuccess;
    kern<<<(unsigned)grid, DEC_WARPS * 32, smem, st>>>(a);
    return cudaGetLastError();
}

template <int JT>
cudaError_t launch_dec_j(const AecDecArgs &a, cudaStream_t st)
{
    switch (a.cfg.B) {
    case 1: return launch_dec<JT, 1>(a, st);
    case 2: return launch_dec<JT, 2>(a, st);
    case 3: return launch_dec<JT, 3>(a, st);
    default: return launch_dec<JT, 4>(a, st);
    }
}

} // namespace

cudaError_t aec_decode_launch(const AecDecArgs &a, int num_sms, cudaStream_t st)
{
    (void)num_sms;
    switch (a.cfg.J) {
    case 8:  return launch_dec_j<8>(a, st);
    case 16: return launch_dec_j<16>(a, st);
    case 32: return launch_dec_j<32>(a, st);
    case 64: return launch_dec_j<64>(a, st);
    default: return launch_dec_j<0>(a, st);
    }
}

namespace {

template <int JT, int B>
cudaError_t launch_decw(const AecDecArgs &a, cudaStream_t st)
{
    auto kern = aec_decode_warp_kernel<JT, B>;
    uint32_t warps = aec_decode_warp_warps(a.cfg);
    if (warps == 0) return cudaErrorInvalidConfiguration;
    uint32_t J = JT ? (uint32_t)JT : a.cfg.J;
    uint32_t stride = (a.grp_G * J) | 1u;
    uint32_t smem = warps * 32u * stride * 4u;
    static AecSmemOptIn optin;
    {
        cudaError_t e = optin.ensure(kern, smem, 48 * 1024);
        if (e != cudaSuccess) return e;
    }
    uint64_t grid = (a.nrsi + warps - 1) / warps;
    if (grid == 0) return cudaSuccess;
    kern<<<(unsigned)grid, warps * 32, smem, st>>>(a);
    return cudaGetLastError();
}

template <int JT>
cudaError_t launch_decw_j(const AecDecArgs &a, cudaStream_t st)
{
    switch (a.cfg.B) {
    case 1: return launch_decw<JT, 1>(a, st);
    case 2: return launch_decw<JT, 2>(a, st);
    case 3: return launch_decw<JT, 3>(a, st);
    default: return launch_decw<JT, 4>(a, st);
    }
}

} // namespace

uint32_t aec_decode_group_blocks(const AecCfg &c) { return (c.rsi + 31u) / 32u; }

/* warps per CTA of the warp-per-RSI kernel (0: an RSI's rows do not fit shared memory -> careful kernel only).
 * The kernel is bound by instruction issue, so what matters is how many warps an SM holds: shared
 * memory is the limit (228 KiB per SM, 1 KiB of it reserved per resident CTA, 227 KiB at most per CTA). */
uint32_t aec_decode_warp_warps(const AecCfg &c)
{
    const uint64_t stride = (aec_decode_group_blocks(c) * c.J) | 1u;
    const uint64_t per_warp = 32u * stride * 4u;
    if (per_warp > 96u * 1024u) return 0;
    uint32_t best_w = 0, best_total = 0;
    for (uint32_t w = 1; w <= (uint32_t)DW_MAX_WARPS; w++) {
        const uint64_t cta = w * per_warp;
        if (cta > 227u * 1024u) break;
        uint64_t ctas = (228u * 1024u) / (cta + 1024u);
        if (ctas > 32u) ctas = 32u;
        uint64_t total = ctas * w;
        if (total > 64u) total = 64u;
        /* ties go to the smaller CTA: less of the SM idles while the last CTAs of the grid finish */
        if (total > best_total) { best_total = (uint32_t)total; best_w = w; }
    }
    return best_w;
}

cudaError_t aec_decode_warp_launch(const AecDecArgs &a, int num_sms, cudaStream_t st)
{
    (void)num_sms;
    switch (a.cfg.J) {
    case 8:  return launch_decw_j<8>(a, st);
    case 16: return launch_decw_j<16>(a, st);
    case 32: return launch_decw_j<32>(a, st);
    case 64: return launch_decw_j<64>(a, st);
    default: return launch_decw_j<0>(a, st);
    }
}

cudaError_t aec_build_group_index_launch(const AecDecArgs &a, uint64_t *grp_index, cudaStream_t st, int only_missing)
{
    if (a.nrsi == 0) return cudaSuccess;
    aec_build_group_index_kernel<<<(unsigned)((a.nrsi + 127) / 128), 128, 0, st>>>(a, grp_index, only_missing);
    return cudaGetLastError();
}

cudaError_t aec_scan_offsets_launch(const AecCfg &c, const uint32_t *in_words, uint64_t in_bytes,
                                    uint64_t start_bit, uint64_t *offsets, uint64_t max_rsi,
                                    uint64_t *result, cudaStream_t st)
{
    aec_scan_offsets_kernel<<<1, 32, 0, st>>>(c, in_words, in_bytes, start_bit, offsets, max_rsi, result);
    return cudaGetLastError();
}
