/*
 * aec_encode.cu -- single-pass CCSDS 121.0-B-2 encoder for sm_100a.
 *
 * Replaces the hot loop of the reference's encoder state machine
 * (/root/reference/src/encode.c:614-754 m_get_block / m_check_zero_block /
 * m_select_code_option / m_encode_*, :235-311 preprocess_*, :313-434 option
 * costs, :61-233 bit emission, src/encode_accessors.c) with one persistent
 * kernel:
 *
 *   thread  = one block of J samples (loaded with 16-byte vector loads,
 *             mapped and costed in registers),
 *   warp    = ballots give the zero-run structure of a 64-block segment,
 *             shuffles give the exclusive scan of CDS bit lengths and of the
 *             k clamp chain,
 *   CTA     = one tile of TB consecutive block slots; tiles are handed out in
 *             order by an atomic ticket, publish their aggregate
 *             (bits, clamp pair) in a 64-bit descriptor and obtain their
 *             absolute bit offset / incoming k with a decoupled look-back,
 *   output  = each thread packs its CDS into a shared-memory staging area
 *             laid out at the tile's final bit phase; the tile then streams
 *             whole 32-bit words to global memory.  Words shared between two
 *             tiles are resolved afterwards by a tiny fix-up kernel, so the
 *             output needs neither atomics nor pre-zeroing.
 *
 * Input is read exactly once and output written exactly once.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "aec_core.cuh"
#include "aec_device.h"

namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;

/* ---- descriptor encoding (one 64-bit word, written/read with single
 * volatile accesses) ------------------------------------------------------ */
constexpr uint64_t ST_AGG = 1, ST_PREFIX = 2;

__device__ __forceinline__ uint64_t desc_pack_agg(const PosFn &f, uint32_t kp)
{
    return ST_AGG | ((uint64_t)(aec_klo(kp) & 0x1Fu) << 2) | ((uint64_t)(aec_khi(kp) & 0x1Fu) << 7) |
           ((uint64_t)(f.has_end & 1u) << 12) | ((f.a & 0x1FFFFFull) << 13) | ((f.rest & 0x1FFFFFull) << 34);
}
__device__ __forceinline__ uint64_t desc_pack_prefix(uint64_t bits, uint32_t k)
{
    return ST_PREFIX | ((uint64_t)(k & 0x1Fu) << 2) | (bits << 12);
}
__device__ __forceinline__ void desc_unpack(uint64_t d, PosFn &f, uint32_t &kp)
{
    if ((d & 3) == ST_PREFIX) {
        uint32_t k = (uint32_t)(d >> 2) & 0x1Fu;
        kp = aec_kpair(k, k);
        f.has_end = 0; f.a = d >> 12; f.rest = 0;
    } else {
        kp = aec_kpair((uint32_t)(d >> 2) & 0x1Fu, (uint32_t)(d >> 7) & 0x1Fu);
        f.has_end = (uint32_t)(d >> 12) & 1u;
        f.a = (d >> 13) & 0x1FFFFFull;
        f.rest = (d >> 34) & 0x1FFFFFull;
    }
}

/* descriptor accesses: relaxed, gpu scope (L2 is the coherence point) */
__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t *p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ PosFn shfl_posfn(const PosFn &v, int src)
{
    PosFn r;
    r.has_end = __shfl_sync(FULL, v.has_end, src);
    r.a = __shfl_sync(FULL, (unsigned long long)v.a, src);
    r.rest = __shfl_sync(FULL, (unsigned long long)v.rest, src);
    return r;
}

/* ---- sample loading ------------------------------------------------------ */

/* Sample i of a block whose bytes sit in the word array w (little-endian
 * words as loaded). Compile-time i after unrolling -> a couple of PRMT/SHF. */
template <int B>
__device__ __forceinline__ uint32_t extract_sample(const uint32_t *w, int i, uint32_t msb)
{
    if (B == 4) return msb ? __byte_perm(w[i], 0, 0x0123) : w[i];
    if (B == 2) {
        uint32_t x = w[i >> 1];
        if (i & 1) return msb ? __byte_perm(x, 0, 0x4423) : (x >> 16);
        return msb ? __byte_perm(x, 0, 0x4401) : (x & 0xFFFFu);
    }
    if (B == 1) return (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
    /* B == 3 */
    int o = i * 3;
    uint32_t lo = w[o >> 2];
    uint32_t hi = ((o & 3) > 1) ? w[(o >> 2) + 1] : 0u;
    uint32_t v = __funnelshift_r(lo, hi, 8 * (o & 3)) & 0xFFFFFFu;
    return msb ? __byte_perm(v, 0, 0x4012) : v;
}

/* Load the J raw samples of one block.  fast: the whole block is inside the
 * input and the base pointer is 16-byte aligned. */
template <int JT, int B>
__device__ __forceinline__ void load_block(const AecCfg &c, const uint8_t *in, uint64_t first,
                                           uint64_t nsamples, bool fast, uint32_t *x)
{
    const uint32_t J = JT ? (uint32_t)JT : c.J;
    if (JT != 0 && fast) {
        constexpr int BB = (JT ? JT : 2) * B;               /* bytes per block */
        constexpr int VW = (BB % 16 == 0) ? 16 : ((BB % 8 == 0) ? 8 : ((BB % 4 == 0) ? 4 : 2));
        constexpr int NW = (BB + 3) / 4;
        uint32_t w[NW + 1];
        const uint8_t *p = in + first * B;
        if (VW == 16) {
#pragma unroll
            for (int j = 0; j < BB / 16; j++) {
                uint4 q = __ldg(reinterpret_cast<const uint4 *>(p) + j);
                w[4 * j] = q.x; w[4 * j + 1] = q.y; w[4 * j + 2] = q.z; w[4 * j + 3] = q.w;
            }
        } else if (VW == 8) {
#pragma unroll
            for (int j = 0; j < BB / 8; j++) {
                uint2 q = __ldg(reinterpret_cast<const uint2 *>(p) + j);
                w[2 * j] = q.x; w[2 * j + 1] = q.y;
            }
        } else if (VW == 4) {
#pragma unroll
            for (int j = 0; j < BB / 4; j++) w[j] = __ldg(reinterpret_cast<const uint32_t *>(p) + j);
        } else {
#pragma unroll
            for (int j = 0; j < NW; j++) {
                uint32_t a = __ldg(reinterpret_cast<const uint16_t *>(p) + 2 * j);
                uint32_t b = (4 * j + 2 < BB) ? __ldg(reinterpret_cast<const uint16_t *>(p) + 2 * j + 1) : 0u;
                w[j] = a | (b << 16);
            }
        }
        w[NW] = 0;
#pragma unroll
        for (int i = 0; i < JT; i++) x[i] = extract_sample<B>(w, i, c.msb);
    } else {
        /* generic: bytewise, index clamped to the last sample (encode.c:681-684) */
#pragma unroll
        for (uint32_t i = 0; i < (JT ? (uint32_t)JT : J); i++) {
            uint64_t idx = first + i;
            if (idx >= nsamples) idx = nsamples - 1;
            x[i] = aec_load_sample(in + idx * c.B, c.B, c.msb);
        }
    }
}

/* Every spin is bounded.  The kernel is launched cooperatively, so the scanner CTA and all worker CTAs
 * are resident together and a poll is normally answered within microseconds.  The scanner (its own
 * function, off the workers' register budget) backs off with nanosleep when a wait gets long anyway
 * (debugger, sanitizer, a preempted context); the workers, which sit exactly at 64 registers, only
 * count.  A wait of the order of a minute -- a broken protocol -- aborts the launch (the host sees a
 * launch failure) instead of hanging the device. */
constexpr uint32_t SPIN_FAST = 1u << 16;     /* scanner: polls before backing off */
constexpr uint32_t SPIN_LIMIT = 1u << 26;    /* scanner: then this many sleeps of ~1 us */
constexpr uint32_t SPIN_WORKER = 1u << 28;   /* workers: polls of an L2 word, ~0.3 us each */
__device__ __forceinline__ void spin_guard_scanner(uint32_t &n)
{
    if (++n > SPIN_FAST) {
        __nanosleep(1000);
        if (n > SPIN_FAST + SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ void spin_guard(uint32_t &n)
{
    if (++n > SPIN_WORKER) __trap();
}

/* ---- dedicated scanner ------------------------------------------------------
 * CTA 0 turns the per-tile aggregates into exclusive prefixes (absolute bit
 * offset + incoming k of every tile).  Worker CTAs never look back: they
 * publish their aggregate, pack their tile, and read their prefix when they
 * are ready to write.  Warp w of the scanner owns the batches w, w+NW, ... of
 * 32 x TPL consecutive tiles: every lane loads TPL aggregates and composes them,
 * the warp scans the lane totals, lane 31 takes the running (position, k) carry
 * from the previous batch through shared memory and passes the carry behind this
 * batch on at once (it only needs the batch totals for that); then the lanes
 * write their tiles' prefixes.  The serial part of the chain is that hand-over
 * only, once per 32 x TPL tiles: with one hand-over per 32 tiles it was what
 * bounded the whole encoder (0.41 us per batch against 0.42 us in which the
 * workers produce 32 tiles: profiles/r2_summary.md). */
template <int NW, int TPL>
__device__ __noinline__ void aec_encode_scanner(const AecEncArgs &a, uint64_t *s_carry_pos, uint32_t *s_carry_k,
                                                volatile uint32_t *s_done)
{
    const AecCfg &c = a.cfg;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t kident = aec_kpair(0, c.kmax);
    const uint64_t ntiles = a.ntiles;
    constexpr uint64_t PER = 32ull * TPL;
    const uint64_t nbatch = (ntiles + PER - 1) / PER;
    if (threadIdx.x == 0) { s_carry_pos[0] = a.seed_bits; s_carry_k[0] = a.dyn ? (uint32_t)a.dyn[0] : a.seed_k; *s_done = 0; }
    __syncthreads();
    for (uint64_t b = warp; b < nbatch; b += NW) {
        const uint64_t t0 = b * PER + (uint64_t)lane * TPL;
        uint64_t dv[TPL];
#pragma unroll
        for (int i = 0; i < TPL; i++)                             /* all loads in flight together */
            dv[i] = (t0 + i < ntiles) ? ld_volatile_u64(&a.desc[t0 + i])
                                      : (ST_AGG | ((uint64_t)c.kmax << 7));   /* identity beyond the last tile */
#pragma unroll
        for (int i = 0; i < TPL; i++) {
            uint32_t spins = 0;
            while ((dv[i] & 3) == 0) { dv[i] = ld_volatile_u64(&a.desc[t0 + i]); spin_guard_scanner(spins); }
        }
        __syncwarp();
        PosFn f[TPL]; uint32_t kj[TPL];
#pragma unroll
        for (int i = 0; i < TPL; i++) desc_unpack(dv[i], f[i], kj[i]);
        /* the lane's own tiles composed, then the inclusive scan over the lanes (lower lane = earlier tiles) */
        PosFn fi = f[0]; uint32_t ki = kj[0];
#pragma unroll
        for (int i = 1; i < TPL; i++) {
            if (c.pad) fi = aec_pcompose(fi, f[i]); else fi.a += f[i].a;
            ki = aec_kcompose(ki, kj[i]);
        }
        if (c.pad) {
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                PosFn o = shfl_posfn(fi, (int)lane - off);
                if (lane >= (uint32_t)off) fi = aec_pcompose(o, fi);
            }
        } else {
            uint32_t li = (uint32_t)fi.a;                         /* 32 x TPL tiles of < 2^21 bits fit 32 bits */
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t o = __shfl_up_sync(FULL, li, off);
                if (lane >= (uint32_t)off) li += o;
            }
            fi.a = li;
        }
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t ok = __shfl_up_sync(FULL, ki, off);
            if (lane >= (uint32_t)off) ki = aec_kcompose(ok, ki);
        }
        PosFn fe = shfl_posfn(fi, (int)lane - 1);
        uint32_t ke = __shfl_up_sync(FULL, ki, 1);
        if (lane == 0) { fe.has_end = 0; fe.a = 0; fe.rest = 0; ke = kident; }
        /* carry of everything before this batch.  Only lane 31 touches the hand-over slots and passes
         * the values on by shuffle: the other lanes never read shared memory here, so nothing depends
         * on the lanes of a warp staying converged between the poll, the read and the publication of
         * the next carry (an earlier version let every lane read the slot: the batch after next
         * could overwrite it before a lane that ran late had read it).  Lane 31 holds the batch
         * totals (its inclusive scan values), so the next carry leaves before anything else is done. */
        uint64_t P = 0;
        uint32_t kc = 0;
        if (lane == 31) {   /* hand-over spin: a sleep quantum here would serialise the chain */
            /* flag protocol between the lanes 31 of consecutive batches: the slots are written before the
             * release store of the batch counter and read after the acquire load that saw it (a slot is
             * reused two batches later, i.e. after its reader has published its own counter value).
             * compute-sanitizer's racecheck, which only knows barriers, reports these three accesses
             * (profiles/r2_sanitize.md); it reports nothing else in the library. */
            const uint32_t done_addr = (uint32_t)__cvta_generic_to_shared(const_cast<uint32_t *>(s_done));
            uint32_t spins = 0, seen;
            for (;;) {
                asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(seen) : "r"(done_addr) : "memory");
                if (seen == (uint32_t)b) break;
                spin_guard_scanner(spins);
            }
            P = *reinterpret_cast<volatile uint64_t *>(&s_carry_pos[b & 1]);
            kc = *reinterpret_cast<volatile uint32_t *>(&s_carry_k[b & 1]);
            const uint64_t bend = aec_papply(fi, P);
            const uint32_t bk = aec_kapply(kc, ki);
            *reinterpret_cast<volatile uint64_t *>(&s_carry_pos[(b + 1) & 1]) = bend;
            *reinterpret_cast<volatile uint32_t *>(&s_carry_k[(b + 1) & 1]) = bk;
            asm volatile("st.release.cta.shared.u32 [%0], %1;" :: "r"(done_addr), "r"((uint32_t)(b + 1)) : "memory");
            if (b + 1 == nbatch && ntiles == a.ntiles_total && !a.dyn) {
                /* the last batch: identity tiles beyond the end keep the totals */
                a.result[0] = bend; a.result[1] = bk;
            }
        }
        P = __shfl_sync(FULL, (unsigned long long)P, 31);
        kc = __shfl_sync(FULL, kc, 31);
        uint64_t pos = aec_papply(fe, P);
        uint32_t k = aec_kapply(kc, ke);
#pragma unroll
        for (int i = 0; i < TPL; i++) {
            const uint64_t end = aec_papply(f[i], pos);
            if (t0 + i < ntiles) {
                st_volatile_u64(&a.pref[t0 + i], desc_pack_prefix(pos, k));
                a.tile_end[t0 + i] = end;
            }
            pos = end;
            k = aec_kapply(k, kj[i]);
        }
    }
}

/* ---- the kernel ----------------------------------------------------------- */

#ifndef AEC_SCAN_TPL
#define AEC_SCAN_TPL 1      /* tiles per scanner lane: one carry hand-over per 32 x AEC_SCAN_TPL tiles */
#endif

template <int JT>
struct TileCfg {
#ifndef AEC_TB16
#define AEC_TB16 256
#endif
    static constexpr int TB = (JT == 0 || JT == 64) ? 128 : ((JT == 16) ? AEC_TB16 : 256);
    static constexpr int NWARP = TB / 32;
    static constexpr int JMAX = JT ? JT : AEC_MAX_J;
    /* resident CTAs per SM the register allocation aims for */
#ifndef AEC_MINB16
#define AEC_MINB16 4
#endif
    static constexpr int MINB = (JT == 8 || JT == 16) ? AEC_MINB16 : ((JT == 32) ? 3 : ((JT == 64) ? 4 : 2));
};

__device__ __forceinline__ void pair_barrier(uint32_t warp)
{
    /* the two warps of one 64-block zero-run segment */
    asm volatile("bar.sync %0, 64;" :: "r"(1u + (warp >> 1)) : "memory");
}

/* One sample at index idx (the lane-0 look-back and the zero-run reference). */
template <int B>
__device__ __forceinline__ uint32_t load_one(const uint8_t *in, uint64_t idx, uint32_t msb, uint32_t aligned)
{
    if (B == 4 && aligned) {
        uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(in) + idx);
        return msb ? __byte_perm(v, 0, 0x0123) : v;
    }
    if (B == 2 && aligned) {
        uint32_t v = __ldg(reinterpret_cast<const uint16_t *>(in) + idx);
        return msb ? __byte_perm(v, 0, 0x4401) : v;
    }
    if (B == 1) return __ldg(in + idx);
    return aec_load_sample(in + idx * B, B, msb);
}

template <int JT, int B>
__global__ void __launch_bounds__(TileCfg<JT>::TB, TileCfg<JT>::MINB)
aec_encode_kernel(const __grid_constant__ AecEncArgs a)
{
    constexpr int TB = TileCfg<JT>::TB;
    constexpr int NWARP = TileCfg<JT>::NWARP;
    constexpr int JMAX = TileCfg<JT>::JMAX;
    const AecCfg &c = a.cfg;
    const uint32_t J = JT ? (uint32_t)JT : c.J;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t kident = aec_kpair(0, c.kmax);
    /* RSIs longer than a tile with RSI padding: a tile may start at an unknown
     * bit phase mod 8, so its bits can only be laid out once its prefix is known */
    const bool late = c.pad && a.RP > (uint32_t)TB;
    const bool want_index = a.rsi_offsets != nullptr || a.grp_index != nullptr;

    /* two staging areas of SW words: tile `it` packs into one while tile it-1 streams out of the
     * other; word 0 of an area stays zero (the word "before" the tile for the funnel shift).  Behind
     * them two areas of TB x 2 words: what each thread has to remember about its block of the tile
     * that still waits for its prefix (bit offset in the tile; CDS length, zero-run length, flags) */
    extern __shared__ uint4 staging_raw[];
    uint32_t *const staging_all = reinterpret_cast<uint32_t *>(staging_raw);
    const uint32_t SW = a.staging_words;
    uint2 *const pend_all = reinterpret_cast<uint2 *>(staging_all + 2u * SW);
    /* tickets and tile geometry live in a ring of three: the tile being streamed out, the tile being
     * packed and the tile claimed ahead */
    __shared__ uint32_t s_ticket[3];
    __shared__ unsigned long long s_rsi0[3]; /* first RSI of the claimed tile */
    __shared__ uint32_t s_b0[3];             /* first block slot of the tile inside that RSI (RP >= TB) */
    __shared__ uint32_t s_zb[NWARP + 1];
    __shared__ uint32_t s_wlen[NWARP];       /* per-warp sums (no-pad mode) */
    __shared__ uint32_t s_wend[NWARP], s_wa[NWARP], s_wrest[NWARP];   /* per-warp PosFn (pad mode) */
    __shared__ uint32_t s_wk[NWARP];
    __shared__ uint32_t s_side[TB];          /* one word per thread: the first completed word of its CDS (BitPack) */
    __shared__ uint32_t s_tend[2], s_ta[2], s_trest[2];   /* position map of the packed tile, per staging area */
    __shared__ unsigned long long s_base;    /* absolute bit offset of the tile being streamed out */
    /* copy-out plan of that tile, worked out once by thread 0: first output word; then sh, nsl, i_lo, i_al,
     * nvec, i_hi, nwhole, flags (1 head word shared, 2 tail word shared, 4 tile not empty) */
    __shared__ unsigned long long s_cp_w0;
    __shared__ uint32_t s_cp[8];
    __shared__ unsigned long long s_base_cur;/* late mode: absolute bit offset of the tile being packed */

    if (a.dyn) {
        /* a repair launch runs only when the number of tiles the device-side plan wants coded again
         * falls into this launch's window (nothing at all when the incoming k was 0) */
        const uint64_t want = a.dyn[1];
        if (want <= a.dyn_lo || want > a.dyn_hi) return;
    }
    if (blockIdx.x == 0) {                  /* CTA 0 is the scanner */
        __shared__ unsigned long long s_cpos[2];
        __shared__ uint32_t s_ck[2];
        __shared__ uint32_t s_sdone;
        aec_encode_scanner<NWARP, AEC_SCAN_TPL>(a, reinterpret_cast<uint64_t *>(s_cpos), s_ck, &s_sdone);
        return;
    }
    /* where a claimed tile sits (thread 0 only) */
    auto place = [&](uint32_t g, uint32_t t) {
        s_ticket[g] = t;
        if (a.tpr) { const uint32_t q = t / a.tpr; s_rsi0[g] = q; s_b0[g] = (t - q * a.tpr) * (uint32_t)TB; }
        else { s_rsi0[g] = (unsigned long long)t << a.tile_rsi_shift; s_b0[g] = 0; }
    };
    /* which block of the tile in ring entry g belongs to this thread */
    auto my_block = [&](uint32_t g, uint64_t &rsi_idx, uint32_t &b) {
        rsi_idx = s_rsi0[g];
        if (a.tpr) b = s_b0[g] + tid;
        else { rsi_idx += tid >> a.rp_shift; b = tid & (a.RP - 1u); }
    };
    for (uint32_t i = tid; i < 2u * SW; i += TB) staging_all[i] = 0;
    /* bit-spreading table of the k-window sums (aec_analyze_block): nibble j of entry v = bit j of v */
    __shared__ uint32_t s_lut[256];
    for (uint32_t v = tid; v < 256u; v += TB) {
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) w |= ((v >> j) & 1u) << (4 * j);
        s_lut[v] = w;
    }
    if (tid == 0) place(0, atomicAdd(a.ticket, 1u));
    __syncthreads();

    bool prev_have = false;                 /* the previous tile is packed and waits to be streamed out */
    /* my block of that tile and of the current one: kept in registers where there is room (short blocks),
     * worked out again from the tile's place otherwise */
    constexpr bool CARRY = (JT == 8 || JT == 16);
    uint64_t prev_rsi = 0, cur_rsi = 0;
    uint32_t prev_b = 0, cur_b = 0;
    uint32_t g = 0;                         /* ring entry of the current tile */

    for (uint32_t it = 0;; it++) {
        const uint32_t slot = it & 1u;
        const uint32_t gp = g == 0 ? 2u : g - 1u, gn = g == 2 ? 0u : g + 1u;
        const uint32_t tile = s_ticket[g];
        const bool have = tile < a.ntiles;
        uint32_t *staging = staging_all + slot * SW + 1u;
        uint64_t pv0 = 0;                   /* thread 0: the previous tile's prefix word, probed before packing */

        if (have) {
            PosFn ptile; ptile.has_end = 0; ptile.a = 0; ptile.rest = 0;      /* position map of the whole tile */
            /* ---- which block is mine ---- */
            uint64_t rsi_idx; uint32_t b;
            my_block(g, rsi_idx, b);
            if (CARRY) { cur_rsi = rsi_idx; cur_b = b; }
            uint32_t nblk = 0;
            if (rsi_idx + 1 < a.nrsi) nblk = c.rsi;
            else if (rsi_idx + 1 == a.nrsi) nblk = a.last_nblk;
            const bool valid = b < nblk;
            const uint32_t ref = (valid && c.pp && b == 0) ? 1u : 0u;

            /* ---- load, map, cost ---- */
            uint32_t d[JMAX];
            uint32_t refs = 0;
            bool small = false;                 /* mapped values known to be below 2^24 */
            BlockInfo bi; bi.opt = OPT_NONE; bi.klo = 0; bi.khi = c.kmax; bi.len = 0;
            {
                const uint64_t first = rsi_idx * (uint64_t)c.R + (uint64_t)b * J;
                const bool fast = valid && a.aligned && (first + J <= a.nsamples);
                if (valid) load_block<JT, B>(c, a.in, first, a.nsamples, fast, d);
                /* last raw sample of the previous block = last sample of the lane below */
                uint32_t prev = __shfl_up_sync(FULL, valid ? d[J - 1] : 0u, 1);
                if (valid && c.pp) {
                    if (b == 0) { refs = d[0]; prev = d[0]; }
                    else if (lane == 0) {
                        uint64_t pi = first - 1;
                        if (pi >= a.nsamples) pi = a.nsamples - 1;
                        prev = load_one<B>(a.in, pi, c.msb, a.aligned);
                    }
                    /* A block whose value range fits between the block and both ends of [0, M] cannot clip
                     * (every |delta| <= range <= min(u, M - u)): its mapped values are plain zig-zag codes
                     * of the signed differences.  Smooth data away from the limits always takes this path.
                     * For unsigned and for signed 32-bit samples the sign flip cancels in the differences
                     * and the range test can do without it. */
                    uint32_t umin, umax, sf = c.sflip;   /* sf: flip still to apply on the exact path */
                    if (c.sflip == 0u) {
                        umin = prev; umax = prev;
#pragma unroll
                        for (uint32_t i = 0; i < (JT ? (uint32_t)JT : J); i++) { umin = min(umin, d[i]); umax = max(umax, d[i]); }
                    } else if (c.sflip == 0x80000000u) {
                        int32_t smin = (int32_t)prev, smax = (int32_t)prev;
#pragma unroll
                        for (uint32_t i = 0; i < (JT ? (uint32_t)JT : J); i++) { smin = min(smin, (int32_t)d[i]); smax = max(smax, (int32_t)d[i]); }
                        umin = (uint32_t)smin ^ 0x80000000u; umax = (uint32_t)smax ^ 0x80000000u;
                    } else {
                        /* n < 32: the flip only cancels modulo 2^n, so flip in place */
                        prev ^= c.sflip; sf = 0u;
                        umin = prev; umax = prev;
#pragma unroll
                        for (uint32_t i = 0; i < (JT ? (uint32_t)JT : J); i++) {
                            d[i] ^= c.sflip;
                            umin = min(umin, d[i]); umax = max(umax, d[i]);
                        }
                    }
                    const uint32_t range = umax - umin;
                    if (JT != 0 && range <= umin && umax <= c.mask && range <= c.mask - umax) {
#pragma unroll
                        for (uint32_t i = 0; i < (JT ? (uint32_t)JT : J); i++) {
                            const uint32_t u = d[i];
                            const int32_t x = (int32_t)(u - prev);
                            d[i] = ((uint32_t)x << 1) ^ (uint32_t)(x >> 31);
                            prev = u;
                        }
                        small = range < (1u << 23);          /* zig-zag codes below 2^24 */
                    } else {
                        prev ^= sf;
#pragma unroll
                        for (uint32_t i = 0; i < (JT ? (uint32_t)JT : J); i++) {
                            const uint32_t u = d[i] ^ sf;
                            d[i] = aec_map_delta(prev, u, c.mask);
                            prev = u;
                        }
                    }
                    if (b == 0) d[0] = 0;
                }
                if (valid) bi = aec_analyze_block<JT>(c, d, ref, small, s_lut);
            }
            const bool is_zero = valid && bi.opt == OPT_ZERO;

            /* ---- zero-run structure of my 64-block segment ---- */
            uint32_t ball = __ballot_sync(FULL, is_zero);
            if (lane == 0) s_zb[warp] = ball;
            pair_barrier(warp);
            uint32_t len = bi.len, zcode = 0, zref = 0, zrun = 0;
            bool zinherit = false;              /* zero block that is not the first of its run */
            if (is_zero) {
                uint64_t m64 = (uint64_t)s_zb[warp & ~1u] | ((uint64_t)s_zb[warp | 1u] << 32);
                uint32_t q = tid & 63u;                       /* position in the aligned 64-slot group */
                uint32_t g0 = q - (b & 63u);                  /* where my segment starts in the group */
                uint32_t seg = b >> 6;
                uint32_t V = nblk - seg * 64u; if (V > 64u) V = 64u;
                uint64_t segmask = (m64 >> g0);
                if (V < 64u) segmask &= ((1ull << V) - 1ull);
                len = aec_zero_run(c, segmask, V, b, &zcode, &zref, &zrun);
                zinherit = (b & 63u) != 0 && ((segmask >> ((b & 63u) - 1u)) & 1ull);
            }
            if (zref)     /* run owner needs the reference sample of block 0 of the RSI */
                refs = load_one<B>(a.in, rsi_idx * (uint64_t)c.R, c.msb, a.aligned);
            const bool rsi_end = valid && (b + 1 == nblk);

            /* ---- intra-warp inclusive scans: CDS lengths and the k clamp chain ---- */
            const uint32_t kp = aec_kpair(bi.klo, bi.khi);    /* identity for zero/invalid/idl<=1 */
            uint32_t kinc = kp;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                /* lanes below `off` get their own pair back, and a clamp pair composed with itself is itself */
                kinc = aec_kcompose(__shfl_up_sync(FULL, kinc, off), kinc);
            }
            uint32_t kexc = __shfl_up_sync(FULL, kinc, 1);
            if (lane == 0) kexc = kident;

            PosFn pinc; pinc.has_end = 0; pinc.a = 0; pinc.rest = 0;   /* pad mode */
            uint32_t linc = len;                                      /* no-pad mode */
            if (c.pad) {
                pinc.has_end = rsi_end ? 1u : 0u; pinc.a = len; pinc.rest = 0;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    PosFn o = shfl_posfn(pinc, (int)lane - off);
                    if (lane >= (uint32_t)off) pinc = aec_pcompose(o, pinc);
                }
            } else {
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    uint32_t o = __shfl_up_sync(FULL, linc, off);
                    if (lane >= (uint32_t)off) linc += o;
                }
            }
            if (lane == 31) {
                s_wk[warp] = kinc;
                s_wlen[warp] = linc;
                if (c.pad) { s_wend[warp] = pinc.has_end; s_wa[warp] = (uint32_t)pinc.a; s_wrest[warp] = (uint32_t)pinc.rest; }
            }
            __syncthreads();                                           /* S2 */

            /* ---- cross-warp: every warp scans the NWARP warp totals with its first lanes ---- */
            PosFn pexc; pexc.has_end = 0; pexc.a = 0; pexc.rest = 0;  /* everything before me in the tile */
            uint32_t kbefore, ktile;
            {
                uint32_t wk = lane < (uint32_t)NWARP ? s_wk[lane] : kident;
#pragma unroll
                for (int off = 1; off < NWARP; off <<= 1) {
                    wk = aec_kcompose(__shfl_up_sync(FULL, wk, off), wk);
                }
                kbefore = __shfl_sync(FULL, wk, warp ? warp - 1 : 0);
                if (warp == 0) kbefore = kident;
                ktile = __shfl_sync(FULL, wk, NWARP - 1);
                kbefore = aec_kcompose(kbefore, kexc);
                if (c.pad) {
                    PosFn w; w.has_end = 0; w.a = 0; w.rest = 0;
                    if (lane < (uint32_t)NWARP) { w.has_end = s_wend[lane]; w.a = s_wa[lane]; w.rest = s_wrest[lane]; }
#pragma unroll
                    for (int off = 1; off < NWARP; off <<= 1) {
                        PosFn o = shfl_posfn(w, (int)lane - off);
                        if (lane >= (uint32_t)off) w = aec_pcompose(o, w);
                    }
                    PosFn bw = shfl_posfn(w, warp ? (int)warp - 1 : 0);
                    ptile = shfl_posfn(w, NWARP - 1);
                    if (warp) pexc = bw;
                    PosFn f = shfl_posfn(pinc, (int)lane - 1);          /* the lanes before me in my own warp */
                    if (lane > 0) pexc = aec_pcompose(pexc, f);
                } else {
                    uint32_t wl = lane < (uint32_t)NWARP ? s_wlen[lane] : 0u;
#pragma unroll
                    for (int off = 1; off < NWARP; off <<= 1) {
                        uint32_t o = __shfl_up_sync(FULL, wl, off);
                        if (lane >= (uint32_t)off) wl += o;
                    }
                    uint32_t before = __shfl_sync(FULL, wl, warp ? warp - 1 : 0);
                    if (warp == 0) before = 0;
                    ptile.a = __shfl_sync(FULL, wl, NWARP - 1);
                    pexc.a = before + (linc - len);
                }
            }

            /* publish this tile's aggregate for the scanner */
            if (tid == 0) {
                st_volatile_u64(&a.desc[tile], desc_pack_agg(ptile, ktile));
                a.tile_kagg[tile] = ktile;
                /* first look at the previous tile's prefix: the answer travels while this tile is packed */
                if (prev_have) pv0 = ld_volatile_u64(&a.pref[s_ticket[gp]]);
            }

            /* ---- pack my CDS into this tile's staging area, at local bit phase 0 ---- */
            uint32_t myoff = (uint32_t)pexc.a;
            if (late) {
                /* bit layout needs the absolute phase: wait for this tile's prefix */
                if (tid == 0) {
                    uint64_t pv = ld_volatile_u64(&a.pref[tile]);
                    uint32_t spins = 0;
                    while ((pv & 3) == 0) { pv = ld_volatile_u64(&a.pref[tile]); spin_guard(spins); }
                    s_base_cur = pv >> 12;
                }
                __syncthreads();
                const uint64_t bs = s_base_cur;
                myoff = (uint32_t)(aec_papply(pexc, bs) - ((bs >> 5) << 5));
            } else if (c.pad) {
                myoff = (uint32_t)aec_papply(pexc, 0);
            }
            if (valid && len) {
                const uint32_t side = (uint32_t)__cvta_generic_to_shared(&s_side[tid]);
                BitPack bp;
                bp.init(staging, myoff, side);
                if (is_zero) {
                    aec_pack_zero(c, bp, zcode, zref, refs);
                } else {
                    uint32_t k = bi.klo;
                    if (bi.opt == OPT_SPLIT && bi.klo != bi.khi) {
                        uint32_t kprev = aec_klo(kbefore);
                        if (kprev != aec_khi(kbefore)) {
                            /* a plateau block before the tile's first fixed k: its split position depends on
                             * the k carried into the tile, i.e. on this tile's prefix (rare) */
                            uint64_t pv = ld_volatile_u64(&a.pref[tile]);
                            uint32_t spins = 0;
                            while ((pv & 3) == 0) { pv = ld_volatile_u64(&a.pref[tile]); spin_guard(spins); }
                            kprev = aec_kapply((uint32_t)(pv >> 2) & 0x1Fu, kbefore);
                        }
                        k = aec_clampu(kprev, bi.klo, bi.khi);
                    }
                    aec_pack_block<JT>(c, bp, d, bi.opt, k, ref, refs);
                }
                bp.finish(side);
            }
            if (want_index)
                pend_all[slot * TB + tid] = make_uint2(myoff, len | (zrun << 12) | ((valid ? 1u : 0u) << 20) |
                                                              ((is_zero ? 1u : 0u) << 21) | ((zinherit ? 1u : 0u) << 22));
            if (tid == 0) { s_tend[slot] = ptile.has_end; s_ta[slot] = (uint32_t)ptile.a; s_trest[slot] = (uint32_t)ptile.rest; }
        }

        /* The previous tile's prefix has had a whole tile time to arrive.  A tile is claimed only when
         * the CTA is about to work on it, so a claimed tile never waits for anything but earlier tiles:
         * the scanner's fixed batches of 32 tiles always complete, and a tile always goes to the CTA
         * that is ready for it first (claiming ahead pins tiles to CTAs and lets one late CTA hold up
         * the whole chain: profiles/r1_g). */
        uint32_t nxt = 0xFFFFFFFFu;
        if (tid == 0) {
            if (prev_have) {
                uint32_t spins = 0;
                if (!have) pv0 = ld_volatile_u64(&a.pref[s_ticket[gp]]);
                while ((pv0 & 3) == 0) { pv0 = ld_volatile_u64(&a.pref[s_ticket[gp]]); spin_guard(spins); }
                const uint64_t base = pv0 >> 12;
                s_base = base;
                /* output word w0+i = staging bits [32 i - sh, 32 i - sh + 32); words [i_lo, i_hi) belong to this
                 * tile alone, the partial ones at either end go to the side arrays for the fix-up kernel */
                PosFn pt; pt.has_end = s_tend[slot ^ 1u]; pt.a = s_ta[slot ^ 1u]; pt.rest = s_trest[slot ^ 1u];
                const uint64_t end = aec_papply(pt, late ? base : 0ull) + (late ? 0ull : base);
                const uint64_t w0 = base >> 5, we = end >> 5;
                const bool head_partial = (base & 31u) != 0;
                const bool tail_partial = (end & 31u) != 0 && (we > w0 || !head_partial);
                const uint32_t nwhole = (uint32_t)(we - w0);
                uint32_t i_hi = nwhole;
                const uint64_t capw = a.out_cap_words > w0 ? a.out_cap_words - w0 : 0ull;
                if ((uint64_t)i_hi > capw) i_hi = (uint32_t)capw;
                /* 16-byte stores between the first and the last 16-byte boundary of the destination,
                 * single words (at most six) for the rest */
                const uint32_t i_lo = head_partial ? 1u : 0u;
                const uint32_t mis = ((uint32_t)(uintptr_t)(a.out_words + w0) >> 2) & 3u;
                uint32_t i_al = i_lo + ((0u - (mis + i_lo)) & 3u);
                if (i_al > i_hi) i_al = i_hi;
                if (i_hi < i_lo) i_al = i_lo;
                s_cp_w0 = w0;
                s_cp[0] = late ? 0u : (uint32_t)(base & 31u);             /* staging is at phase 0 unless late */
                s_cp[1] = ((late ? (uint32_t)(base & 31u) : 0u) + (uint32_t)(end - base) + 31u) >> 5;
                s_cp[2] = i_lo;
                s_cp[3] = i_al;
                s_cp[4] = i_hi > i_al ? (i_hi - i_al) >> 2 : 0u;
                s_cp[5] = i_hi;
                s_cp[6] = nwhole;
                s_cp[7] = (head_partial ? 1u : 0u) | (tail_partial ? 2u : 0u) | (end > base ? 4u : 0u);
            }
            /* claim the next tile now that the previous tile's prefix is here; the ticket's round trip
             * overlaps the copy-out.  (Claiming after the copy-out instead was tried: it needs one more
             * barrier per tile and the prefix is hardly ever late -- 766 polls for 16384 tiles,
             * profiles/r1_f_summary.md.) */
            if (have) nxt = atomicAdd(a.ticket, 1u);
            if (!prev_have || !have) { if (have) place(gn, nxt); else s_ticket[gn] = 0xFFFFFFFFu; }
        }
        __syncthreads();                                               /* S3 */

        /* ---- stream the previous tile's words out, shifted to the absolute bit phase ---- */
        if (prev_have) {
            uint32_t *pstage = staging_all + (slot ^ 1u) * SW + 1u;
            const uint64_t base = s_base;
            const uint32_t prev_tile = s_ticket[gp];
            const uint2 pe = want_index ? pend_all[(slot ^ 1u) * TB + tid] : make_uint2(0u, 0u);
            if (pe.y & (1u << 20)) {
                uint64_t p_rsi = prev_rsi; uint32_t p_b = prev_b;
                if (!CARRY) my_block(gp, p_rsi, p_b);
                const uint32_t p_len = pe.y & 0xFFFu, p_zrun = (pe.y >> 12) & 0xFFu;
                const uint64_t myabs = (late ? ((base >> 5) << 5) : base) + pe.x;
                if (p_b == 0 && a.rsi_offsets) a.rsi_offsets[p_rsi] = myabs;
                if (a.grp_index) {
                    /* group index for the warp-per-RSI decoder (aec_device.h) */
                    const uint32_t G = a.grp_G;
                    const uint32_t q = a.grp_magic ? __umulhi(p_b, a.grp_magic) : p_b;    /* b / G */
                    if (q * G == p_b && !(pe.y & (1u << 22))) a.grp_index[p_rsi * 32ull + q] = myabs;
                    if ((pe.y & (1u << 21)) && p_len && p_zrun > 1) {
                        /* I own a zero run: group starts inside it inherit their leading blocks from me */
                        uint32_t b0 = p_b + 1u - p_zrun;
                        for (uint32_t gs = (b0 / G + 1u) * G; gs <= p_b; gs += G)
                            a.grp_index[p_rsi * 32ull + gs / G] = ((uint64_t)(p_b - gs + 1u) << 56) | (myabs + p_len);
                    }
                }
            }
            const uint32_t sh = s_cp[0], nsl = s_cp[1], i_lo = s_cp[2], i_al = s_cp[3], nvec = s_cp[4], i_hi = s_cp[5];
            uint32_t *dst = a.out_words + s_cp_w0;
            {
                const uint32_t i_tl = i_al + 4u * nvec;
#pragma unroll 1
                for (uint32_t q = tid; q < nvec; q += TB) {
                    const uint32_t i = i_al + 4u * q;
                    const uint32_t p0 = pstage[(int)i - 1], p1 = pstage[i], p2 = pstage[i + 1], p3 = pstage[i + 2],
                                   p4 = pstage[i + 3];
                    uint4 o;
                    o.x = __byte_perm(__funnelshift_r(p1, p0, sh), 0, 0x0123);
                    o.y = __byte_perm(__funnelshift_r(p2, p1, sh), 0, 0x0123);
                    o.z = __byte_perm(__funnelshift_r(p3, p2, sh), 0, 0x0123);
                    o.w = __byte_perm(__funnelshift_r(p4, p3, sh), 0, 0x0123);
                    *reinterpret_cast<uint4 *>(dst + i) = o;
                }
                const uint32_t nh = i_al - i_lo;
                const uint32_t i = tid < nh ? i_lo + tid : i_tl + (tid - nh);
                if (i < i_hi) {
                    const uint32_t v = __funnelshift_r(pstage[i], pstage[(int)i - 1], sh);
                    dst[i] = __byte_perm(v, 0, 0x0123);
                }
            }
            if (tid == 0)
                a.head_c[prev_tile] = ((s_cp[7] & 5u) == 5u) ? (pstage[0] >> sh) : 0u;
            if (tid == 32) {
                const uint32_t nwhole = s_cp[6];
                a.tail_c[prev_tile] = (s_cp[7] & 2u) ? __funnelshift_r(pstage[nwhole], pstage[(int)nwhole - 1], sh) : 0u;
            }
            if (tid == 0 && have) place(gn, nxt);                      /* the ticket claimed before S3 has arrived by now */
            __syncthreads();                                           /* S4: everyone has read the words */
            {
                uint4 *z = reinterpret_cast<uint4 *>(staging_all + (slot ^ 1u) * SW);
                const uint32_t n4 = (nsl + 5u) >> 2;                   /* pad word + nsl words + one spare */
#pragma unroll 1
                for (uint32_t i = tid; i < n4; i += TB) z[i] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        if (!have) break;
        prev_have = true;
        if (CARRY) { prev_rsi = cur_rsi; prev_b = cur_b; }
        g = gn;
    }
}

/* Resolve the words shared between tiles (and with the bits a previous call
 * left in the stream's last partial word): one thread per tile boundary. */
__global__ void aec_encode_fixup_kernel(const AecEncArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - 1;   /* -1 = the seed */
    if (i >= (int64_t)a.ntiles_total) return;
    /* leave the control words as the next launch expects them (this kernel runs after the coder) */
    if (i < 0) *a.ticket = 0u;
    else { a.desc[i] = 0ull; a.pref[i] = 0ull; }
    const uint64_t total = a.tile_end[a.ntiles_total - 1];
    uint64_t bi = (i <= 0) ? a.seed_bits : a.tile_end[i - 1];
    uint64_t ei = (i < 0) ? a.seed_bits : a.tile_end[i];
    uint32_t v;
    if (i < 0) {
        if ((ei & 31u) == 0) return;
        v = a.seed_word;
    } else {
        if ((ei & 31u) == 0) return;
        bool head_partial = (bi & 31u) != 0;
        if (!((ei >> 5) > (bi >> 5) || !head_partial)) return;   /* no tail contribution of its own */
        v = a.tail_c[i];
    }
    const uint64_t word = ei >> 5;
    for (int64_t j = i + 1; j < (int64_t)a.ntiles_total; j++) {
        uint64_t ej = a.tile_end[j];
        v |= a.head_c[j];
        if ((ej >> 5) > word) break;
    }
    /* bytewise store: only bytes that belong to the stream and fit the buffer */
    uint64_t limit = (total + 7) >> 3;
    if (limit > a.out_cap_bytes) limit = a.out_cap_bytes;
    uint8_t *o = reinterpret_cast<uint8_t *>(a.out_words);
    for (int bq = 0; bq < 4; bq++) {
        uint64_t bidx = word * 4 + bq;
        if (bidx < limit) o[bidx] = (uint8_t)(v >> (24 - 8 * bq));
    }
}

/* Clamp pair of the whole launch (ordered composition of the per-tile pairs)
 * and a tile index after which k no longer depends on the incoming k: what a
 * multi-GPU shard publishes so its successors can chain k. */
__global__ void __launch_bounds__(1024) aec_encode_summary_kernel(const AecEncArgs a)
{
    /* 1024 threads, a contiguous run of tiles each; ordered composition inside the warps by shuffles,
     * across the warps by warp 0 */
    __shared__ uint32_t s_acc[32];
    __shared__ unsigned long long s_first[32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t n = a.ntiles_total;
    const uint64_t per = (n + 1023) / 1024;
    uint64_t b0 = threadIdx.x * per, b1 = b0 + per;
    if (b0 > n) b0 = n;
    if (b1 > n) b1 = n;
    const uint32_t ident = aec_kpair(0, a.cfg.kmax);
    uint32_t acc = ident;
    unsigned long long firstc = ~0ull;
    for (uint64_t t = b0; t < b1; t++) {
        acc = aec_kcompose(acc, a.tile_kagg[t]);
        /* once a run of tiles maps every k to one value, so does everything up to there */
        if (firstc == ~0ull && aec_klo(acc) == aec_khi(acc)) firstc = t;
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t o = __shfl_up_sync(FULL, acc, off);       /* lower lane = earlier tiles */
        if (lane >= (uint32_t)off) acc = aec_kcompose(o, acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        unsigned long long o = __shfl_xor_sync(FULL, firstc, off);
        if (o < firstc) firstc = o;
    }
    if (lane == 31) { s_acc[warp] = acc; s_first[warp] = firstc; }
    __syncthreads();
    if (warp != 0) return;
    acc = s_acc[lane];
    firstc = s_first[lane];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t o = __shfl_up_sync(FULL, acc, off);
        if (lane >= (uint32_t)off) acc = aec_kcompose(o, acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        unsigned long long o = __shfl_xor_sync(FULL, firstc, off);
        if (o < firstc) firstc = o;
    }
    if (lane == 31) {
        a.result[2] = aec_klo(acc);
        a.result[3] = aec_khi(acc);
        a.result[4] = (firstc == ~0ull) ? n : firstc;
        /* last 64 bits of the stream, right-aligned: what the next shard needs to complete the
         * word its own first bits share with this shard's last bits */
        const uint64_t total = a.tile_end[n - 1];
        const uint64_t endbyte = (total + 7) >> 3;
        const uint8_t *pb = reinterpret_cast<const uint8_t *>(a.out_words);
        uint64_t lo8 = 0, hi1 = 0;                      /* the last 9 bytes of the stream, big endian */
        for (int j = 0; j < 9; j++) {
            int64_t bi = (int64_t)endbyte - 9 + j;
            uint64_t byte = (bi >= 0 && (uint64_t)bi < a.out_cap_bytes) ? pb[bi] : 0u;
            if (j == 0) hi1 = byte; else lo8 = (lo8 << 8) | byte;
        }
        const uint32_t padb = (8u - (uint32_t)(total & 7u)) & 7u;   /* zero fill bits after the last stream bit */
        uint64_t t64 = padb ? ((lo8 >> padb) | (hi1 << (64u - padb))) : lo8;
        if (total < 64) t64 &= (total ? ((1ull << total) - 1ull) : 0ull);
        a.result[5] = t64;
        if (a.shard_out) {
            a.shard_out[0] = a.result[0];               /* bits of the shard (coded from bit 0) */
            a.shard_out[1] = aec_klo(acc); a.shard_out[2] = aec_khi(acc); a.shard_out[3] = t64;
        }
    }
}

/* One thread turns the gathered shard summaries into this rank's plan: exclusive scan of the bit lengths,
 * clamp chain of k, the predecessor's bits of the shared word (libaec_b200/parallel.py
 * plan_shards is the host model the gloo tests pin). */
__global__ void aec_shard_plan_kernel(const uint64_t *all, uint32_t world, uint32_t rank, const uint64_t *result, uint64_t *plan,
                                      uint64_t *plan_copy)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint64_t off = 0, total = 0, prev_tail = 0;
    uint32_t k = 0;
    uint64_t my_off = 0, my_bits = 0, my_head = 0;
    uint32_t my_k = 0;
    for (uint32_t r = 0; r < world; r++) {
        const uint64_t bits = all[4 * r];
        if (r == rank) {
            my_off = off; my_k = k; my_bits = bits;
            const uint32_t n = (uint32_t)(off & 31u);
            my_head = n ? ((prev_tail & ((1ull << n) - 1ull)) << (32u - n)) & 0xFFFFFFFFull : 0ull;
        }
        k = aec_clampu(k, (uint32_t)all[4 * r + 1], (uint32_t)all[4 * r + 2]);
        off += bits;
        if (bits) prev_tail = all[4 * r + 3];
        total += bits;
    }
    plan[PLAN_K_IN] = my_k;
    plan[PLAN_REPAIR_TILES] = (my_k != 0u && my_bits) ? result[4] + 1ull : 0ull;
    plan[PLAN_BIT_OFFSET] = my_off;
    plan[PLAN_HEAD_OR] = my_head;
    plan[PLAN_TOTAL_BITS] = total;
    plan[PLAN_MY_BITS] = my_bits;
    if (plan_copy)
        for (int i = 0; i < PLAN_WORDS; i++) plan_copy[i] = plan[i];
}

/* aec_place_bits_kernel with everything read from the plan */
__global__ void aec_place_bits_planned_kernel(const uint32_t *src, const uint64_t *plan, uint32_t *dst, uint64_t dst_cap_words,
                                              uint32_t global, uint32_t last_rank)
{
    const uint64_t nbits = plan[PLAN_MY_BITS];
    const uint64_t dst_bit = global ? plan[PLAN_BIT_OFFSET] : (plan[PLAN_BIT_OFFSET] & 31ull);
    const uint32_t head_or = (uint32_t)plan[PLAN_HEAD_OR];
    const uint64_t w0 = dst_bit >> 5;
    const uint32_t sh = (uint32_t)(dst_bit & 31u);
    const uint64_t endbit = dst_bit + nbits;
    /* words written: every word the shard touches, or -- when the destination is the whole stream -- the
     * words it owns: its partial last word belongs to the successor, who completes it */
    const uint64_t we = (global && !last_rank) ? (endbit >> 5) : ((endbit + 31) >> 5);
    const uint64_t nw = we > w0 ? we - w0 : 0;
    const uint64_t src_words = (nbits + 31) >> 5;
    /* four destination words per thread: five source words (one 16-byte load when aligned), one 16-byte
     * store when the destination allows it */
    const uint64_t nq = (nw + 3) >> 2;
    const bool dst16 = ((reinterpret_cast<uintptr_t>(dst + w0)) & 15u) == 0;
    const bool src16 = ((reinterpret_cast<uintptr_t>(src)) & 15u) == 0;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i0 = q << 2;
        uint32_t sw[5];
        sw[0] = (i0 >= 1 && i0 - 1 < src_words) ? __byte_perm(src[i0 - 1], 0, 0x0123) : 0u;
        if (src16 && i0 + 4 <= src_words) {
            const uint4 x = *reinterpret_cast<const uint4 *>(src + i0);
            sw[1] = __byte_perm(x.x, 0, 0x0123); sw[2] = __byte_perm(x.y, 0, 0x0123);
            sw[3] = __byte_perm(x.z, 0, 0x0123); sw[4] = __byte_perm(x.w, 0, 0x0123);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) sw[1 + j] = (i0 + j < src_words) ? __byte_perm(src[i0 + j], 0, 0x0123) : 0u;
        }
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint64_t i = i0 + j;
            v[j] = sh ? ((sw[j] << (32u - sh)) | (sw[j + 1] >> sh)) : sw[j + 1];
            if (w0 + i == (endbit >> 5) && (endbit & 31u)) v[j] &= ~(0xFFFFFFFFu >> (endbit & 31u));
            if (i == 0 && sh) v[j] = (v[j] & (0xFFFFFFFFu >> sh)) | (head_or & ~(0xFFFFFFFFu >> sh));
            v[j] = __byte_perm(v[j], 0, 0x0123);
        }
        if (dst16 && i0 + 4 <= nw && w0 + i0 + 4 <= dst_cap_words) {
            *reinterpret_cast<uint4 *>(dst + w0 + i0) = make_uint4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (i0 + j < nw && w0 + i0 + j < dst_cap_words) dst[w0 + i0 + j] = v[j];
        }
    }
}


/* dst[dst_bit ..) = src[0 .. nbits): one thread per destination word. */
__global__ void aec_place_bits_kernel(const uint32_t *src, uint64_t nbits, uint32_t *dst, uint64_t dst_bit,
                                      uint64_t dst_cap_words, uint32_t head_or)
{
    const uint64_t w0 = dst_bit >> 5;
    const uint32_t sh = (uint32_t)(dst_bit & 31u);
    const uint64_t nw = ((dst_bit + nbits + 31) >> 5) - w0;
    const uint64_t src_words = (nbits + 31) >> 5;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (uint64_t)gridDim.x * blockDim.x) {
        /* destination word i holds source bits [32*i - sh, 32*i - sh + 32) */
        uint32_t hi = (i >= 1 && i - 1 < src_words) ? __byte_perm(src[i - 1], 0, 0x0123) : 0u;
        uint32_t lo = (i < src_words) ? __byte_perm(src[i], 0, 0x0123) : 0u;
        uint32_t v = sh ? ((hi << (32u - sh)) | (lo >> sh)) : lo;
        /* clear bits past the end of the stream in the last word */
        uint64_t endbit = dst_bit + nbits;
        if (w0 + i == (endbit >> 5) && (endbit & 31u)) v &= ~(0xFFFFFFFFu >> (endbit & 31u));
        if (i == 0 && sh) v = (v & (0xFFFFFFFFu >> sh)) | (head_or & ~(0xFFFFFFFFu >> sh));
        if (w0 + i < dst_cap_words) dst[w0 + i] = __byte_perm(v, 0, 0x0123);
    }
}

template <int JT, int B>
cudaError_t launch_variant(const AecEncArgs &a, uint32_t smem_bytes, int num_sms, cudaStream_t st)
{
    auto kern = aec_encode_kernel<JT, B>;
    /* opt in to large dynamic shared memory (per device); the 48 KiB default limit counts the
     * kernel's static shared variables as well, so leave room for them */
    static AecSmemOptIn optin;
    cudaError_t e = optin.ensure(kern, smem_bytes, 40 * 1024);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TileCfg<JT>::TB, smem_bytes);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    /* CTA 0 is the scanner, the others are workers; all of them must be resident */
    uint64_t grid = (uint64_t)occ * (uint64_t)num_sms;
    if (grid > a.ntiles + 1) grid = a.ntiles + 1;
    if (grid < 2) grid = 2;
    if (a.ntiles == 0) return cudaSuccess;
    /* cooperative launch: the runtime guarantees that the scanner and every worker are resident at
     * the same time (or refuses the launch) whatever else shares the device */
    static const bool coop = getenv("AECB200_NO_COOP") == nullptr;
    if (!coop) {
        kern<<<(unsigned)grid, TileCfg<JT>::TB, smem_bytes, st>>>(a);
        return cudaGetLastError();
    }
    void *kargs[] = {const_cast<AecEncArgs *>(&a)};
    return cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)grid), dim3(TileCfg<JT>::TB), kargs,
                                       smem_bytes, st);
}

template <int JT>
cudaError_t launch_j(const AecEncArgs &a, uint32_t smem, int sms, cudaStream_t st)
{
    switch (a.cfg.B) {
    case 1: return launch_variant<JT, 1>(a, smem, sms, st);
    case 2: return launch_variant<JT, 2>(a, smem, sms, st);
    case 3: return launch_variant<JT, 3>(a, smem, sms, st);
    default: return launch_variant<JT, 4>(a, smem, sms, st);
    }
}

} // namespace

uint32_t aec_encode_tile_blocks(uint32_t J)
{
    if (J == 16) return AEC_TB16;
    return (J == 8 || J == 32) ? 256u : 128u;
}

uint32_t aec_encode_staging_words(const AecCfg &c)
{
    uint32_t TB = aec_encode_tile_blocks(c.J);
    /* every CDS is at most idl + 1 + n + J*n bits, plus RSI padding; one zero pad word in
     * front, spare words behind, a multiple of four words so that both areas stay 16-byte aligned */
    uint64_t bits = 31ull + (uint64_t)TB * (c.idl + 1ull + (uint64_t)c.J * c.n + c.n + 8ull) + 64ull;
    return (uint32_t)((bits / 32ull + 8ull + 3ull) & ~3ull);
}

cudaError_t aec_encode_summary_launch(const AecEncArgs &a, cudaStream_t st)
{
    aec_encode_summary_kernel<<<1, 1024, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t aec_shard_plan_launch(const uint64_t *all, uint32_t world, uint32_t rank, const uint64_t *result, uint64_t *plan,
                                  uint64_t *plan_copy, cudaStream_t st)
{
    aec_shard_plan_kernel<<<1, 32, 0, st>>>(all, world, rank, result, plan, plan_copy);
    return cudaGetLastError();
}

cudaError_t aec_place_bits_planned_launch(const uint32_t *src, const uint64_t *plan, uint32_t *dst, uint64_t dst_cap_words,
                                          uint32_t global, uint32_t last_rank, int num_sms, cudaStream_t st)
{
    aec_place_bits_planned_kernel<<<(unsigned)(num_sms * 8), 256, 0, st>>>(src, plan, dst, dst_cap_words, global, last_rank);
    return cudaGetLastError();
}

cudaError_t aec_place_bits_launch(const uint32_t *src, uint64_t nbits, uint32_t *dst, uint64_t dst_bit,
                                  uint64_t dst_cap_words, uint32_t head_or, cudaStream_t st)
{
    uint64_t nw = ((dst_bit + nbits + 31) >> 5) - (dst_bit >> 5);
    if (nw == 0) return cudaSuccess;
    unsigned grid = (unsigned)((nw + 255) / 256 > 148 * 16 ? 148 * 16 : (nw + 255) / 256);
    aec_place_bits_kernel<<<grid, 256, 0, st>>>(src, nbits, dst, dst_bit, dst_cap_words, head_or);
    return cudaGetLastError();
}

cudaError_t aec_encode_launch(const AecEncArgs &args, int num_sms, cudaStream_t st)
{
    AecEncArgs a = args;
    {   /* derived geometry the kernel would otherwise divide for */
        const uint32_t TB = aec_encode_tile_blocks(a.cfg.J);
        a.tpr = a.RP >= TB ? a.RP / TB : 0u;
        a.rp_shift = 0; while ((1u << a.rp_shift) < a.RP) a.rp_shift++;
        a.tile_rsi_shift = 0; while (a.RP < TB && ((a.RP << a.tile_rsi_shift) < TB)) a.tile_rsi_shift++;
        a.grp_magic = a.grp_G > 1u ? (uint32_t)((1ull << 32) / a.grp_G + 1ull) : 0u;
    }
    /* two staging areas and two areas of per-thread notes (aec_encode_kernel) */
    uint32_t smem = 2u * a.staging_words * 4u + 2u * aec_encode_tile_blocks(a.cfg.J) * 8u;
    cudaError_t e;
    switch (a.cfg.J) {
    case 8:  e = launch_j<8>(a, smem, num_sms, st); break;
    case 16: e = launch_j<16>(a, smem, num_sms, st); break;
    case 32: e = launch_j<32>(a, smem, num_sms, st); break;
    case 64: e = launch_j<64>(a, smem, num_sms, st); break;
    default: e = launch_j<0>(a, smem, num_sms, st); break;
    }
    if (e != cudaSuccess) return e;
    unsigned nthreads = (unsigned)(a.ntiles_total + 1);
    aec_encode_fixup_kernel<<<(nthreads + 255) / 256, 256, 0, st>>>(a);
    return cudaGetLastError();
}
