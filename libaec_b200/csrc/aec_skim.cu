/*
 * aec_skim.cu -- parallel discovery of RSI boundaries in a stream that comes without an index.
 *
 * The reference finds the start of RSI r+1 only by decoding RSI r (/root/reference/src/decode.c:402-421
 * m_id, :288-340 direct_get_fs): every libaec 0.3.4 stream a caller can hand to aec_decode /
 * aec_buffer_decode / SZ_BufftoBuffDecompress is of that kind.  A GPU cannot afford that chain
 * (one thread: ~80 us per RSI), so the chain is cut into pieces that do not depend on where the
 * stream is entered:
 *
 *   1. aec_skim_level0_kernel   for EVERY bit position p of a window: if a coded data set (CDS) started
 *                               at p, how long would it be and how many blocks would it stand for
 *                               (T0[p]; with dense tables also R[p] = the same for the first CDS of an RSI,
 *                               which carries the reference sample).  One thread per position, the window's words and
 *                               their running popcount staged in shared memory; a unary section is
 *                               skipped by rank/select on that popcount instead of bit by bit.
 *   2. aec_skim_double_kernel   T(j+1)[p] = T(j)[p] then T(j)[p + length]: the CDS chain from p after
 *                               2^(j+1) steps (pointer doubling; bits and blocks add up).
 *   3. aec_skim_rsi_kernel      H[p] = length of a whole RSI that starts at p: the first CDS from R[p],
 *                               then a greedy descent through the levels until exactly `rsi` blocks
 *                               are accounted for (run-of-zero-segment codes stand for "up to the end of
 *                               the 64-block segment" and are resolved here, where the block number is known).
 *      aec_skim_rsi_sparse_kernel  R and H for the candidates only: an RSI starts where a chain of CDSs ends, and
 *                               the top-level chains of ALL positions end on a few per cent of them (the
 *                               doubling passes mark those in H, aec_skim_core.cuh: SK_CAND); the walk works
 *                               out the rare start that was not marked itself and turns the stream to dense
 *                               tables (aec_skim_rsi_kernel: every position) when that happens often.
 *      aec_skim_hchase_list_kernel / aec_skim_hdouble_kernel  streams of many RSIs per window: the length of
 *                               eight RSIs in a row, so that the walk takes an eighth of its steps.
 *   4. aec_skim_walk_kernel     the only serial part: one load of H per RSI from the stream's known
 *                               first bit; RSIs whose H is not available (truncated or corrupt stream,
 *                               a chain leaving the window) are skimmed CDS by CDS like the reference does.
 *
 * Everything up to the walk is independent of the entry point, so it runs on all SMs; windows bound
 * the table memory (4 bytes x (levels + 2 .. 4.25) per stream bit).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "aec_skim_core.cuh"
#include "aec_device.h"

namespace {

constexpr int SK_THREADS = 256;
#ifndef AEC_SK_TILE
#define AEC_SK_TILE 8192
#endif
constexpr uint32_t SK_TILE = AEC_SK_TILE;   /* bit positions per CTA of the level-0 kernel */

/* state[5] bit 0: the walk asked for dense tables (it runs on its own stream, next to the table kernels of the
 * following window).  A window latches the request once, before its first kernel, into bit 63 of its list
 * counter, so that all its kernels and its walk agree on what the tables hold. */
constexpr unsigned long long SK_DENSE_BIT = 1ull << 63;
__global__ void aec_skim_begin_kernel(const AecSkimArgs a)
{
    a.state[6 + a.set] = (*reinterpret_cast<volatile const uint64_t *>(a.state + 5) & 1ull) ? SK_DENSE_BIT : 0ull;
}
__device__ __forceinline__ bool sk_sparse_now(const AecSkimArgs &a)
{
    return a.sparse && !(a.state[6 + a.set] & SK_DENSE_BIT);
}

/* Stage the stream words of `span` bit positions from window-relative bit tile0 (plus the look-ahead) as
 * big-endian words w[0..nwords] with their running popcount pre[0..nwords]; all threads of the CTA call it.
 * One bulk asynchronous copy (cp.async.bulk, the TMA engine's 1-D form) brings the 16-byte aligned part
 * straight into shared memory and signals an mbarrier; what lies behind the last whole 16 bytes of the
 * stream comes by ordinary loads (zeros past the end). */
__device__ __forceinline__ void sk_stage_words(const AecSkimArgs &a, uint32_t tile0, uint32_t nwords, uint32_t *w, uint32_t *pre,
                                               uint32_t *s_part, unsigned long long *s_mbar)
{
    const uint32_t tid = threadIdx.x;
    const uint64_t word0 = (a.wb + tile0) >> 5;
    const uint64_t total_words = (a.nbits + 31ull) >> 5;
    uint32_t nbulk = 0;                                 /* words the bulk copy delivers */
    if (a.bulk && word0 < total_words) {
        const uint64_t avail = (total_words - word0) & ~3ull;
        const uint32_t want = (nwords + 1u) & ~3u;
        nbulk = avail < want ? (uint32_t)avail : want;
    }
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(s_mbar);
    if (nbulk) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = nbulk * 4u;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(w);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst), "l"(a.in_words + word0), "r"(bytes), "r"(mbar) : "memory");
        }
    }
    for (uint32_t i = nbulk + tid; i <= nwords; i += SK_THREADS) {
        const uint64_t wi = word0 + i;
        w[i] = wi < total_words ? __byte_perm(__ldg(a.in_words + wi), 0, 0x0123) : 0u;
    }
    if (nbulk) {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar) : "memory");
        for (uint32_t i = tid; i < nbulk; i += SK_THREADS) w[i] = __byte_perm(w[i], 0, 0x0123);   /* to big endian */
    }
    __syncthreads();
    /* rounds of a block-wide exclusive scan */
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nwords + 1u; base += SK_THREADS) {
        const uint32_t i = base + tid;
        const uint32_t v = i < nwords ? (uint32_t)__popc(w[i]) : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, off);
            if ((tid & 31u) >= (uint32_t)off) inc += o;
        }
        if ((tid & 31u) == 31u) s_part[tid >> 5] = inc;
        __syncthreads();
        uint32_t before = carry;
        for (uint32_t ww = 0; ww < (tid >> 5); ww++) before += s_part[ww];
        if (i <= nwords) pre[i] = before + inc - v;
        uint32_t tot = 0;
        for (uint32_t ww = 0; ww < SK_THREADS / 32; ww++) tot += s_part[ww];
        carry += tot;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SK_THREADS)
aec_skim_level0_kernel(const AecSkimArgs a)
{
    if (a.state[2] & 1ull) return;                      /* the walk has already ended */
    const AecCfg &c = a.cfg;
    extern __shared__ __align__(16) uint32_t sk_smem[];
    const uint32_t nwords = SK_TILE / 32u + a.la_words;
    uint32_t *w = sk_smem;                              /* [nwords + 1], padded to a multiple of 4 words */
    uint32_t *pre = sk_smem + ((nwords + 1u + 3u) & ~3u);   /* [nwords + 1] */
    __shared__ uint32_t s_part[SK_THREADS / 32];
    __shared__ __align__(8) unsigned long long s_mbar;
    const uint32_t tid = threadIdx.x;
    const uint32_t tile0 = blockIdx.x * SK_TILE;        /* window-relative */
    sk_stage_words(a, tile0, nwords, w, pre, s_part, &s_mbar);
    const bool sparse = sk_sparse_now(a);
    /* bits of the stream that exist, relative to the tile */
    const uint64_t tile_abs = a.wb + tile0;
    const uint32_t limit = a.nbits > tile_abs ? (uint32_t)min((unsigned long long)(a.nbits - tile_abs), 0x7FFFFFFFull) : 0u;
    for (uint32_t q = tid; q < SK_TILE; q += SK_THREADS) {
        const uint32_t p = tile0 + q;
        if (p >= a.np) break;
        uint32_t t0 = 0u, r0 = 0u;
        if (q < limit) {
            t0 = sk_entry(c, w, pre, nwords, q, limit, 0u);
            r0 = (c.pp && !sparse) ? sk_entry(c, w, pre, nwords, q, limit, 1u) : t0;
        }
        a.T[p] = t0;
        if (sparse) a.H[p] = 0u;                        /* the doubling passes mark the candidates in here; R: candidates only, later */
        else a.R[p] = r0;
    }
}

/* SK_DP consecutive positions per thread: 16-byte loads, SK_DP gathers in flight, 16-byte stores */
#ifndef AEC_SK_DP
#define AEC_SK_DP 4
#endif
constexpr int SK_DP = AEC_SK_DP;
__global__ void __launch_bounds__(SK_THREADS)
aec_skim_double_kernel(const AecSkimArgs a, uint32_t level)
{
    if (a.state[2] & 1ull) return;
    const uint32_t p = (blockIdx.x * SK_THREADS + threadIdx.x) * (uint32_t)SK_DP;      /* np is a multiple of 32 */
    if (p >= a.np) return;
    const uint32_t *src = a.T + (size_t)level * a.np;
    uint32_t *dst = a.T + (size_t)(level + 1u) * a.np;
    uint32_t xs[SK_DP], y[SK_DP], r[SK_DP];
#pragma unroll
    for (int v = 0; v < SK_DP / 4; v++) {
        const uint4 x = *reinterpret_cast<const uint4 *>(src + p + 4 * v);
        xs[4 * v] = x.x; xs[4 * v + 1] = x.y; xs[4 * v + 2] = x.z; xs[4 * v + 3] = x.w;
    }
#pragma unroll
    for (int i = 0; i < SK_DP; i++) {
        const uint32_t q = p + i + sk_len(xs[i]);
        y[i] = (sk_jump(xs[i]) && q < a.np) ? __ldg(src + q) : 0u;
    }
#pragma unroll
    for (int i = 0; i < SK_DP; i++) {
        const uint32_t blk = sk_blk(xs[i]) + sk_blk(y[i]), len = sk_len(xs[i]) + sk_len(y[i]);
        r[i] = (sk_jump(y[i]) && blk <= 0xFFFu && len <= 0xFFFFFu) ? ((len << 12) | blk) : 0u;
    }
#pragma unroll
    for (int v = 0; v < SK_DP / 4; v++)
        *reinterpret_cast<uint4 *>(dst + p + 4 * v) = make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
    if (!sk_sparse_now(a)) return;
    /* candidates for RSI starts: behind a run-of-zero-segment code (level 0 is at hand in the first pass) and
     * where the chains of the top level end (the last pass has just worked them out) */
    if (level == 0u) {
#pragma unroll
        for (int i = 0; i < SK_DP; i++)
            if (sk_ros(xs[i])) { const uint32_t t = sk_mark_pos(a.cfg, p + i + sk_len(xs[i])); if (t < a.nh_eff) a.H[t] = SK_CAND; }
    }
    if (level + 2u == a.LV) {
#pragma unroll
        for (int i = 0; i < SK_DP; i++)
            if (sk_jump(r[i])) { const uint32_t t = sk_mark_pos(a.cfg, p + i + sk_len(r[i])); if (t < a.nh_eff) a.H[t] = SK_CAND; }
    }
}

/* H[p] <- bits from p to the start of the next RSI when an RSI starts at p (0: not available).
 * Same walk as sk_rsi_len (aec_skim_core.cuh), NC candidates per thread in lock step so that their
 * table look-ups -- each a dependent, mostly uncached load -- are in flight together. */
#ifndef AEC_SK_NC
#define AEC_SK_NC 4
#endif
constexpr int SK_NC = AEC_SK_NC;
#ifndef AEC_SK_ROUNDS
#define AEC_SK_ROUNDS 2
#endif
constexpr int SK_ROUNDS = AEC_SK_ROUNDS;   /* a CTA takes SK_ROUNDS x SK_NC x 256 CONSECUTIVE candidates: the chains of neighbouring
                                     * candidates stay within a few thousand positions of each other at every level, so the
                                     * look-ups of one CTA fall into a handful of compact table regions that its L1 keeps */
/* The descent of sk_rsi_len for SK_NC starts p[i] at once (live[i]: there is a candidate, first_e[i] the entry of
 * its first CDS); out[i] = RSI length or 0. */
__device__ __forceinline__ void sk_descend_nc(const AecSkimArgs &a, const uint32_t (&p)[SK_NC], const uint32_t (&first_e)[SK_NC],
                                              bool (&live)[SK_NC], uint32_t (&out)[SK_NC])
{
    const AecCfg &c = a.cfg;
    const uint32_t np = a.np, rsi = c.rsi;
    const int top = (int)a.LV - 1;
    uint32_t q[SK_NC], rem[SK_NC];
#pragma unroll
    for (int i = 0; i < SK_NC; i++) {
        const uint32_t first = live[i] ? first_e[i] : 0u;
        uint32_t b = sk_blk(first);
        if (b == 0u) b = rsi < 64u ? rsi : 64u;                 /* run-of-zero-segment at block 0 */
        live[i] = live[i] && first >= 0x1000u && b <= rsi;
        q[i] = p[i] + sk_len(first);
        rem[i] = live[i] ? rsi - b : 0u;
    }
    for (int guard = 0; guard < 4096; guard++) {
        for (int j = top; j >= 0; j--) {
            const uint32_t *Tj = a.T + (size_t)j * np;
            bool again;
            do {
                uint32_t e[SK_NC];
#pragma unroll
                for (int i = 0; i < SK_NC; i++)          /* 2^j CDSs stand for 2^j blocks at least: no look-up that cannot fit */
                    e[i] = (rem[i] >= (1u << j) && q[i] < np) ? __ldg(Tj + q[i]) : 0u;
                again = false;
#pragma unroll
                for (int i = 0; i < SK_NC; i++) {
                    if (sk_jump(e[i]) && sk_blk(e[i]) <= rem[i]) {
                        q[i] += sk_len(e[i]); rem[i] -= sk_blk(e[i]);
                        again = again || rem[i] != 0u;
                    }
                }
            } while (again && j == top);                        /* below the top level a step fits at most once */
        }
        /* whoever still has blocks left stands at a CDS that is not a plain step */
        bool progress = false;
#pragma unroll
        for (int i = 0; i < SK_NC; i++) {
            if (rem[i] == 0u) continue;
            const uint32_t e = q[i] < np ? __ldg(a.T + q[i]) : 0u;
            const uint32_t b = rsi - rem[i];
            if (sk_ros(e)) {
                const uint32_t seg = 64u - (b & 63u);
                rem[i] -= rem[i] < seg ? rem[i] : seg;          /* decode.c:528-530 */
                q[i] += sk_len(e);
                progress = true;
            } else if (sk_jump(e) && sk_blk(e) <= rem[i]) {
                q[i] += sk_len(e); rem[i] -= sk_blk(e);
                progress = true;
            } else { live[i] = false; rem[i] = 0u; }
        }
        if (!progress) break;
    }
#pragma unroll
    for (int i = 0; i < SK_NC; i++) {
        uint32_t end = q[i];
        if (c.pad) end = (end + 7u) & ~7u;                      /* windows start on byte boundaries */
        out[i] = (live[i] && rem[i] == 0u) ? end - p[i] : 0u;
    }
}

__global__ void __launch_bounds__(SK_THREADS)
aec_skim_rsi_kernel(const AecSkimArgs a)
{
    if (a.state[2] & 1ull) return;
    if (sk_sparse_now(a)) return;                       /* aec_skim_rsi_sparse_kernel has done the candidates */
    const uint32_t step = a.cfg.pad ? 8u : 1u;          /* padded RSIs start on byte boundaries */
    for (int round = 0; round < SK_ROUNDS; round++) {
        /* candidate i of this thread: consecutive threads take consecutive positions (coalesced R / H accesses) */
        const uint32_t base = (blockIdx.x * (uint32_t)SK_ROUNDS + (uint32_t)round) * (uint32_t)(SK_NC * SK_THREADS) + threadIdx.x;
        uint32_t p[SK_NC], first_e[SK_NC], out[SK_NC];
        bool live[SK_NC];
#pragma unroll
        for (int i = 0; i < SK_NC; i++) {
            p[i] = (base + (uint32_t)i * SK_THREADS) * step;
            live[i] = p[i] < a.nh_eff;
            first_e[i] = live[i] ? a.R[p[i]] : 0u;
        }
        sk_descend_nc(a, p, first_e, live, out);
#pragma unroll
        for (int i = 0; i < SK_NC; i++)
            if (p[i] < a.nh_eff) a.H[p[i]] = out[i];
    }
}

/* The candidates only.  A CTA takes SK_CHUNK consecutive bit positions: it stages their stream words like the
 * level-0 kernel does (the first CDS of an RSI carries the reference sample: its entry R is worked out here,
 * for the candidates, instead of for every position), collects the marked positions in shared memory -- in
 * batches, so that a window full of marks (fixed-length CDSs: chains that never merge) still fits -- works
 * out their RSI lengths SK_NC per thread like the dense pass, and appends the positions that have one to the
 * window's list (the long-jump passes run over that list). */
#ifndef AEC_SK_CHUNK
#define AEC_SK_CHUNK 16384
#endif
constexpr uint32_t SK_CHUNK = AEC_SK_CHUNK;
constexpr uint32_t SK_LIST = 6144;
constexpr uint32_t SK_CHUNK_WORDS = SK_CHUNK / 32u + 72u;   /* + the longest CDS (sk_lookahead_words <= 68) */
__global__ void __launch_bounds__(SK_THREADS)
aec_skim_rsi_sparse_kernel(const AecSkimArgs a)
{
    if (a.state[2] & 1ull) return;
    if (!sk_sparse_now(a)) return;
    const AecCfg &c = a.cfg;
    __shared__ __align__(16) uint32_t s_w[SK_CHUNK_WORDS + 4];
    __shared__ uint32_t s_pre[SK_CHUNK_WORDS + 4];
    __shared__ uint32_t s_list[SK_LIST];                /* the batch; its front is reused for the positions to be listed */
    __shared__ uint32_t s_part[SK_THREADS / 32];
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ uint32_t s_n, s_m;
    __shared__ unsigned long long s_at;
    const uint32_t tid = threadIdx.x;
    const uint32_t sh = c.pad ? 3u : 0u;
    const uint32_t pos0 = blockIdx.x * SK_CHUNK;        /* window-relative */
    if (pos0 >= a.nh_eff) return;
    const uint32_t pos1 = pos0 + SK_CHUNK < a.nh_eff ? pos0 + SK_CHUNK : a.nh_eff;
    const uint32_t nwords = SK_CHUNK / 32u + a.la_words;
    sk_stage_words(a, pos0, nwords, s_w, s_pre, s_part, &s_mbar);
    const uint64_t chunk_abs = a.wb + pos0;
    const uint32_t limit = a.nbits > chunk_abs ? (uint32_t)min((unsigned long long)(a.nbits - chunk_abs), 0x7FFFFFFFull) : 0u;
    constexpr uint32_t PER_PASS = SK_THREADS * 16u;     /* slots looked at between two checks of the batch's fill */
    const uint32_t slot0 = pos0 >> sh, slot1 = (pos1 + (1u << sh) - 1u) >> sh;
    uint32_t next = slot0;
    while (next < slot1) {
        if (tid == 0) { s_n = 0u; s_m = 0u; }
        __syncthreads();
        /* collect: whole passes while the batch surely has room for another one */
        while (next < slot1) {
            if (sh == 0u) {
                uint4 h[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {           /* four 16-byte loads in flight per thread */
                    const uint32_t sl = next + ((uint32_t)k * SK_THREADS + tid) * 4u;
                    h[k] = sl + 3u < slot1 ? *reinterpret_cast<const uint4 *>(a.H + sl) : make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t sl = next + ((uint32_t)k * SK_THREADS + tid) * 4u;
                    if (sl + 3u < slot1) {
                        if (h[k].x == SK_CAND) s_list[atomicAdd(&s_n, 1u)] = sl;
                        if (h[k].y == SK_CAND) s_list[atomicAdd(&s_n, 1u)] = sl + 1u;
                        if (h[k].z == SK_CAND) s_list[atomicAdd(&s_n, 1u)] = sl + 2u;
                        if (h[k].w == SK_CAND) s_list[atomicAdd(&s_n, 1u)] = sl + 3u;
                    } else {
                        for (uint32_t t = 0; t < 4u; t++)
                            if (sl + t < slot1 && a.H[sl + t] == SK_CAND) s_list[atomicAdd(&s_n, 1u)] = sl + t;
                    }
                }
            } else {
                for (uint32_t k = 0; k < 16u; k++) {
                    const uint32_t sl = next + k * SK_THREADS + tid;
                    if (sl < slot1 && a.H[sl << sh] == SK_CAND) s_list[atomicAdd(&s_n, 1u)] = sl << sh;
                }
            }
            next += PER_PASS;
            __syncthreads();
            if (s_n + PER_PASS > SK_LIST) break;
            __syncthreads();                            /* nobody adds to s_n before everybody has read it */
        }
        __syncthreads();
        const uint32_t n = s_n;
        for (uint32_t base = 0; base < n; base += SK_NC * SK_THREADS) {
            uint32_t p[SK_NC], first_e[SK_NC], out[SK_NC];
            bool live[SK_NC], have[SK_NC];
#pragma unroll
            for (int i = 0; i < SK_NC; i++) {
                const uint32_t k = base + (uint32_t)i * SK_THREADS + tid;
                live[i] = have[i] = k < n;
                p[i] = live[i] ? s_list[k] : 0u;
                const uint32_t q = p[i] - pos0;
                first_e[i] = (live[i] && q < limit) ? sk_entry(c, s_w, s_pre, nwords, q, limit, c.pp ? 1u : 0u) : 0u;
            }
            __syncthreads();                            /* this round's entries are read: the front of s_list takes the results */
            sk_descend_nc(a, p, first_e, live, out);
#pragma unroll
            for (int i = 0; i < SK_NC; i++) {
                if (!have[i]) continue;
                a.H[p[i]] = out[i];                     /* replaces the mark */
                a.R[p[i]] = first_e[i];                 /* the group index wants it again */
                if (out[i] && a.cand_list) s_list[atomicAdd(&s_m, 1u)] = p[i];
            }
        }
        __syncthreads();
        if (a.cand_list) {
            const uint32_t m = s_m;
            if (tid == 0) s_at = m ? atomicAdd(reinterpret_cast<unsigned long long *>(a.state + 6 + a.set), (unsigned long long)m) : 0ull;
            __syncthreads();
            const unsigned long long at = s_at;
            for (uint32_t j = tid; j < m; j += SK_THREADS) {
                if (at + j < a.cand_cap) a.cand_list[at + j] = s_list[j];
                else a.H[s_list[j]] = 0u;               /* no room in the list: not a candidate (the walk works it out itself) */
            }
        }
        __syncthreads();
    }
}

/* state: [0] bit position of the next RSI, [1] RSIs found, [2] flags (1 ended, 2 data error),
 * [3] RSIs taken from the tables (diagnostics) */
__global__ void aec_skim_walk_kernel(const AecSkimArgs a)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (a.state[2] & 1ull) return;
    const AecCfg &c = a.cfg;
    SkWalk s; s.pos = a.state[0]; s.found = a.state[1]; s.flags = 0; s.fast = a.state[3]; s.slow = 0;
    a.state[4] = s.found;
    const uint64_t found0 = s.found;
    BitRd br;
    br.init(a.in_words, (a.nbits + 31ull) >> 5, a.nbits);
    const uint32_t *H = a.H;
    const uint32_t *H8 = a.H8 ? a.H8 + a.np : nullptr;
    while (sk_walk_step(c, br, a.nbits, a.wb, a.nh_eff, a.last, a.offsets, a.max_rsi, s,
                        [H](uint64_t rel) { return __ldcg(H + rel); }, a.grp_index, H8 != nullptr,
                        [H8](uint64_t rel) { return __ldcg(H8 + rel); }, sk_sparse_now(a),
                        [&a, &c, &br](uint64_t rel) {
                            /* a start nobody marked has no R entry either: parse its first CDS, keep it for the group index */
                            const uint32_t first = sk_first_entry_serial(c, br, a.wb + rel);
                            a.R[rel] = first;
                            return sk_rsi_len(c, a.T, a.LV, a.np, (uint32_t)rel, first); })) { }
    a.state[0] = s.pos; a.state[1] = s.found; a.state[2] = s.flags; a.state[3] = s.fast;
    if (sk_sparse_now(a) && sk_walk_wants_dense(s.slow, s.found - found0)) a.state[5] = 1ull;
}

/* RSI lengths doubled: dst[p] = src[p] + src[p + src[p]] for the candidates of the window */
__global__ void __launch_bounds__(SK_THREADS)
aec_skim_hdouble_kernel(const AecSkimArgs a, const uint32_t *src, uint32_t *dst)
{
    if (a.state[2] & 1ull) return;
    if (sk_sparse_now(a)) return;                       /* aec_skim_hdouble_list_kernel does the listed positions */
    const uint32_t sh = a.cfg.pad ? 3u : 0u;
    const uint32_t slots = (a.nh_eff + (1u << sh) - 1u) >> sh;
    /* a fixed grid that strides: when the window is sparse the launch costs next to nothing */
    for (uint32_t sl = blockIdx.x * SK_THREADS + threadIdx.x; sl < slots; sl += gridDim.x * SK_THREADS) {
        const uint32_t p = sl << sh;
        dst[p] = sk_hdouble(src, a.nh_eff, p);
    }
}

/* Sparse candidates: the length of eight RSIs in a row for the window's list of candidates that have an RSI
 * length, hop by hop through H (eight dependent look-ups per candidate, a million candidates in flight); the
 * buffer holds values at listed positions only (the walk trusts it only where H is not 0). */
__global__ void __launch_bounds__(SK_THREADS)
aec_skim_hchase_list_kernel(const AecSkimArgs a, uint32_t *dst)
{
    if (a.state[2] & 1ull) return;
    if (!sk_sparse_now(a)) return;
    uint64_t n = a.state[6 + a.set];                    /* sparse: bit 63 is clear */
    if (n > a.cand_cap) n = a.cand_cap;
    for (uint64_t i = (uint64_t)blockIdx.x * SK_THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * SK_THREADS) {
        const uint32_t p = a.cand_list[i];
        dst[p] = sk_hchase(a.H, a.nh_eff, p);
    }
}

/* the offsets the walk skipped over (heads of its long jumps know where they start) */
__global__ void __launch_bounds__(SK_THREADS)
aec_skim_fill_kernel(const AecSkimArgs a)
{
    const uint64_t r0 = a.state[4], r1 = a.state[1];
    for (uint64_t r = r0 + (uint64_t)blockIdx.x * SK_THREADS + threadIdx.x; r < r1; r += (uint64_t)gridDim.x * SK_THREADS)
        sk_fill(a.H, a.wb, a.offsets, r, r1);
}

/* Group index of the RSIs the walk has just taken from the tables: lane l of a warp finds where block l * G of
 * its RSI starts by the same descent through the levels (no second pass over the stream). */
__global__ void __launch_bounds__(SK_THREADS)
aec_skim_group_index_kernel(const AecSkimArgs a)
{
    const AecCfg &c = a.cfg;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t r0 = a.state[4], r1 = a.state[1];
    const uint64_t warps = (uint64_t)gridDim.x * (SK_THREADS / 32);
    for (uint64_t r = r0 + (uint64_t)blockIdx.x * (SK_THREADS / 32) + (threadIdx.x >> 5); r < r1; r += warps) {
        uint64_t *g = a.grp_index + r * 32ull;
        const uint64_t mark = g[0];
        __syncwarp();
        if (mark != SK_GRP_FAST) continue;              /* skimmed serially: the builder kernel does this RSI */
        const uint64_t start = a.offsets[r];
        const uint32_t p = (uint32_t)(start - a.wb);
        const uint32_t m = lane * a.grp_G;
        g[lane] = m < c.rsi ? sk_group_entry(c, a.T, a.R, a.LV, a.np, a.wb, p, m) : 0ull;
    }
}

} // namespace

uint32_t aec_skim_levels(const AecCfg &c) { return sk_levels(c); }
uint32_t aec_skim_sparse_min_levels(void) { return SK_SPARSE_MIN_LEVELS; }
uint64_t aec_skim_margin_bits(const AecCfg &c) { return sk_margin_bits(c); }

cudaError_t aec_skim_window_launch(const AecSkimArgs &args, cudaStream_t st)
{
    AecSkimArgs a = args;
    if (a.np == 0) return cudaSuccess;
    a.la_words = sk_lookahead_words(a.cfg);
    /* the bulk copy wants 16-byte aligned addresses on both sides */
    a.bulk = ((reinterpret_cast<uintptr_t>(a.in_words) & 15u) == 0 && (a.wb & 127ull) == 0) ? 1u : 0u;
    const uint32_t smem = (((SK_TILE / 32u + a.la_words + 1u + 3u) & ~3u) + SK_TILE / 32u + a.la_words + 1u) * 4u;
    if (a.sparse) aec_skim_begin_kernel<<<1, 1, 0, st>>>(a);
    aec_skim_level0_kernel<<<(a.np + SK_TILE - 1u) / SK_TILE, SK_THREADS, smem, st>>>(a);
    const uint32_t grid = (a.np / (uint32_t)SK_DP + SK_THREADS - 1u) / SK_THREADS;
    for (uint32_t j = 0; j + 1u < a.LV; j++)
        aec_skim_double_kernel<<<grid, SK_THREADS, 0, st>>>(a, j);
    const uint32_t cand = a.cfg.pad ? (a.nh_eff + 7u) / 8u : a.nh_eff;
    const uint32_t per_cta = SK_THREADS * SK_NC * SK_ROUNDS;
    /* candidates first; the dense pass only runs when the walk has asked for it (or sparse is off) */
    if (a.sparse) aec_skim_rsi_sparse_kernel<<<(a.nh_eff + SK_CHUNK - 1u) / SK_CHUNK, SK_THREADS, 0, st>>>(a);
    aec_skim_rsi_kernel<<<(cand + per_cta - 1u) / per_cta, SK_THREADS, 0, st>>>(a);
    if (a.H8) {
        /* dense tables: H -> 2 RSIs -> 4 -> 8, between two buffers, the last result lands in the second one;
         * sparse candidates: eight hops per listed candidate, straight into the second one.  Both sets of
         * kernels are launched; the window's latched mode lets one of them return at once. */
        uint32_t g2 = (cand + SK_THREADS - 1u) / SK_THREADS;
        if (g2 > 148u * 32u) g2 = 148u * 32u;
        if (a.sparse) {
            /* one look-up chain per listed candidate: enough threads for all of them to be in flight at once */
            uint32_t g3 = (a.cand_cap / 8u + SK_THREADS - 1u) / SK_THREADS;
            if (g3 > 4096u) g3 = 4096u;
            if (g3 < 1u) g3 = 1u;
            aec_skim_hchase_list_kernel<<<g3, SK_THREADS, 0, st>>>(a, a.H8 + a.np);
        }
        aec_skim_hdouble_kernel<<<g2, SK_THREADS, 0, st>>>(a, a.H, a.H8 + a.np);
        aec_skim_hdouble_kernel<<<g2, SK_THREADS, 0, st>>>(a, a.H8 + a.np, a.H8);
        aec_skim_hdouble_kernel<<<g2, SK_THREADS, 0, st>>>(a, a.H8, a.H8 + a.np);
    }
    return cudaGetLastError();
}

/* the serial part, on its own stream so that it runs next to the tables of the following window */
cudaError_t aec_skim_walk_launch(const AecSkimArgs &a, cudaStream_t st)
{
    if (a.np == 0) return cudaSuccess;
    aec_skim_walk_kernel<<<1, 32, 0, st>>>(a);
    if (a.H8) aec_skim_fill_kernel<<<64, SK_THREADS, 0, st>>>(a);
    if (a.grp_index) aec_skim_group_index_kernel<<<64, SK_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}
