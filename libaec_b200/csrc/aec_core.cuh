/*
 * aec_core.cuh -- per-block and per-RSI building blocks of the B200 AEC coder.
 *
 * Everything here is `__host__ __device__` so the very same code runs inside
 * the CUDA kernels (aec_encode.cu / aec_decode.cu) and inside the CPU model
 * harness (cpu_model.cpp) that tests/ uses to check the block logic against
 * the oracle without a GPU.  No reference code is used; citations point at the
 * reference lines whose *results* each function has to reproduce
 * (/root/reference/src/...).
 */
#ifndef AEC_CORE_CUH
#define AEC_CORE_CUH

#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define AEC_HD __host__ __device__ __forceinline__
#define AEC_HDM __host__ __device__ __forceinline__      /* member functions */
#define AEC_HDM_COLD __host__ __device__ __noinline__     /* rare paths kept out of the unrolled code */
#else
#define AEC_HD static inline
#define AEC_HDM inline
#define AEC_HDM_COLD inline
#endif

/* flag bits: same values as include/libaec.h */
#define AECF_SIGNED     1u
#define AECF_3BYTE      2u
#define AECF_MSB        4u
#define AECF_PREPROCESS 8u
#define AECF_RESTRICTED 16u
#define AECF_PAD_RSI    32u
#define AECF_NOT_ENFORCE 64u

#define AEC_MAX_J 64

/* code options of one block */
enum { OPT_ZERO = 0, OPT_SE = 1, OPT_SPLIT = 2, OPT_UNCOMP = 3, OPT_NONE = 4 };

/* Coding configuration derived once per stream (host) and passed by value. */
struct AecCfg {
    uint32_t n;       /* bits per sample 1..32 */
    uint32_t J;       /* block size (even, <= 64) */
    uint32_t rsi;     /* blocks per RSI */
    uint32_t flags;
    uint32_t B;       /* storage bytes per sample 1..4 */
    uint32_t idl;     /* id length */
    uint32_t kmax;    /* 2^idl - 3 */
    uint32_t pp;      /* preprocessing on */
    uint32_t msb;     /* MSB-first sample storage */
    uint32_t pad;     /* byte-align every RSI (AEC_PAD_RSI honoured) */
    uint32_t mask;    /* 2^n - 1 */
    uint32_t sflip;   /* 2^(n-1) for signed+pp, else 0: u = raw ^ sflip maps [xmin,xmax] -> [0,mask] */
    uint32_t sext;    /* signed: decoder sign-extends outputs (decode.c:78-84) */
    uint32_t R;       /* samples per RSI = rsi*J */
};

/* Validation + derivation shared by encoder and decoder
 * (results of encode.c:777-872 and decode.c:699-763). enc!=0 applies the
 * encoder-only checks. Returns 0 or AEC_CONF_ERROR (-1). */
AEC_HD int aec_cfg_init(AecCfg *c, uint32_t n, uint32_t J, uint32_t rsi, uint32_t flags,
                        int enc, int honour_pad)
{
    c->n = n; c->J = J; c->rsi = rsi; c->flags = flags;
    if (n == 0 || n > 32) return -1;
    if (enc) {
        if (flags & AECF_NOT_ENFORCE) { if (J & 1u) return -1; }
        else if (J != 8 && J != 16 && J != 32 && J != 64) return -1;
        if (rsi > 4096) return -1;
    }
    if (n > 16) { c->idl = 5; c->B = (n <= 24 && (flags & AECF_3BYTE)) ? 3 : 4; }
    else if (n > 8) { c->idl = 4; c->B = 2; }
    else {
        if (flags & AECF_RESTRICTED) {
            if (n <= 2) c->idl = 1; else if (n <= 4) c->idl = 2; else return -1;
        } else c->idl = 3;
        c->B = 1;
    }
    c->kmax = c->idl > 1 ? (1u << c->idl) - 3u : 0u;   /* no split option when idl <= 1 (encode.c:595-598) */
    c->pp = (flags & AECF_PREPROCESS) ? 1u : 0u;
    c->msb = (flags & AECF_MSB) ? 1u : 0u;
    c->pad = (honour_pad && (flags & AECF_PAD_RSI)) ? 1u : 0u;
    c->mask = (n == 32) ? 0xFFFFFFFFu : ((1u << n) - 1u);
    uint32_t sg = (flags & AECF_SIGNED) ? 1u : 0u;
    c->sflip = (sg && c->pp) ? (1u << (n - 1)) : 0u;
    c->sext = sg;
    c->R = rsi * J;
    return 0;
}

/* ---- small bit helpers with host equivalents ---- */
AEC_HD int aec_clz32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
AEC_HD int aec_clz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
AEC_HD uint32_t aec_bswap32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}
/* 32 bits starting `sh` bits into the 64-bit big-endian pair (hi:lo) */
AEC_HD uint32_t aec_funnel(uint32_t hi, uint32_t lo, uint32_t sh)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, sh);
#else
    sh &= 31u;
    return sh ? ((hi << sh) | (lo >> (32u - sh))) : hi;
#endif
}

/* x << s for s in 0..2^32-1 (zero when s >= 32) */
AEC_HD uint32_t aec_shl(uint32_t x, uint32_t s)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));
    return r;
#else
    return s < 32u ? x << s : 0u;
#endif
}

/* One sample from `B` storage bytes (results of encode_accessors.c:61-143). */
AEC_HD uint32_t aec_load_sample(const uint8_t *p, uint32_t B, uint32_t msb)
{
    uint32_t v = 0;
    if (msb) { for (uint32_t i = 0; i < B; i++) v = (v << 8) | p[i]; }
    else     { for (uint32_t i = 0; i < B; i++) v |= (uint32_t)p[i] << (8u * i); }
    return v;
}

/* CCSDS mapper for one sample given the previous one, both already normalised
 * to u = x - xmin in [0, M] (results of encode.c:255-269 / :294-309; the two
 * signedness variants collapse to this single unsigned form). */
AEC_HD uint32_t aec_map_delta(uint32_t u0, uint32_t u1, uint32_t M)
{
#if defined(__CUDA_ARCH__)
    const uint32_t D = __usad(u1, u0, 0u);                 /* |u1 - u0| in one instruction */
    const uint32_t zig = D + D - (u1 < u0 ? 1u : 0u);      /* only used when D <= th <= M/2: no overflow */
    const uint32_t th = min(u0, M - u0);
    return (D <= th) ? zig : th + D;
#else
    uint32_t ge = u1 >= u0;
    uint32_t D = ge ? (u1 - u0) : (u0 - u1);
    uint32_t th = (u0 < M - u0) ? u0 : (M - u0);
    return (D <= th) ? (2u * D - (ge ? 0u : 1u)) : (th + D);
#endif
}

/* Inverse of aec_map_delta (results of decode.c:89-135). */
AEC_HD uint32_t aec_unmap_delta(uint32_t u0, uint32_t d, uint32_t M)
{
    uint32_t h = (d >> 1) + (d & 1u);
    uint32_t mu = M - u0;
    uint32_t th = (u0 < mu) ? u0 : mu;
    uint32_t step = (d & 1u) ? (u0 - h) : (u0 + h);
    uint32_t clip = (u0 <= mu) ? d : (M - d);
    return (h <= th) ? step : clip;
}

/* ------------------------------------------------------------------------- */
/* Encoder: per-block analysis                                                */
/* ------------------------------------------------------------------------- */

struct BlockInfo {
    uint32_t opt;     /* OPT_* (OPT_ZERO = all-zero block, length decided by the run logic) */
    uint32_t klo, khi;/* argmin plateau of the split length (identity 0..kmax when unused) */
    uint32_t len;     /* CDS bits for SE/SPLIT/UNCOMP incl. id and reference sample */
};

/* Sum of d[i] >> k over the block. */
template <int JT>
AEC_HD uint32_t aec_sum_shift(const uint32_t *d, uint32_t J, uint32_t k)
{
    uint32_t s = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t i = 0; i < (JT ? (uint32_t)JT : J); i++) s += d[i] >> k;
    return s;
}

/*
 * Analyse one non-trivial block: pick the code option exactly as the
 * reference does (encode.c:585-612) and report the k plateau [klo,khi]
 * (encode.c:329-410 returns clamp(k_prev, klo, khi)).
 *   d[0..J)   mapped samples (d[0] == 0 in a reference block)
 *   ref       1 when the block carries the reference sample
 * Returns opt == OPT_ZERO when every sample is zero (caller runs the zero-run
 * logic); then klo/khi are the identity.
 */
/*   small   the caller knows every d[i] < 2^24 (then the 32-bit sum stands in for the OR of the values) */
/*   lut     device only, may be null: 256 words in shared memory, nibble j of lut[v] = bit j of v */
template <int JT>
AEC_HD BlockInfo aec_analyze_block(const AecCfg &c, const uint32_t *d, uint32_t ref, bool small = false,
                                   const uint32_t *lut = nullptr)
{
    const uint32_t J = JT ? (uint32_t)JT : c.J;
    BlockInfo bi;
    bi.klo = 0; bi.khi = c.kmax; bi.len = 0; bi.opt = OPT_ZERO;

    uint32_t orv = 0, s32 = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t i = 0; i < J; i++) s32 += d[i];
    if (small) orv = s32 ? 1u : 0u;          /* only "all zero", ">= 2^25" and ">= 2^31" are asked of orv */
    else {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t i = 0; i < J; i++) orv |= d[i];
    }
    if (orv == 0) return bi;
    uint64_t S0 = s32;
    if (orv >> 25) {       /* J <= 64 values below 2^25 cannot overflow 32 bits; otherwise add again in 64 */
        S0 = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t i = 0; i < J; i++) S0 += d[i];
    }

    const uint32_t thisbs = J - ref;
    const uint32_t unc = thisbs * c.n;                 /* encode.c:270, :746 */

    /* ---- split option: T(k) = S(k) - S(k+1) = sum ceil((d>>k)/2) is
     * non-increasing; klo = first k < kmax with T(k) <= thisbs, khi = first
     * k < kmax with T(k) < thisbs, both kmax when none. ---- */
    uint32_t split = 0xFFFFFFFFu;
    if (c.idl > 1) {
        /* guess kg: smallest k with (S0 >> (k+1)) <= thisbs; the true klo is
         * within one of it on every data set we measured */
        const uint32_t kmax = c.kmax;
        int kg = 0;
        if ((orv >> 25) == 0) {
            /* 32-bit sum: q >> kg0 is as long as thisbs when kg0 is the difference of their bit
             * lengths, so the answer is kg0 or kg0 + 1 */
            const uint32_t q = s32 >> 1;
            if (q > thisbs) {
                const int kg0 = aec_clz32(thisbs) - aec_clz32(q);
                kg = kg0 + ((q >> kg0) > thisbs ? 1 : 0);
            }
        } else {
            const uint64_t q = S0 >> 1;
            if (q > thisbs) {
                kg = (64 - aec_clz64(q)) - (32 - aec_clz32(thisbs));
                if (kg < 0) kg = 0;
                while (kg > 0 && (q >> (kg - 1)) <= thisbs) kg--;
                while ((q >> kg) > thisbs) kg++;
            }
        }
        uint32_t kgc = (uint32_t)kg > kmax ? kmax : (uint32_t)kg;
        uint32_t kb = kgc >= 2 ? kgc - 2 : 0;               /* window base: T known for kb..kb+3 */
        /* S at kb..kb+4; all sums fit 32 bits: S(kb) <= S0 >> kb <= 8*thisbs+7
         * when kb = kg-2, and d >> (kmax-2) <= 31 when the guess was clamped */
        uint32_t S1 = 0, S2 = 0, S3 = 0, S4 = 0, S5 = 0;
#if defined(__CUDA_ARCH__)
        if ((JT == 8 || JT == 16) && lut != nullptr && (orv >> 25) == 0 && (s32 >> kb) <= 255u) {
            /* Every d >> kb fits a byte (their sum does).  Count, per bit position j, the samples whose
             * shifted value has bit j set: C_j, one nibble each, accumulated from a table look-up per
             * sample (eight samples per accumulator, so a nibble holds its count).  Then
             * S(kb + m) = sum_{j >= m} 2^(j-m) C_j: ten byte dot products.  One shift per sample on the
             * ALU pipe, which bounds the kernel, instead of five shifts and five adds. */
            const uint32_t lb = (uint32_t)__cvta_generic_to_shared(lut);
            uint32_t acc[JT == 16 ? 2 : 1] = {0u};
#pragma unroll
            for (uint32_t i = 0; i < J; i++) {
                uint32_t w;
                asm("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"((d[i] >> kb) * 4u + lb));
                acc[i >> 3] += w;
            }
            uint32_t E = acc[0] & 0x0F0F0F0Fu, O = (acc[0] >> 4) & 0x0F0F0F0Fu;     /* bytes C0 C2 C4 C6 / C1 C3 C5 C7 */
            if (JT == 16) { E += acc[JT == 16 ? 1 : 0] & 0x0F0F0F0Fu; O += (acc[JT == 16 ? 1 : 0] >> 4) & 0x0F0F0F0Fu; }
            S1 = __dp4a(E, 0x40100401u, __dp4a(O, 0x80200802u, 0u));
            S2 = __dp4a(E, 0x20080200u, __dp4a(O, 0x40100401u, 0u));
            S3 = __dp4a(E, 0x10040100u, __dp4a(O, 0x20080200u, 0u));
            S4 = __dp4a(E, 0x08020000u, __dp4a(O, 0x10040100u, 0u));
            S5 = __dp4a(E, 0x04010000u, __dp4a(O, 0x08020000u, 0u));
        } else
#endif
        {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (uint32_t i = 0; i < J; i++) {
                uint32_t v = d[i] >> kb;
                S1 += v; S2 += v >> 1; S3 += v >> 2; S4 += v >> 3; S5 += v >> 4;
            }
        }
        const uint32_t T0 = S1 - S2, T1 = S2 - S3, T2 = S3 - S4, T3 = S4 - S5;
        uint32_t lo = 0xFFFFFFFFu, hi = 0xFFFFFFFFu, slo = 0;
        /* first k in the window with T <= thisbs (lo) and with T < thisbs (hi) */
        if (kb + 3 < kmax && T3 <= thisbs) { lo = kb + 3; slo = S4; }
        if (kb + 2 < kmax && T2 <= thisbs) { lo = kb + 2; slo = S3; }
        if (kb + 1 < kmax && T1 <= thisbs) { lo = kb + 1; slo = S2; }
        if (kb + 0 < kmax && T0 <= thisbs) { lo = kb + 0; slo = S1; }
        if (kb + 3 < kmax && T3 < thisbs) hi = kb + 3;
        if (kb + 2 < kmax && T2 < thisbs) hi = kb + 2;
        if (kb + 1 < kmax && T1 < thisbs) hi = kb + 1;
        if (kb + 0 < kmax && T0 < thisbs) hi = kb + 0;
        if (lo == kb && kb > 0) {
            /* left edge: smaller k may qualify as well -> walk down (rare) */
            uint32_t snext = S1, k = kb;
            while (k > 0) {
                uint32_t sk = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (uint32_t i = 0; i < J; i++) sk += d[i] >> (k - 1);
                uint32_t T = sk - snext;
                if (T > thisbs) break;
                k--; lo = k; slo = sk;
                if (T < thisbs) hi = k;
                snext = sk;
            }
        }
        if (hi == 0xFFFFFFFFu) {
            /* nothing strictly below thisbs inside the window -> walk up (rare) */
            uint32_t k = kb + 4, sk = S5;
            while (k < kmax) {
                uint32_t sn = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (uint32_t i = 0; i < J; i++) sn += d[i] >> (k + 1);
                uint32_t T = sk - sn;
                if (lo == 0xFFFFFFFFu && T <= thisbs) { lo = k; slo = sk; }
                if (T < thisbs) { hi = k; break; }
                sk = sn; k++;
            }
            if (hi == 0xFFFFFFFFu) hi = kmax;
        }
        if (lo == 0xFFFFFFFFu) {
            lo = kmax;
            uint32_t sk = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (uint32_t i = 0; i < J; i++) sk += d[i] >> kmax;
            slo = sk;
        }
        bi.klo = lo; bi.khi = hi;
        split = slo + thisbs * (lo + 1);
    }

    /* ---- second extension (encode.c:412-434) ---- */
    uint32_t se = 0xFFFFFFFFu;
    /* se >= 1 + J/2 + S0, so a larger S0 can never win -- except that the reference adds in u64
     * with wrap-around, reachable only when a pair sum reaches 2^32 */
    if (S0 <= unc || orv >= 0x80000000u) {
        uint64_t len = 1;
        bool inf = false;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t i = 0; i < J; i += 2) {
            uint64_t s = (uint64_t)d[i] + (uint64_t)d[i + 1];
            len += s * (s + 1) / 2 + d[i + 1] + 1;
            if (len > unc) { inf = true; break; }
        }
        if (!inf) se = (uint32_t)len;
    }

    /* ---- selection with the reference's tie-breaks (encode.c:600-611) ---- */
    uint32_t opt, body;
    if (split < unc) {
        if (split < se) { opt = OPT_SPLIT; body = split; }
        else            { opt = OPT_SE;    body = se; }
    } else {
        if (unc <= se)  { opt = OPT_UNCOMP; body = unc; }
        else            { opt = OPT_SE;     body = se; }
    }
    bi.opt = opt;
    /* the SE cost already counts the extra selector bit (its sum starts at 1, encode.c:424) */
    if (opt == OPT_SE || opt == OPT_SPLIT) bi.len = c.idl + ref * c.n + body;
    else                        bi.len = c.idl + J * c.n;       /* reference replaces sample 0 */
    return bi;
}

/* Zero-run bookkeeping for the block at position p of its 64-block segment.
 *   segmask  bit i set <=> block i of the segment is a valid all-zero block
 *   V        valid blocks in this segment (1..64)
 *   b        block index inside the RSI (p == b % 64)
 * Returns the CDS bit length this block owns (0 when it is not the last block
 * of its run) and fills *fs_code / *zref.  Results of encode.c:614-659 and
 * :565-583. */
AEC_HD uint32_t aec_zero_run(const AecCfg &c, uint64_t segmask, uint32_t V, uint32_t b,
                             uint32_t *fs_code, uint32_t *zref, uint32_t *runlen)
{
    uint32_t p = b & 63u;
    bool last = (p + 1 == V);
    bool next_zero = !last && ((segmask >> (p + 1)) & 1ull);
    *runlen = 0;
    if (next_zero) return 0;                           /* run continues */
    /* run length: consecutive ones ending at bit p */
    uint64_t sh = segmask << (63u - p);
    uint32_t L = (uint32_t)aec_clz64(~sh);
    if (L > p + 1) L = p + 1;
    *runlen = L;
    uint32_t code;
    if (last && L > 4) code = 4;                       /* ROS */
    else if (L >= 5)   code = L;
    else               code = L - 1;
    *fs_code = code;
    *zref = (c.pp && (b + 1 == L)) ? 1u : 0u;          /* run starts at block 0 of the RSI */
    return c.idl + 1 + (*zref ? c.n : 0) + code + 1;
}

/* clamp-map composition for the k chain: apply X then Y. */
AEC_HD uint32_t aec_clampu(uint32_t v, uint32_t lo, uint32_t hi)
{
    return v < lo ? lo : (v > hi ? hi : v);
}
/* pair layout: lo in bits 15..0, hi in bits 31..16 (two u16 lanes, so that the device
 * composes both bounds with one VIMNMX.U16x2 max and one min) */
AEC_HD uint32_t aec_kpair(uint32_t lo, uint32_t hi) { return lo | (hi << 16); }
AEC_HD uint32_t aec_klo(uint32_t p) { return p & 0xFFFFu; }
AEC_HD uint32_t aec_khi(uint32_t p) { return p >> 16; }
AEC_HD uint32_t aec_kapply(uint32_t k, uint32_t p) { return aec_clampu(k, aec_klo(p), aec_khi(p)); }
AEC_HD uint32_t aec_kcompose(uint32_t x, uint32_t y)
{
#if defined(__CUDA_ARCH__)
    const uint32_t ylo = __byte_perm(y, 0, 0x1010), yhi = __byte_perm(y, 0, 0x3232);
    return __vminu2(__vmaxu2(x, ylo), yhi);
#else
    uint32_t ylo = aec_klo(y), yhi = aec_khi(y);
    return aec_kpair(aec_clampu(aec_klo(x), ylo, yhi), aec_clampu(aec_khi(x), ylo, yhi));
#endif
}

/* Position monoid for AEC_PAD_RSI: f(p) = has_end ? roundup8(p + a) + rest : p + a. */
struct PosFn { uint32_t has_end; uint64_t a; uint64_t rest; };
AEC_HD uint64_t aec_up8(uint64_t x) { return (x + 7ull) & ~7ull; }
AEC_HD PosFn aec_pcompose(const PosFn &x, const PosFn &y)
{
    PosFn r;
    if (!y.has_end) {
        if (x.has_end) { r.has_end = 1; r.a = x.a; r.rest = x.rest + y.a; }
        else           { r.has_end = 0; r.a = x.a + y.a; r.rest = 0; }
    } else {
        if (x.has_end) { r.has_end = 1; r.a = x.a; r.rest = aec_up8(x.rest + y.a) + y.rest; }
        else           { r.has_end = 1; r.a = x.a + y.a; r.rest = y.rest; }
    }
    return r;
}
AEC_HD uint64_t aec_papply(const PosFn &f, uint64_t p)
{
    return f.has_end ? aec_up8(p + f.a) + f.rest : p + f.a;
}

/* ------------------------------------------------------------------------- */
/* Encoder: bit packer writing one CDS into a zero-initialised word buffer    */
/* ------------------------------------------------------------------------- */

/* The buffer holds big-endian-bit-order 32-bit words (bit 31 of word 0 is the
 * first stream bit).  A CDS shares its first and last word with its
 * neighbours, so those two are merged with OR (atomic on the GPU, where the
 * buffer is shared memory written by all threads of the CTA); interior words
 * are owned exclusively and stored plainly.
 *
 * The bits wait in a right-aligned 64-bit accumulator (hi:lo, `fill` valid
 * bits, fill < 32 between calls); a put shifts the field in and writes one
 * word out when 32 bits are complete.  On the device put() has neither a
 * data-dependent branch nor an atomic: the first completed word goes to a
 * slot private to the thread (`side`), every later one to the staging word it
 * owns, all with one predicated store; finish() merges the first word and the
 * unfinished last one into the staging area. */
struct BitPack {
    uint32_t lo, hi;  /* accumulator, valid bits [0, fill) */
    uint32_t fill;
    uint32_t wfirst;  /* the CDS's first word (shared with the previous CDS) */
#if defined(__CUDA_ARCH__)
    uint32_t wst;     /* shared-window byte address the next completed word is stored to */
    uint32_t wnext;   /* address of the staging word after the one being filled */
#else
    uint32_t *buf;
    uint32_t wcur;    /* word index of the word being filled */
#endif

    /* b: staging area, bitpos: where the CDS starts in it; side (device only): shared-window address of a
     * word private to this thread */
    AEC_HDM void init(uint32_t *b, uint32_t bitpos, uint32_t side)
    {
        lo = 0; hi = 0; fill = bitpos & 31u;
#if defined(__CUDA_ARCH__)
        wfirst = (uint32_t)__cvta_generic_to_shared(b) + ((bitpos >> 5) << 2);
        wst = side; wnext = wfirst + 4u;
#else
        (void)side;
        buf = b; wcur = bitpos >> 5; wfirst = wcur;
#endif
    }
    /* append len (0..32) bits of v (v < 2^len) */
    AEC_HDM void put(uint32_t v, uint32_t len)
    {
#if defined(__CUDA_ARCH__)
        hi = __funnelshift_lc(lo, hi, len);
        asm("shl.b32 %0, %1, %2;" : "=r"(lo) : "r"(lo), "r"(len));     /* PTX shifts clamp: len == 32 gives 0 */
        lo |= v;
        fill += len;
        /* predicated, not branched: whether a put completes a word differs from lane to lane */
        asm volatile("{\n\t"
                     ".reg .pred p;\n\t"
                     ".reg .b32 w;\n\t"
                     "setp.ge.u32 p, %0, 32;\n\t"
                     "shf.r.wrap.b32 w, %3, %4, %0;\n\t"            /* bits [fill-32, fill) */
                     "@p st.shared.b32 [%1], w;\n\t"
                     "@p mov.b32 %1, %2;\n\t"
                     "@p add.u32 %2, %2, 4;\n\t"
                     "and.b32 %0, %0, 31;\n\t"
                     "}"
                     : "+r"(fill), "+r"(wst), "+r"(wnext) : "r"(lo), "r"(hi) : "memory");
#else
        uint64_t acc = ((uint64_t)hi << 32) | lo;
        acc = (len >= 32u ? (acc << 16) << 16 : acc << len) | v;
        fill += len;
        if (fill >= 32u) {
            const uint32_t w = (uint32_t)(acc >> (fill - 32u));
            if (wcur == wfirst) buf[wcur] |= w; else buf[wcur] = w;
            wcur++;
            fill -= 32u;
        }
        lo = (uint32_t)acc; hi = (uint32_t)(acc >> 32);
#endif
    }
    /* fundamental sequence: fs zeros then a one (the loop only runs for codes longer than 32 bits) */
    AEC_HDM void put_fs(uint32_t fs)
    {
        while (fs >= 32u) { put(0u, 32u); fs -= 32u; }
        put(1u, fs + 1u);
    }
    AEC_HDM void finish(uint32_t side)
    {
#if defined(__CUDA_ARCH__)
        if (wnext != wfirst + 4u) {                 /* at least one word was completed: the first sits in the side slot */
            uint32_t h;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(h) : "r"(side) : "memory");
            asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(wfirst), "r"(h) : "memory");
        }
        if (fill) asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(wnext - 4u), "r"(lo << (32u - fill)) : "memory");
#else
        (void)side;
        if (fill) buf[wcur] |= lo << (32u - fill);
#endif
    }
};

/*
 * Emit the CDS of one non-zero block (results of encode.c:520-563).
 *   d      mapped samples (d[0] == 0 in a reference block); refs: raw reference sample when ref
 *   k      split position (already clamp(k_prev, klo, khi))
 * Fields are combined four samples at a time before they go through the
 * word packer: a group's unary codes (or its k-bit remainders) usually fit
 * one 32-bit field.
 */
template <int JT>
AEC_HD void aec_pack_block(const AecCfg &c, BitPack &bp, const uint32_t *d, uint32_t opt,
                           uint32_t k, uint32_t ref, uint32_t refs)
{
    const uint32_t J = JT ? (uint32_t)JT : c.J;
    constexpr int NQ = (JT ? JT : AEC_MAX_J) / 4 + ((JT ? JT : AEC_MAX_J) % 4 ? 1 : 0);
    if (opt == OPT_SPLIT) {
        bp.put(k + 1, c.idl);
        if (ref) bp.put(refs, c.n);
        /* unary part: the codes of four samples as one field when they fit 32 bits */
        uint32_t qa[NQ], ql[NQ], lmax = 0;
        if (JT == 0) { for (int q = 0; q < NQ; q++) { qa[q] = 0; ql[q] = 0; } }   /* quads beyond J stay empty */
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t g = 0; g < J; g += 4) {
            uint32_t acc = 0, len = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (uint32_t j = 0; j < 4; j++) {
                const uint32_t i = g + j;
                if (i >= J) continue;
                uint32_t s1 = (d[i] >> k) + 1u, one = 1u;
                if (i == 0) { s1 -= ref; one -= ref; }        /* d[0] == 0 in a reference block: nothing to emit */
                acc = aec_shl(acc, s1) | one;
                len += s1;                                     /* a chosen option has sum(fs) < J*n: no overflow */
            }
            qa[g >> 2] = acc; ql[g >> 2] = len;
            lmax = len > lmax ? len : lmax;
        }
        /* two quads as one field when every such pair fits 32 bits (the usual case: a put costs
         * about as much as packing three samples) */
        uint32_t pmax = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q + 1 < NQ; q += 2) { const uint32_t l2 = ql[q] + ql[q + 1]; pmax = l2 > pmax ? l2 : pmax; }
        if (NQ > 1 && pmax <= 32u && lmax <= 32u) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int q = 0; q < NQ; q += 2) {
                if ((uint32_t)q * 4u >= J) continue;
                if (q + 1 < NQ) bp.put(aec_shl(qa[q], ql[q + 1]) | qa[q + 1], ql[q] + ql[q + 1]);
                else bp.put(qa[q], ql[q]);
            }
        } else if (lmax <= 32u) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (uint32_t g = 0; g < J; g += 4) bp.put(qa[g >> 2], ql[g >> 2]);
        } else {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (uint32_t i = 0; i < J; i++)
                if (i >= ref) bp.put_fs(d[i] >> k);
        }
        /* binary part: k low bits of every sample */
        if (k) {
            const uint32_t m = (1u << k) - 1u;
            if (k <= 8) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (uint32_t g = 0; g < J; g += 4) {
                    uint32_t acc = 0, len = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                    for (uint32_t j = 0; j < 4; j++) {
                        const uint32_t i = g + j;
                        if (i >= J) continue;
#if defined(__CUDA_ARCH__)
                        acc = acc * (m + 1u) + (d[i] & m);        /* the shift as a multiply-add: FMA pipe */
#else
                        acc = (acc << k) | (d[i] & m);
#endif
                        len += k;
                    }
                    if (g == 0) len -= ref * k;
                    bp.put(acc, len);
                }
            } else if (k <= 16) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (uint32_t g = 0; g < J; g += 2) {
#if defined(__CUDA_ARCH__)
                    const uint32_t acc = (d[g] & m) * (m + 1u) + (d[g + 1] & m);
#else
                    const uint32_t acc = ((d[g] & m) << k) | (d[g + 1] & m);
#endif
                    bp.put(acc, (g == 0 && ref) ? k : 2u * k);
                }
            } else {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (uint32_t i = 0; i < J; i++) bp.put(d[i] & m, (i == 0 && ref) ? 0u : k);
            }
        }
    } else if (opt == OPT_SE) {
        bp.put(1, c.idl + 1);
        if (ref) bp.put(refs, c.n);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t i = 0; i < J; i += 2) {
            uint32_t s = d[i] + d[i + 1];
            bp.put_fs(s * (s + 1) / 2 + d[i + 1]);      /* u32 like encode.c:558-559 */
        }
    } else { /* OPT_UNCOMP */
        bp.put((1u << c.idl) - 1u, c.idl);
        bp.put(ref ? refs : d[0], c.n);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t i = 1; i < J; i++) bp.put(d[i], c.n);
    }
}

/* Zero-run CDS (results of encode.c:565-583). */
AEC_HD void aec_pack_zero(const AecCfg &c, BitPack &bp, uint32_t fs_code, uint32_t zref, uint32_t refs)
{
    bp.put(0, c.idl + 1);
    if (zref) bp.put(refs, c.n);
    bp.put_fs(fs_code);
}

/* ------------------------------------------------------------------------- */
/* Decoder: big-endian bit reader over 32-bit words                           */
/* ------------------------------------------------------------------------- */

struct BitRd {
    const uint32_t *w;   /* 4-byte aligned, raw (byte order as stored) */
    uint64_t nwords;     /* words that may be read; beyond that zeros */
    uint64_t nbits;      /* valid stream bits */
    uint64_t ci;         /* index of cached word pair */
    uint32_t c0, c1;

    AEC_HDM uint32_t word(uint64_t i) const { return i < nwords ? aec_bswap32(w[i]) : 0u; }
    AEC_HDM void init(const uint32_t *base, uint64_t nw, uint64_t nb)
    {
        w = base; nwords = nw; nbits = nb; ci = 0xFFFFFFFFFFFFFFF0ull; c0 = c1 = 0;
    }
    /* 32 stream bits starting at bit position pos (zeros past the end) */
    AEC_HDM uint32_t peek(uint64_t pos)
    {
        uint64_t i = pos >> 5;
        if (i != ci) {
            if (i == ci + 1) { c0 = c1; c1 = word(i + 1); }
            else { c0 = word(i); c1 = word(i + 1); }
            ci = i;
        }
        return aec_funnel(c0, c1, (uint32_t)(pos & 31u));
    }
};

#endif /* AEC_CORE_CUH */
