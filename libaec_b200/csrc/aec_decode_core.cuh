/*
 * aec_decode_core.cuh -- per-RSI decoding steps shared by the CUDA decode
 * kernels and the CPU model harness (host+device, see aec_core.cuh).
 *
 * One *lane* owns one RSI and walks it block by block.  Each call of
 * aec_decode_block() parses (or continues) one block and leaves its J mapped
 * values in a caller-provided row; the caller undoes the predictor and stores
 * the samples.  Results reproduce /root/reference/src/decode.c (states m_id,
 * m_split, m_zero_block, m_se, m_uncomp, m_low_entropy*), restated for a
 * position-addressed bit reader instead of a byte-fed accumulator.
 */
#ifndef AEC_DECODE_CORE_CUH
#define AEC_DECODE_CORE_CUH

#include "aec_core.cuh"

enum { DEC_OK = 0, DEC_EXHAUSTED = 1, DEC_ERROR = 2 };

struct RsiDec {
    uint64_t pos;        /* bit cursor */
    uint32_t zero_left;  /* all-zero blocks still owed by the current zero run */
    uint32_t status;     /* DEC_* */
};

/* Read `len` (1..32) bits at st.pos; false when they are not all inside the stream. */
AEC_HD bool aec_rd_bits(BitRd &br, RsiDec &st, uint32_t len, uint32_t *v)
{
    if (st.pos + len > br.nbits) return false;
    uint32_t wnd = br.peek(st.pos);
    *v = wnd >> (32u - len);
    st.pos += len;
    return true;
}

/* Fundamental sequence at st.pos (decode.c:288-340 / :361-379). */
AEC_HD bool aec_rd_fs(BitRd &br, RsiDec &st, uint32_t *fs)
{
    uint32_t acc = 0;
    uint64_t p = st.pos;
    for (;;) {
        if (p >= br.nbits) return false;
        uint32_t wnd = br.peek(p);
        if (wnd == 0) { acc += 32; p += 32; continue; }
        uint32_t z = (uint32_t)aec_clz32(wnd);
        acc += z; p += z + 1;
        break;
    }
    if (p > br.nbits) return false;   /* terminator lies in the zero padding: cannot happen, kept for safety */
    *fs = acc; st.pos = p;
    return true;
}

/*
 * Decode block `b` of the lane's RSI into row[0..J).  row[] receives mapped
 * values d (row[0] is the raw reference sample when the block carries one).
 * Returns how many leading values of the row are valid: J normally, fewer when
 * the stream ran out (st.status = DEC_EXHAUSTED) or is corrupt (DEC_ERROR).
 */
template <int JT>
AEC_HD uint32_t aec_decode_block(const AecCfg &c, BitRd &br, RsiDec &st, uint32_t b, uint32_t *row)
{
    const uint32_t J = JT ? (uint32_t)JT : c.J;
    if (st.zero_left) {                                   /* inside a zero run */
        st.zero_left--;
        for (uint32_t i = 0; i < J; i++) row[i] = 0;
        return J;
    }
    const uint32_t ref = (c.pp && b == 0) ? 1u : 0u;
    uint32_t id, v;
    if (!aec_rd_bits(br, st, c.idl, &id)) { st.status = DEC_EXHAUSTED; return 0; }

    if (id == 0) {                                        /* low entropy (decode.c:634-644) */
        uint32_t sel;
        if (!aec_rd_bits(br, st, 1, &sel)) { st.status = DEC_EXHAUSTED; return 0; }
        uint32_t produced = 0;
        if (ref) {
            if (!aec_rd_bits(br, st, c.n, &v)) { st.status = DEC_EXHAUSTED; return 0; }
            row[0] = v; produced = 1;
        }
        if (sel == 0) {                                   /* zero run (decode.c:518-558) */
            uint32_t fs;
            if (!aec_rd_fs(br, st, &fs)) { st.status = DEC_EXHAUSTED; return produced; }
            uint32_t zb = fs + 1;
            if (zb == 5) {                                /* ROS */
                uint32_t a1 = c.rsi - b, a2 = 64u - (b & 63u);
                zb = a1 < a2 ? a1 : a2;
            } else if (zb > 5) {
                zb--;
            }
            if (zb > c.rsi - b) { st.status = DEC_ERROR; return produced; }   /* decode.c:543-544 */
            st.zero_left = zb - 1;
            for (uint32_t i = ref; i < J; i++) row[i] = 0;
            return J;
        }
        /* second extension (decode.c:589-616) */
        uint32_t i = ref;
        while (i < J) {
            uint32_t m;
            if (!aec_rd_fs(br, st, &m)) { st.status = DEC_EXHAUSTED; return i; }
            /* s = max{s : s(s+1)/2 <= m}; valid encoder output has m <= 90 */
            uint32_t s = 0;
            if (m > 90) { st.status = DEC_ERROR; return i; }
            while ((s + 1) * (s + 2) / 2 <= m) s++;
            uint32_t d1 = m - s * (s + 1) / 2;
            if ((i & 1u) == 0) { row[i] = s - d1; i++; }
            row[i] = d1; i++;
        }
        return J;
    }
    if (id == (1u << c.idl) - 1u) {                       /* uncompressed (decode.c:659-677) */
        for (uint32_t i = 0; i < J; i++) {
            if (!aec_rd_bits(br, st, c.n, &v)) { st.status = DEC_EXHAUSTED; return i; }
            row[i] = v;
        }
        return J;
    }
    /* split, k = id - 1 (decode.c:462-502) */
    const uint32_t k = id - 1;
    if (ref) {
        if (!aec_rd_bits(br, st, c.n, &v)) { st.status = DEC_EXHAUSTED; return 0; }
        row[0] = v;
    }
    for (uint32_t i = ref; i < J; i++) {
        uint32_t fs;
        if (!aec_rd_fs(br, st, &fs)) { st.status = DEC_EXHAUSTED; return ref; }
        row[i] = fs << k;
    }
    if (k) {
        for (uint32_t i = ref; i < J; i++) {
            if (!aec_rd_bits(br, st, k, &v)) { st.status = DEC_EXHAUSTED; return i; }
            row[i] += v;
        }
    }
    return J;
}

/*
 * Skim one CDS without producing values: advances st.pos / st.zero_left the
 * same way aec_decode_block does.  Used by the sequential RSI-boundary scan for
 * streams that come without an offset index.
 */
AEC_HD bool aec_skim_block(const AecCfg &c, BitRd &br, RsiDec &st, uint32_t b)
{
    const uint32_t J = c.J;
    if (st.zero_left) { st.zero_left--; return true; }
    const uint32_t ref = (c.pp && b == 0) ? 1u : 0u;
    uint32_t id, v, fs;
    if (!aec_rd_bits(br, st, c.idl, &id)) { st.status = DEC_EXHAUSTED; return false; }
    if (id == 0) {
        uint32_t sel;
        if (!aec_rd_bits(br, st, 1, &sel)) { st.status = DEC_EXHAUSTED; return false; }
        if (ref && !aec_rd_bits(br, st, c.n, &v)) { st.status = DEC_EXHAUSTED; return false; }
        if (sel == 0) {
            if (!aec_rd_fs(br, st, &fs)) { st.status = DEC_EXHAUSTED; return false; }
            uint32_t zb = fs + 1;
            if (zb == 5) { uint32_t a1 = c.rsi - b, a2 = 64u - (b & 63u); zb = a1 < a2 ? a1 : a2; }
            else if (zb > 5) zb--;
            if (zb > c.rsi - b) { st.status = DEC_ERROR; return false; }
            st.zero_left = zb - 1;
            return true;
        }
        for (uint32_t i = ref; i < J; i += (i & 1u) ? 1u : 2u)
            if (!aec_rd_fs(br, st, &fs)) { st.status = DEC_EXHAUSTED; return false; }
        return true;
    }
    if (id == (1u << c.idl) - 1u) {
        uint64_t need = (uint64_t)J * c.n;
        if (st.pos + need > br.nbits) { st.status = DEC_EXHAUSTED; return false; }
        st.pos += need;
        return true;
    }
    const uint32_t k = id - 1;
    if (ref && !aec_rd_bits(br, st, c.n, &v)) { st.status = DEC_EXHAUSTED; return false; }
    /* skip (J-ref) terminators: count ones 32 bits at a time */
    uint32_t need = J - ref;
    uint64_t p = st.pos;
    while (need) {
        if (p >= br.nbits) { st.status = DEC_EXHAUSTED; return false; }
        uint32_t wnd = br.peek(p);
#if defined(__CUDA_ARCH__)
        uint32_t ones = (uint32_t)__popc(wnd);
#else
        uint32_t ones = (uint32_t)__builtin_popcount(wnd);
#endif
        if (ones < need) { need -= ones; p += 32; continue; }
        /* the need-th one is inside this window: peel from the top */
        while (need > 1) { wnd &= ~(0x80000000u >> aec_clz32(wnd)); need--; }
        p += (uint32_t)aec_clz32(wnd) + 1;
        need = 0;
    }
    if (p > br.nbits) { st.status = DEC_EXHAUSTED; return false; }
    uint64_t lsb = (uint64_t)(J - ref) * k;
    if (p + lsb > br.nbits) { st.status = DEC_EXHAUSTED; return false; }
    st.pos = p + lsb;
    return true;
}

/* Sample -> storage bytes (results of decode.c:144-189). */
AEC_HD void aec_store_sample(uint8_t *o, uint32_t v, uint32_t B, uint32_t msb)
{
    if (msb) { for (uint32_t i = 0; i < B; i++) o[i] = (uint8_t)(v >> (8u * (B - 1u - i))); }
    else     { for (uint32_t i = 0; i < B; i++) o[i] = (uint8_t)(v >> (8u * i)); }
}

/* Undo the predictor over one row in place: row[] holds d (row[0] the raw
 * reference when first), on return the output words to be stored (low 8*B
 * bits matter).  `u_prev` carries the normalised previous sample across
 * blocks. Results of decode.c:67-141. */
AEC_HD void aec_unmap_row(const AecCfg &c, uint32_t *row, uint32_t cnt, uint32_t first_is_ref,
                          uint32_t *u_prev)
{
    if (!c.pp) return;                                    /* decode.c:136-139 */
    uint32_t u = *u_prev;
    const uint32_t sflip = c.sext ? (1u << (c.n - 1)) : 0u;
    for (uint32_t i = 0; i < cnt; i++) {
        if (i == 0 && first_is_ref) u = (row[0] ^ sflip) & c.mask;   /* reference sample */
        else u = aec_unmap_delta(u, row[i], c.mask);
        uint32_t x = u ^ sflip;                           /* back to the n-bit pattern */
        if (c.sext && c.n < 32 && (x >> (c.n - 1)) & 1u) x |= ~c.mask;   /* sign-extend (decode.c:78-84, :131) */
        row[i] = x;
    }
    *u_prev = u;
}

#endif /* AEC_DECODE_CORE_CUH */
