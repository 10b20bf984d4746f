/*
 * aec_skim_core.cuh -- building blocks of the parallel RSI-boundary discovery (aec_skim.cu), host+device
 * so that the CPU model harness (cpu_model.cpp) runs the very same code against the oracle.
 *
 * Table entry (32 bits): bits 31..12 = length in bits of a piece of the CDS chain (0 = nothing valid
 * starts here), bits 11..0 = blocks it stands for; a length with 0 blocks is a run-of-zero-segment CDS,
 * whose block count depends on the block number (decode.c:528-530).
 */
#ifndef AEC_SKIM_CORE_CUH
#define AEC_SKIM_CORE_CUH

#include "aec_decode_core.cuh"

#define SK_FAIL 0xFFFFFFFFu
#define SK_MAX_LEVELS 8

AEC_HD uint32_t sk_len(uint32_t e) { return e >> 12; }
AEC_HD uint32_t sk_blk(uint32_t e) { return e & 0xFFFu; }
AEC_HD bool sk_jump(uint32_t e) { return e >= 0x1000u && (e & 0xFFFu) != 0u; }
AEC_HD bool sk_ros(uint32_t e) { return e >= 0x1000u && (e & 0xFFFu) == 0u; }

AEC_HD uint32_t sk_popc(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}

/* position (0 = most significant bit) of the r-th set bit of x counted from the top, 1 <= r <= popc(x) */
AEC_HD uint32_t sk_select_from_top(uint32_t x, uint32_t r)
{
    uint32_t pos = 0, c;
    bool lo;
    c = sk_popc(x >> 16); lo = r > c; r -= lo ? c : 0u; pos += lo ? 16u : 0u; x = lo ? (x << 16) : x;
    c = sk_popc(x >> 24); lo = r > c; r -= lo ? c : 0u; pos += lo ? 8u : 0u;  x = lo ? (x << 8) : x;
    c = sk_popc(x >> 28); lo = r > c; r -= lo ? c : 0u; pos += lo ? 4u : 0u;  x = lo ? (x << 4) : x;
    c = sk_popc(x >> 30); lo = r > c; r -= lo ? c : 0u; pos += lo ? 2u : 0u;  x = lo ? (x << 2) : x;
    c = x >> 31; pos += (r > c) ? 1u : 0u;
    return pos;
}

/* tile-relative position of the m-th one at or after tile-relative bit q (m >= 1); SK_FAIL when it is not
 * inside the staged words.  w[0..nwords]: big-endian words, pre[i] = ones in w[0..i), i <= nwords */
AEC_HD uint32_t sk_find_one(const uint32_t *w, const uint32_t *pre, uint32_t nwords, uint32_t q, uint32_t m)
{
    uint32_t i = q >> 5;
    if (i >= nwords) return SK_FAIL;
    const uint32_t first = w[i] & (0xFFFFFFFFu >> (q & 31u));
    const uint32_t c = sk_popc(first);
    if (m <= c) return 32u * i + sk_select_from_top(first, m);
    i++;
    const uint32_t target = pre[i] + (m - c);
    while (i < nwords && pre[i + 1] < target) i++;
    if (i >= nwords) return SK_FAIL;
    return 32u * i + sk_select_from_top(w[i], target - pre[i]);
}

/* The CDS that would start at tile-relative bit q0 (decode.c:402-677 restated as "how long, how many
 * blocks"); ref != 0: it is the first CDS of an RSI and carries the reference sample.  limit = stream
 * bits that exist, tile-relative.  q0 < 32 * (nwords - 1). */
AEC_HD uint32_t sk_entry(const AecCfg &c, const uint32_t *w, const uint32_t *pre, uint32_t nwords,
                         uint32_t q0, uint32_t limit, uint32_t ref)
{
    const uint32_t i = q0 >> 5;
    const uint32_t win = aec_funnel(w[i], w[i + 1], q0 & 31u);
    const uint32_t id = win >> (32u - c.idl);
    const uint32_t refb = ref ? c.n : 0u;
    uint32_t q, m, extra = 0;
    bool zero = false;
    if (id == 0u) {
        const uint32_t sel = (win >> (31u - c.idl)) & 1u;
        q = q0 + c.idl + 1u + refb;
        m = sel ? (c.J >> 1) : 1u;                      /* second extension: J/2 codes; zero run: one */
        zero = sel == 0u;
    } else if (id == (1u << c.idl) - 1u) {
        const uint32_t len = c.idl + c.J * c.n;         /* the reference sample takes the place of sample 0 */
        return (q0 + len <= limit) ? ((len << 12) | 1u) : 0u;
    } else {
        const uint32_t k = id - 1u;
        q = q0 + c.idl + refb;
        m = c.J - (ref ? 1u : 0u);
        extra = m * k;
    }
    const uint32_t e = sk_find_one(w, pre, nwords, q, m);
    if (e == SK_FAIL) return 0u;
    const uint32_t end = e + 1u + extra;
    if (end > limit) return 0u;
    uint32_t blocks = 1u;
    if (zero) {
        const uint32_t zb = e - q + 1u;                 /* decode.c:525-533 */
        blocks = zb == 5u ? 0u : (zb > 5u ? zb - 1u : zb);
        if (blocks > 0xFFFu) return 0u;
    }
    return ((end - q0) << 12) | blocks;
}

/* one pointer-doubling step: the chain from p after twice as many CDSs */
AEC_HD uint32_t sk_double(const uint32_t *src, uint32_t np, uint32_t p)
{
    const uint32_t x = src[p];
    if (!sk_jump(x)) return 0u;
    const uint32_t q = p + sk_len(x);
    if (q >= np) return 0u;
    const uint32_t y = src[q];
    if (!sk_jump(y)) return 0u;
    const uint32_t blk = sk_blk(x) + sk_blk(y), len = sk_len(x) + sk_len(y);
    return (blk <= 0xFFFu && len <= 0xFFFFFu) ? ((len << 12) | blk) : 0u;
}

/* Bits from window-relative p to the start of the next RSI when an RSI starts at p (0: the tables cannot
 * tell).  first = entry of the RSI's first CDS; T[LV][np] = the chain tables. */
AEC_HD uint32_t sk_rsi_len(const AecCfg &c, const uint32_t *T, uint32_t LV, uint32_t np, uint32_t p, uint32_t first)
{
    if (first < 0x1000u) return 0u;
    uint32_t b = sk_blk(first);
    if (b == 0u) b = c.rsi < 64u ? c.rsi : 64u;        /* run-of-zero-segment at block 0 */
    uint32_t q = p + sk_len(first);
    if (b > c.rsi) return 0u;
    while (b < c.rsi) {
        uint32_t rem = c.rsi - b;
        for (int j = (int)LV - 1; j >= 0; j--) {
            const uint32_t *Tj = T + (size_t)j * np;
            for (;;) {
                if (q >= np || (1u << j) > rem) break;  /* 2^j CDSs stand for 2^j blocks at least: no need to look */
                const uint32_t e = Tj[q];
                if (!sk_jump(e) || sk_blk(e) > rem) break;
                q += sk_len(e); rem -= sk_blk(e);
                if (j != (int)LV - 1) break;            /* below the top level a step fits at most once */
            }
        }
        b = c.rsi - rem;
        if (rem == 0u) break;
        /* the CDS at q is not a plain step: run-of-zero-segment, or the chain ends here */
        if (q >= np) return 0u;
        const uint32_t e = T[q];
        if (sk_ros(e)) {
            const uint32_t seg = 64u - (b & 63u);
            b += rem < seg ? rem : seg;                 /* decode.c:528-530 */
            q += sk_len(e);
        } else if (sk_jump(e) && sk_blk(e) <= rem) {
            q += sk_len(e); b += sk_blk(e);
        } else return 0u;
    }
    if (c.pad) q = (q + 7u) & ~7u;                      /* windows start on byte boundaries */
    return q - p;
}

#define SK_GRP_POISON 0x00FFFFFFFFFFFFFFull     /* an entry the fast decoder refuses (it hands the RSI to the careful kernel) */
#define SK_GRP_FAST    0xFFFFFFFFFFFFFFFEull     /* walk -> group kernel: this RSI came from the tables, its entries are still to be written */
#define SK_GRP_MISSING 0xFFFFFFFFFFFFFFFFull     /* walk -> builder: this RSI was skimmed serially, build its entries by skimming */

/* Group index entry (aec_device.h: AecDecArgs::grp_index) of the group that starts at block m of the RSI at
 * window-relative p, from the chain tables: bits 55..0 absolute bit offset of the first CDS the lane parses,
 * bits 63..56 blocks at the head of the group that still belong to a zero run begun before it.  R[p] = entry
 * of the RSI's first CDS, wb = the window's first bit. */
AEC_HD uint64_t sk_group_entry(const AecCfg &c, const uint32_t *T, const uint32_t *R, uint32_t LV, uint32_t np, uint64_t wb,
                               uint32_t p, uint32_t m)
{
    if (m == 0u) return wb + p;
    const uint32_t first = R[p];
    if (first < 0x1000u) return SK_GRP_POISON;
    uint32_t c0 = sk_blk(first);
    if (c0 == 0u) c0 = c.rsi < 64u ? c.rsi : 64u;
    uint32_t q = p + sk_len(first);
    if (c0 > m) return ((uint64_t)(c0 - m) << 56) | (wb + q);      /* the group starts inside the first CDS's zero run */
    uint32_t rem = m - c0;
    for (int guard = 0; rem != 0u && guard < 8192; guard++) {
        for (int j = (int)LV - 1; j >= 0; j--) {
            const uint32_t *Tj = T + (size_t)j * np;
            for (;;) {
                if (q >= np || (1u << j) > rem) break;
                const uint32_t e = Tj[q];
                if (!sk_jump(e) || sk_blk(e) > rem) break;
                q += sk_len(e); rem -= sk_blk(e);
                if (j != (int)LV - 1) break;
            }
        }
        if (rem == 0u) break;
        if (q >= np) return SK_GRP_POISON;
        const uint32_t e = T[q];
        const uint32_t b = m - rem;                     /* block number of the CDS at q */
        uint32_t cnt;
        if (sk_ros(e)) { const uint32_t seg = 64u - (b & 63u), left = c.rsi - b; cnt = left < seg ? left : seg; }
        else if (sk_jump(e)) cnt = sk_blk(e);
        else return SK_GRP_POISON;
        if (cnt > rem) return ((uint64_t)(cnt - rem) << 56) | (wb + q + sk_len(e));   /* a zero run straddles the group start */
        q += sk_len(e); rem -= cnt;
    }
    return rem == 0u ? wb + q : SK_GRP_POISON;
}

/* levels of the chain tables: level j = 2^j CDSs; 2^LV - 1 >= rsi - 1 up to the cap */
AEC_HD uint32_t sk_levels(const AecCfg &c)
{
    uint32_t lv = 1;
    while (lv < SK_MAX_LEVELS && (1u << lv) < c.rsi) lv++;
    return lv;
}

/* bits an RSI can take at most, in whole words, plus slack */
AEC_HD uint64_t sk_margin_bits(const AecCfg &c)
{
    const uint64_t cds = c.idl + 1ull + (uint64_t)(c.J + 1u) * c.n;
    return ((cds * c.rsi + 8ull + 31ull) & ~31ull) + 64ull;
}

/* words staged beyond a tile: the longest CDS plus the funnel shift's neighbour */
AEC_HD uint32_t sk_lookahead_words(const AecCfg &c)
{
    return (uint32_t)((c.idl + 1ull + (uint64_t)(c.J + 1u) * c.n + 31ull) / 32ull) + 2u;
}

#define SK_OFF_PENDING 0xFFFFFFFFFFFFFFFFull    /* walk -> fill: this RSI's offset follows from its predecessor's H */
#define SK_SKIP 8u                             /* RSIs per long jump of the walk */

/* one doubling step of the RSI lengths: from p over two (four, eight) RSIs, all of which start inside the
 * part of the window that has lengths (0: not available) */
AEC_HD uint32_t sk_hdouble(const uint32_t *src, uint32_t nh_eff, uint32_t p)
{
    const uint32_t h = src[p];
    if (h == 0u) return 0u;
    const uint64_t q = (uint64_t)p + h;
    if (q >= nh_eff) return 0u;
    const uint32_t h2 = src[q];
    return h2 ? h + h2 : 0u;
}

/* Sparse candidates.  An RSI can only start where a CDS chain ends, and chains that start anywhere merge
 * quickly: after 2^j CDSs the chains of ALL positions of a window end on about 2 / 2^j of its positions.
 * The doubling passes therefore mark, in H itself, the positions where the top-level chains end (and where a
 * run-of-zero-segment code ends, after which no chain was followed); RSI lengths are worked out for the marked
 * positions only, the rest of H stays 0 = "not available".  A true RSI start that is not marked (fewer than
 * 2^top plain CDSs before it: the first RSIs of a window, RSIs made of zero runs) is worked out by the walk
 * itself with the same descent; when that happens often the walk switches the stream to dense tables. */
#define SK_CAND 0xFFFFFFFFu                    /* H[p] between the marking and the RSI-length pass: an RSI may start at p */
#define SK_SPARSE_MIN_LEVELS 4u                /* sparse candidates need a top level of 8 CDSs or more (<= 19 % marked) */

AEC_HD uint32_t sk_mark_pos(const AecCfg &c, uint32_t t) { return c.pad ? ((t + 7u) & ~7u) : t; }

/* The entry of an RSI's first CDS (what R[p] holds) by parsing the stream at absolute bit `pos` the way the
 * one-thread scan does: for starts nobody prepared an R entry for (sparse candidates compute R at marked
 * positions only).  0 when the stream does not hold a whole CDS there; a run-of-zero-segment code comes out
 * resolved for block 0 (sk_rsi_len and sk_group_entry resolve the 0 of the tables to the same count). */
AEC_HD uint32_t sk_first_entry_serial(const AecCfg &c, BitRd &br, uint64_t pos)
{
    RsiDec st; st.pos = pos; st.zero_left = 0; st.status = DEC_OK;
    if (!aec_skim_block(c, br, st, 0u)) return 0u;
    const uint64_t len = st.pos - pos;
    const uint32_t blk = 1u + st.zero_left;
    return (len <= 0xFFFFFull && blk <= 0xFFFu) ? (((uint32_t)len << 12) | blk) : 0u;
}

/* the length of SK_SKIP RSIs in a row from listed position p (H[p] != 0), hop by hop: 0 unless every one of them
 * starts inside the part of the window that has lengths and has one (the same value three sk_hdouble steps give) */
AEC_HD uint32_t sk_hchase(const uint32_t *H, uint32_t nh_eff, uint32_t p)
{
    uint64_t q = p;
    for (uint32_t t = 0; t < SK_SKIP; t++) {
        if (q >= nh_eff) return 0u;
        const uint32_t h = H[q];
        if (h == 0u) return 0u;
        q += h;
    }
    return (q - p) <= 0xFFFFFFFFull ? (uint32_t)(q - p) : 0u;
}

/* One RSI of the walk: the serial part of the discovery.  Returns false when the walk ends.
 * state: pos, found, flags (1 ended, 2 data error), fast; slow = RSI lengths the walk had to work out itself. */
struct SkWalk { uint64_t pos, found, flags, fast, slow; };

/* load_h8 is trusted only where load_h is not 0 (with sparse candidates the long-jump buffers hold values at
 * listed positions only); descend(rel) = the RSI length from the chain tables for a start the tables did
 * not prepare (sparse candidates), 0 when there is none. */
template <class LoadH, class LoadH8, class Descend>
AEC_HD bool sk_walk_step(const AecCfg &c, BitRd &br, uint64_t nbits, uint64_t wb, uint32_t nh_eff, uint32_t last,
                         uint64_t *offsets, uint64_t max_rsi, SkWalk &s, LoadH load_h, uint64_t *grp, bool have_h8, LoadH8 load_h8,
                         bool sparse, Descend descend)
{
    if (s.found >= max_rsi) { s.flags = 1; return false; }
    const uint64_t start = c.pad ? ((s.pos + 7ull) & ~7ull) : s.pos;
    if (start >= nbits) { s.flags = 1; return false; }
    if (start >= wb + nh_eff && !last) return false;    /* the next window takes over */
    const uint64_t rel = start - wb;
    const bool inside = rel < nh_eff;
    uint32_t h = inside ? load_h(rel) : 0u;
    const uint32_t h8 = (inside && have_h8) ? load_h8(rel) : 0u;      /* both loads in flight together */
    if (have_h8 && h && h8 && s.found + SK_SKIP <= max_rsi) {
        /* eight RSIs with one look-up: their offsets are filled in afterwards, in parallel (sk_fill) */
        offsets[s.found] = start;
        for (uint32_t t = 1; t < SK_SKIP; t++) offsets[s.found + t] = SK_OFF_PENDING;
        if (grp) for (uint32_t t = 0; t < SK_SKIP; t++) grp[(s.found + t) * 32ull] = SK_GRP_FAST;
        s.found += SK_SKIP; s.fast += SK_SKIP;
        s.pos = start + h8;
        return true;
    }
    offsets[s.found++] = start;                         /* even a truncated RSI may still deliver leading samples */
    if (inside && h == 0u && sparse) { h = descend(rel); s.slow++; }
    if (grp) grp[(s.found - 1) * 32ull] = h ? SK_GRP_FAST : SK_GRP_MISSING;
    if (h) { s.pos = start + h; s.fast++; return true; }
    /* not in the tables: skim this RSI CDS by CDS (truncated or damaged stream, chain leaving the window) */
    RsiDec st; st.pos = start; st.zero_left = 0; st.status = DEC_OK;
    for (uint32_t b = 0; b < c.rsi; b++)
        if (!aec_skim_block(c, br, st, b)) break;
    s.pos = st.pos;
    if (st.status != DEC_OK) { s.flags = 1ull | (st.status == DEC_ERROR ? 2ull : 0ull); return false; }
    return true;
}

/* after a window: so many starts were not among the candidates that dense tables are cheaper from here on */
AEC_HD bool sk_walk_wants_dense(uint64_t slow, uint64_t found) { return slow > 16ull && slow * 4ull > found; }

/* offsets the walk left pending behind RSI r (the head of a long jump): one H look-up each */
AEC_HD void sk_fill(const uint32_t *H, uint64_t wb, uint64_t *offsets, uint64_t r, uint64_t found)
{
    if (offsets[r] == SK_OFF_PENDING || r + 1 >= found || offsets[r + 1] != SK_OFF_PENDING) return;
    uint64_t q = offsets[r];
    for (uint64_t t = 1; t < SK_SKIP && r + t < found && offsets[r + t] == SK_OFF_PENDING; t++) {
        q += H[q - wb];
        offsets[r + t] = q;
    }
}

#endif /* AEC_SKIM_CORE_CUH */
