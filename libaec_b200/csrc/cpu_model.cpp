/*
 * cpu_model.cpp -- TEST HARNESS, not part of the product library.
 *
 * Runs the host+device building blocks of aec_core.cuh / aec_decode_core.cuh on
 * the CPU, arranged exactly like the CUDA kernels arrange them (tiles of TB
 * block slots, padded RSI slots, 64-block zero-run segments, scan monoids,
 * phase-aligned staging words, head/tail boundary words + fix-up), with the
 * warp shuffles and the look-back replaced by serial loops.  tests/ compares
 * its output with the oracle so that the block logic, the geometry and the
 * boundary handling are verified on machines without a GPU.  The product
 * library (libaec.so) does not contain this file.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "aec_decode_core.cuh"

static uint32_t next_pow2(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

struct Slot {
    bool valid, is_zero, rsi_end;
    uint64_t rsi_idx; uint32_t b, ref, refs, len, zcode, zref;
    BlockInfo bi;
    uint32_t d[AEC_MAX_J];
};

/* optional per-block trace for debugging: 4 words per coded block (opt, klo, khi, len) */
static uint32_t *g_trace = nullptr; static size_t g_trace_cap = 0;
extern "C" void model_set_trace(uint32_t *buf, size_t cap) { g_trace = buf; g_trace_cap = cap; }

extern "C" int model_encode(uint32_t n, uint32_t J, uint32_t rsi, uint32_t flags, int honour_pad,
                            const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                            size_t *out_len, uint64_t *offsets, uint64_t seed_bits, uint32_t seed_k,
                            uint32_t seed_word, uint64_t *end_bits, uint32_t *end_k)
{
    AecCfg c;
    if (aec_cfg_init(&c, n, J, rsi, flags, 1, honour_pad) != 0) return -1;
    if (J == 0 || J > AEC_MAX_J || rsi == 0) return -1;
    const uint64_t nsamples = in_len / c.B;
    const uint64_t nrsi = (nsamples + c.R - 1) / c.R;
    if (nsamples == 0) { *out_len = 0; *end_bits = seed_bits; *end_k = seed_k; return 0; }
    const uint64_t last_s = nsamples - (nrsi - 1) * (uint64_t)c.R;
    const uint32_t last_nblk = (uint32_t)((last_s + J - 1) / J);
    const uint32_t TB = (J == 8 || J == 16 || J == 32) ? 256u : 128u;
    const uint32_t NW = TB / 32;
    const uint32_t RP = rsi <= TB ? next_pow2(rsi) : ((rsi + TB - 1) / TB) * TB;
    const uint64_t ntiles = (nrsi * (uint64_t)RP + TB - 1) / TB;

    std::vector<uint32_t> head_c(ntiles, 0), tail_c(ntiles, 0);
    std::vector<uint64_t> tile_end(ntiles, 0);
    std::vector<Slot> S(TB);
    std::vector<uint32_t> staging;
    uint32_t *ow = (uint32_t *)out;
    const uint64_t cap_words = out_cap / 4;

    uint64_t base = seed_bits;
    uint32_t kin = seed_k;
    for (uint64_t tile = 0; tile < ntiles; tile++) {
        std::vector<uint32_t> zb(NW + 1, 0);
        for (uint32_t tid = 0; tid < TB; tid++) {
            Slot &s = S[tid];
            if (RP >= TB) { uint32_t tpr = RP / TB; s.rsi_idx = tile / tpr; s.b = (uint32_t)(tile % tpr) * TB + tid; }
            else { s.rsi_idx = tile * (TB / RP) + tid / RP; s.b = tid % RP; }
            uint32_t nblk = 0;
            if (s.rsi_idx + 1 < nrsi) nblk = c.rsi; else if (s.rsi_idx + 1 == nrsi) nblk = last_nblk;
            s.valid = s.b < nblk;
            s.ref = (s.valid && c.pp && s.b == 0) ? 1u : 0u;
            s.refs = 0; s.len = 0; s.zcode = 0; s.zref = 0;
            s.bi.opt = OPT_NONE; s.bi.klo = 0; s.bi.khi = c.kmax; s.bi.len = 0;
            if (s.valid) {
                uint64_t first = s.rsi_idx * (uint64_t)c.R + (uint64_t)s.b * J;
                for (uint32_t i = 0; i < J; i++) {
                    uint64_t idx = first + i; if (idx >= nsamples) idx = nsamples - 1;
                    s.d[i] = aec_load_sample(in + idx * c.B, c.B, c.msb);
                }
                if (c.pp) {
                    uint32_t prev;
                    if (s.b == 0) { s.refs = s.d[0]; prev = s.d[0] ^ c.sflip; }
                    else { uint64_t pi = first - 1; if (pi >= nsamples) pi = nsamples - 1;
                           prev = aec_load_sample(in + pi * c.B, c.B, c.msb) ^ c.sflip; }
                    for (uint32_t i = 0; i < J; i++) { uint32_t u = s.d[i] ^ c.sflip; s.d[i] = aec_map_delta(prev, u, c.mask); prev = u; }
                    if (s.b == 0) s.d[0] = 0;
                }
                s.bi = aec_analyze_block<0>(c, s.d, s.ref);
            }
            s.is_zero = s.valid && s.bi.opt == OPT_ZERO;
            s.rsi_end = s.valid && (s.b + 1 == nblk);
            if (s.is_zero) zb[tid >> 5] |= 1u << (tid & 31);
        }
        for (uint32_t tid = 0; tid < TB; tid++) {
            Slot &s = S[tid];
            s.len = s.bi.len;
            if (s.is_zero) {
                uint32_t warp = tid >> 5;
                uint32_t nblk = (s.rsi_idx + 1 < nrsi) ? c.rsi : last_nblk;
                uint64_t m64 = (uint64_t)zb[warp & ~1u] | ((uint64_t)((warp | 1u) < NW ? zb[warp | 1u] : 0u) << 32);
                uint32_t q = tid & 63u, g0 = q - (s.b & 63u), seg = s.b >> 6;
                uint32_t V = nblk - seg * 64u; if (V > 64u) V = 64u;
                uint64_t segmask = m64 >> g0;
                if (V < 64u) segmask &= ((1ull << V) - 1ull);
                uint32_t rl = 0; s.len = aec_zero_run(c, segmask, V, s.b, &s.zcode, &s.zref, &rl);
                if (s.zref) s.refs = aec_load_sample(in + s.rsi_idx * (uint64_t)c.R * c.B, c.B, c.msb);
            }
        }
        if (g_trace) for (uint32_t tid = 0; tid < TB; tid++) {
            Slot &s = S[tid];
            if (!s.valid) continue;
            size_t bidx = (size_t)(s.rsi_idx * c.rsi + s.b);
            if (bidx * 4 + 3 < g_trace_cap) { g_trace[bidx*4] = s.bi.opt; g_trace[bidx*4+1] = s.bi.klo; g_trace[bidx*4+2] = s.bi.khi; g_trace[bidx*4+3] = s.len; }
        }
        /* serial scans with the same monoids */
        std::vector<PosFn> pexc(TB); std::vector<uint32_t> kbefore(TB);
        PosFn acc; acc.has_end = 0; acc.a = 0; acc.rest = 0;
        uint32_t kacc = aec_kpair(0, c.kmax);
        for (uint32_t tid = 0; tid < TB; tid++) {
            pexc[tid] = acc; kbefore[tid] = kacc;
            PosFn e; e.has_end = (c.pad && S[tid].rsi_end) ? 1u : 0u; e.a = S[tid].len; e.rest = 0;
            acc = aec_pcompose(acc, e);
            kacc = aec_kcompose(kacc, aec_kpair(S[tid].bi.klo, S[tid].bi.khi));
        }
        const uint64_t end = aec_papply(acc, base);
        const uint32_t kout = aec_kapply(kin, kacc);
        tile_end[tile] = end;
        const uint64_t w0 = base >> 5, we = end >> 5;
        staging.assign((size_t)(we - w0) + 4, 0);
        for (uint32_t tid = 0; tid < TB; tid++) {
            Slot &s = S[tid];
            uint64_t myoff = aec_papply(pexc[tid], base);
            if (s.valid && s.b == 0 && offsets) offsets[s.rsi_idx] = myoff;
            if (s.valid && s.len) {
                BitPack bp; bp.init(staging.data(), (uint32_t)(myoff - (w0 << 5)), 0u);
                if (s.is_zero) aec_pack_zero(c, bp, s.zcode, s.zref, s.refs);
                else {
                    uint32_t kprev = aec_kapply(kin, kbefore[tid]);
                    uint32_t k = aec_clampu(kprev, s.bi.klo, s.bi.khi);
                    aec_pack_block<0>(c, bp, s.d, s.bi.opt, k, s.ref, s.refs);
                }
                bp.finish(0u);
            }
        }
        const uint32_t nw = (uint32_t)(we - w0) + ((end & 31u) ? 1u : 0u);
        const bool head_partial = (base & 31u) != 0;
        const bool tail_partial = (end & 31u) != 0 && (we > w0 || !head_partial);
        for (uint32_t i = 0; i < nw; i++) {
            uint64_t wi = w0 + i;
            if (i == 0 && head_partial) continue;
            if (wi == we) continue;
            if (wi < cap_words) ow[wi] = aec_bswap32(staging[i]);
        }
        head_c[tile] = (head_partial && end > base) ? staging[0] : 0u;
        tail_c[tile] = tail_partial ? staging[(size_t)(we - w0)] : 0u;
        base = end; kin = kout;
    }
    /* fix-up */
    const uint64_t total = tile_end[ntiles - 1];
    for (int64_t i = -1; i < (int64_t)ntiles; i++) {
        uint64_t bi = (i <= 0) ? seed_bits : tile_end[i - 1];
        uint64_t ei = (i < 0) ? seed_bits : tile_end[i];
        uint32_t v;
        if ((ei & 31u) == 0) continue;
        if (i < 0) v = seed_word;
        else {
            bool head_partial = (bi & 31u) != 0;
            if (!((ei >> 5) > (bi >> 5) || !head_partial)) continue;
            v = tail_c[i];
        }
        uint64_t word = ei >> 5;
        for (int64_t j = i + 1; j < (int64_t)ntiles; j++) { v |= head_c[j]; if ((tile_end[j] >> 5) > word) break; }
        uint64_t limit = (total + 7) >> 3; if (limit > out_cap) limit = out_cap;
        for (int bq = 0; bq < 4; bq++) { uint64_t bidx = word * 4 + bq; if (bidx < limit) out[bidx] = (uint8_t)(v >> (24 - 8 * bq)); }
    }
    *out_len = (size_t)((total + 7) / 8);
    *end_bits = total; *end_k = kin;
    return 0;
}

/* Decode with the device building blocks: discover RSI offsets by skimming,
 * then decode every RSI independently from its offset. */
extern "C" int model_decode(uint32_t n, uint32_t J, uint32_t rsi, uint32_t flags,
                            const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                            size_t *out_len, const uint64_t *given_offsets, size_t n_given)
{
    AecCfg c;
    if (aec_cfg_init(&c, n, J, rsi, flags, 0, 0) != 0) return -1;
    if (J == 0 || J > AEC_MAX_J || rsi == 0) return -1;
    c.pad = (flags & AECF_PAD_RSI) ? 1u : 0u;
    const uint64_t out_samples = out_cap / c.B;
    const uint64_t need_rsi = (out_samples + c.R - 1) / c.R;
    size_t in_pad = (in_len + 3) & ~(size_t)3;
    std::vector<uint32_t> words(in_pad / 4 + 1, 0);
    memcpy(words.data(), in, in_len);
    BitRd br; br.init(words.data(), in_pad / 4, (uint64_t)in_len * 8);
    std::vector<uint64_t> offs;
    if (given_offsets) { for (size_t i = 0; i < n_given && i < need_rsi; i++) offs.push_back(given_offsets[i]); }
    else {
        RsiDec st; st.pos = 0; st.zero_left = 0; st.status = DEC_OK;
        for (uint64_t r = 0; r < need_rsi; r++) {
            if (c.pad) st.pos = (st.pos + 7ull) & ~7ull;
            uint64_t start = st.pos; st.zero_left = 0;
            if (start >= br.nbits) break;
            offs.push_back(start);          /* even a truncated RSI may still deliver leading samples */
            for (uint32_t b = 0; b < c.rsi; b++) if (!aec_skim_block(c, br, st, b)) break;
            if (st.status != DEC_OK) break;
        }
    }
    uint64_t total = out_samples < offs.size() * (uint64_t)c.R ? out_samples : offs.size() * (uint64_t)c.R;
    int err = 0;
    for (size_t r = 0; r < offs.size(); r++) {
        uint64_t startS = r * (uint64_t)c.R;
        uint64_t limit = out_samples > startS ? out_samples - startS : 0; if (limit > c.R) limit = c.R;
        BitRd b2; b2.init(words.data(), in_pad / 4, (uint64_t)in_len * 8);
        RsiDec st; st.pos = offs[r]; st.zero_left = 0; st.status = DEC_OK;
        uint32_t row[AEC_MAX_J + 1], u_prev = 0, delivered = 0;
        bool active = limit > 0;
        for (uint32_t b = 0; b < c.rsi && active; b++) {
            uint32_t cnt = aec_decode_block<0>(c, b2, st, b, row);
            uint64_t room = limit - delivered; if (cnt > room) cnt = (uint32_t)room;
            aec_unmap_row(c, row, cnt, (c.pp && b == 0) ? 1u : 0u, &u_prev);
            for (uint32_t i = 0; i < cnt; i++) aec_store_sample(out + (startS + (uint64_t)b * J + i) * c.B, row[i], c.B, c.msb);
            delivered += cnt;
            if (cnt < J || delivered >= limit) active = false;
        }
        if (delivered < limit && startS + delivered < total) total = startS + delivered;
        if (st.status == DEC_ERROR) err = 1;
    }
    *out_len = (size_t)(total * c.B);
    return err ? -3 : 0;
}

/* ------------------------------------------------------------------------ */
/* RSI boundary discovery (aec_skim.cu) on the CPU: the same window loop, tile staging, level tables,
 * RSI lengths and walk, with the kernels' grids replaced by loops.                                   */
/* ------------------------------------------------------------------------ */
#include "aec_skim_core.cuh"

static int g_skip8 = 0;
extern "C" void model_set_skip8(int on) { g_skip8 = on; }   /* the walk's eight-RSI jumps (aec_skim_hdouble_kernel, aec_skim_fill_kernel) */
static int g_sparse = 1;
extern "C" void model_set_sparse(int on) { g_sparse = on; } /* RSI lengths at marked chain ends only (aec_skim_rsi_sparse_kernel) */
static uint64_t g_slow = 0, g_dense_windows = 0, g_marked = 0, g_positions = 0;
extern "C" void model_sparse_stats(uint64_t *slow, uint64_t *dense_windows, uint64_t *marked, uint64_t *positions)
{
    *slow = g_slow; *dense_windows = g_dense_windows; *marked = g_marked; *positions = g_positions;
}

/* the group index the way aec_build_group_index_kernel builds it: one skim of the RSI from its start offset */
static void group_index_by_skim(const AecCfg &c, BitRd &br, uint64_t start, uint64_t *g)
{
    const uint32_t G = (c.rsi + 31u) / 32u;
    RsiDec st; st.pos = start; st.zero_left = 0; st.status = DEC_OK;
    for (uint32_t i = 0; i < 32; i++) g[i] = 0;
    for (uint32_t b = 0; b < c.rsi; b++) {
        if (b % G == 0) g[b / G] = ((uint64_t)st.zero_left << 56) | (st.pos & 0x00FFFFFFFFFFFFFFull);
        if (!aec_skim_block(c, br, st, b)) {
            for (uint32_t q = b / G + 1; q * G < c.rsi; q++) g[q] = 0x00FFFFFFFFFFFFFFull;
            break;
        }
    }
}

extern "C" int model_scan_offsets_grp(uint32_t n, uint32_t J, uint32_t rsi, uint32_t flags,
                                      const uint8_t *in, size_t in_bytes, uint64_t start_bit,
                                      uint64_t *offsets, uint64_t max_rsi, uint64_t window_bits, int serial,
                                      uint64_t *found, uint64_t *out_flags, uint64_t *fast, uint64_t *end_pos,
                                      uint64_t *grp, uint64_t *grp_ref);

extern "C" int model_scan_offsets(uint32_t n, uint32_t J, uint32_t rsi, uint32_t flags,
                                  const uint8_t *in, size_t in_bytes, uint64_t start_bit,
                                  uint64_t *offsets, uint64_t max_rsi, uint64_t window_bits, int serial,
                                  uint64_t *found, uint64_t *out_flags, uint64_t *fast, uint64_t *end_pos)
{
    return model_scan_offsets_grp(n, J, rsi, flags, in, in_bytes, start_bit, offsets, max_rsi, window_bits, serial,
                                  found, out_flags, fast, end_pos, nullptr, nullptr);
}

/* grp (optional, max_rsi * 32): the group index from the tables (skim for RSIs the tables could not give);
 * grp_ref (optional): the same index built by skimming every RSI, for comparison */
extern "C" int model_scan_offsets_grp(uint32_t n, uint32_t J, uint32_t rsi, uint32_t flags,
                                      const uint8_t *in, size_t in_bytes, uint64_t start_bit,
                                      uint64_t *offsets, uint64_t max_rsi, uint64_t window_bits, int serial,
                                      uint64_t *found, uint64_t *out_flags, uint64_t *fast, uint64_t *end_pos,
                                      uint64_t *grp, uint64_t *grp_ref)
{
    AecCfg c;
    if (aec_cfg_init(&c, n, J, rsi, flags, 0, 0) != 0) return -1;
    if (J == 0 || J > AEC_MAX_J || rsi == 0) return -1;
    c.pad = (flags & AECF_PAD_RSI) ? 1u : 0u;
    const uint64_t nbits = (uint64_t)in_bytes * 8ull;
    const uint64_t total_words = (nbits + 31ull) >> 5;
    std::vector<uint32_t> words(total_words + 4, 0);            /* raw byte order like the device buffer */
    memcpy(words.data(), in, in_bytes);
    BitRd br;
    br.init(words.data(), total_words, nbits);
    SkWalk s; s.pos = start_bit; s.found = 0; s.flags = 0; s.fast = 0; s.slow = 0;
    g_slow = g_dense_windows = g_marked = g_positions = 0;
    *found = 0; *out_flags = 0; *fast = 0; *end_pos = start_bit;
    if (max_rsi == 0) return 0;
    const uint64_t base = start_bit & ~127ull;
    if (serial || base >= nbits) {
        /* the one-thread scan (aec_scan_offsets_kernel) */
        RsiDec st; st.pos = start_bit; st.zero_left = 0; st.status = DEC_OK;
        uint64_t f = 0;
        for (uint64_t r = 0; r < max_rsi; r++) {
            if (c.pad) st.pos = (st.pos + 7ull) & ~7ull;
            st.zero_left = 0;
            if (st.pos >= br.nbits) break;
            offsets[f++] = st.pos;
            for (uint32_t b = 0; b < c.rsi; b++)
                if (!aec_skim_block(c, br, st, b)) break;
            if (st.status != DEC_OK) break;
        }
        *found = f; *out_flags = (st.status == DEC_ERROR) ? 2 : 0; *end_pos = st.pos;
        return 0;
    }
    const uint32_t LV = sk_levels(c);
    const uint64_t margin = sk_margin_bits(c);
    uint64_t nh = (window_bits + 127ull) & ~127ull;
    if (nh < 1024) nh = 1024;
    const uint64_t span = ((nbits - base) + 31ull) & ~31ull;
    const uint64_t nwin = (span + nh - 1) / nh;
    const uint32_t TILE = 8192, la = sk_lookahead_words(c), nwords = TILE / 32u + la;
    std::vector<uint32_t> T, H, Rv, w(nwords + 1), pre(nwords + 2);
    const uint32_t G = (c.rsi + 31u) / 32u;
    const bool sparse_ok = g_sparse && LV >= SK_SPARSE_MIN_LEVELS;
    bool dense = false;                                                 /* state[5]: the walk found too many unmarked starts */
    for (uint64_t wi = 0; wi < nwin && !(s.flags & 1ull); wi++) {
        const uint64_t wb = base + wi * nh;
        const uint64_t rem = ((nbits - wb) + 31ull) & ~31ull;
        const uint32_t np = (uint32_t)(nh + margin < rem ? nh + margin : rem);
        const uint32_t last = (wi + 1 == nwin) ? 1u : 0u;
        const uint32_t nh_eff = last ? np : (uint32_t)nh;
        const bool sparse = sparse_ok && !dense;
        T.assign((size_t)LV * np, 0u); H.assign(np, 0u); Rv.assign(np, sparse ? 0xDEADBEEFu : 0u);   /* sparse: R exists at candidates only */
        for (uint32_t tile0 = 0; tile0 < np; tile0 += TILE) {           /* level-0 kernel, one CTA */
            const uint64_t word0 = (wb + tile0) >> 5;
            for (uint32_t i = 0; i <= nwords; i++) {
                const uint64_t x = word0 + i;
                w[i] = x < total_words ? aec_bswap32(words[x]) : 0u;
            }
            pre[0] = 0;
            for (uint32_t i = 0; i <= nwords; i++) pre[i + 1] = pre[i] + (i < nwords ? sk_popc(w[i]) : 0u);
            const uint64_t tile_abs = wb + tile0;
            const uint32_t limit = nbits > tile_abs ? (uint32_t)((nbits - tile_abs) < 0x7FFFFFFFull ? (nbits - tile_abs) : 0x7FFFFFFFull) : 0u;
            for (uint32_t q = 0; q < TILE && tile0 + q < np; q++) {
                uint32_t t0 = 0, r0 = 0;
                if (q < limit) {
                    t0 = sk_entry(c, w.data(), pre.data(), nwords, q, limit, 0u);
                    r0 = (c.pp && !sparse) ? sk_entry(c, w.data(), pre.data(), nwords, q, limit, 1u) : t0;
                }
                T[tile0 + q] = t0;
                if (!sparse) Rv[tile0 + q] = r0;
            }
        }
        for (uint32_t j = 0; j + 1 < LV; j++)
            for (uint32_t p = 0; p < np; p++) T[(size_t)(j + 1) * np + p] = sk_double(T.data() + (size_t)j * np, np, p);
        const uint32_t stp = c.pad ? 8u : 1u;
        std::vector<uint32_t> list;
        if (sparse) {
            /* marks: where the top-level chains end (last doubling pass) and where run-of-zero-segment codes end (first pass) */
            const uint32_t *Tt = T.data() + (size_t)(LV - 1u) * np;
            for (uint32_t p = 0; p < np; p++) {
                if (sk_jump(Tt[p])) { const uint32_t t = sk_mark_pos(c, p + sk_len(Tt[p])); if (t < nh_eff) H[t] = SK_CAND; }
                if (sk_ros(T[p])) { const uint32_t t = sk_mark_pos(c, p + sk_len(T[p])); if (t < nh_eff) H[t] = SK_CAND; }
            }
            /* aec_skim_rsi_sparse_kernel: a CTA stages the words of its chunk of positions and works out R for its candidates */
            const uint32_t CH = 16384, cwords = CH / 32u + la;
            std::vector<uint32_t> cw(cwords + 1), cpre(cwords + 2);
            uint32_t staged = 0xFFFFFFFFu;
            for (uint32_t p = 0; p < nh_eff; p += stp) {
                if (H[p] != SK_CAND) { H[p] = 0u; continue; }
                const uint32_t ch0 = p / CH * CH;
                if (staged != ch0) {
                    const uint64_t word0 = (wb + ch0) >> 5;
                    for (uint32_t i = 0; i <= cwords; i++) { const uint64_t x = word0 + i; cw[i] = x < total_words ? aec_bswap32(words[x]) : 0u; }
                    cpre[0] = 0;
                    for (uint32_t i = 0; i <= cwords; i++) cpre[i + 1] = cpre[i] + (i < cwords ? sk_popc(cw[i]) : 0u);
                    staged = ch0;
                }
                const uint64_t ch_abs = wb + ch0;
                const uint32_t climit = nbits > ch_abs ? (uint32_t)((nbits - ch_abs) < 0x7FFFFFFFull ? (nbits - ch_abs) : 0x7FFFFFFFull) : 0u;
                const uint32_t q = p - ch0;
                Rv[p] = q < climit ? sk_entry(c, cw.data(), cpre.data(), cwords, q, climit, c.pp ? 1u : 0u) : 0u;
                H[p] = sk_rsi_len(c, T.data(), LV, np, p, Rv[p]);
                g_marked++;
                if (H[p]) list.push_back(p);
            }
            g_positions += nh_eff / stp;
        } else {
            for (uint32_t p = 0; p < nh_eff; p += stp) H[p] = sk_rsi_len(c, T.data(), LV, np, p, Rv[p]);
            g_dense_windows++;
        }
        const uint32_t *Hp = H.data();
        std::vector<uint32_t> Ha, Hb;
        if (g_skip8 && sparse) {
            /* the long-jump buffers hold values at listed positions only; everything else is stale */
            Hb.assign(np, 0xDEADBEEFu);
            for (uint32_t p : list) Hb[p] = sk_hchase(Hp, nh_eff, p);       /* aec_skim_hchase_list_kernel */
        } else if (g_skip8) {
            Ha.assign(np, 0u); Hb.assign(np, 0u);
            for (uint32_t p = 0; p < nh_eff; p += stp) Hb[p] = sk_hdouble(H.data(), nh_eff, p);
            for (uint32_t p = 0; p < nh_eff; p += stp) Ha[p] = sk_hdouble(Hb.data(), nh_eff, p);
            for (uint32_t p = 0; p < nh_eff; p += stp) Hb[p] = sk_hdouble(Ha.data(), nh_eff, p);
        }
        const uint32_t *H8p = g_skip8 ? Hb.data() : nullptr;
        const uint64_t f0 = s.found, slow0 = s.slow;
        const uint32_t *Tp = T.data(); uint32_t *Rp = Rv.data();
        while (sk_walk_step(c, br, nbits, wb, nh_eff, last, offsets, max_rsi, s,
                            [Hp](uint64_t rel) { return Hp[rel]; }, grp, H8p != nullptr,
                            [H8p](uint64_t rel) { return H8p[rel]; }, sparse,
                            [&c, &br, Tp, Rp, LV, np, wb, sparse](uint64_t rel) {
                                /* a start nobody marked has no R entry either: parse its first CDS, keep it for the group index */
                                if (sparse) Rp[rel] = sk_first_entry_serial(c, br, wb + rel);
                                return sk_rsi_len(c, Tp, LV, np, (uint32_t)rel, Rp[rel]); })) { }
        if (sparse && sk_walk_wants_dense(s.slow - slow0, s.found - f0)) dense = true;
        g_slow = s.slow;
        if (g_skip8)
            for (uint64_t r = f0; r < s.found; r++) sk_fill(Hp, wb, offsets, r, s.found);
        if (grp) {                                                      /* aec_skim_group_index_kernel */
            for (uint64_t r = f0; r < s.found; r++) {
                uint64_t *g = grp + r * 32;
                if (g[0] != SK_GRP_FAST) continue;
                const uint32_t p = (uint32_t)(offsets[r] - wb);
                for (uint32_t l = 0; l < 32; l++)
                    g[l] = l * G < c.rsi ? sk_group_entry(c, T.data(), Rv.data(), LV, np, wb, p, l * G) : 0ull;
            }
        }
    }
    for (uint64_t r = 0; r < s.found; r++) {
        if (grp && grp[r * 32] == SK_GRP_MISSING) group_index_by_skim(c, br, offsets[r], grp + r * 32);   /* builder, only_missing */
        if (grp_ref) group_index_by_skim(c, br, offsets[r], grp_ref + r * 32);
    }
    *found = s.found; *out_flags = s.flags; *fast = s.fast; *end_pos = s.pos;
    return 0;
}
