/*
 * aec_oracle.c -- plain-C, whole-buffer CPU oracle for CCSDS 121.0-B-2 as
 * implemented by reference libaec 0.3.4.
 *
 * TEST INFRASTRUCTURE ONLY (see aec_oracle.h).  The reference is a pair of
 * resumable state machines; this file restates what they compute as straight
 * loops over RSIs and blocks, writing one bit field at a time.  It is slow on
 * purpose: every step is meant to be obviously equal to the cited reference
 * lines.  Citations are relative to /root/reference/.
 *
 * Parity: PINNED against the compiled reference and its golden vector, see
 * tests/test_oracle.py.
 */
#include "aec_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* Derived coding parameters                                                  */
/* ------------------------------------------------------------------------- */

typedef struct {
    uint32_t n;        /* bits per sample */
    uint32_t J;        /* block size */
    uint32_t rsi;      /* blocks per reference sample interval */
    uint32_t flags;
    uint32_t B;        /* storage bytes per sample */
    uint32_t idl;      /* option id length in bits */
    uint32_t kmax;
    int      pp, sgn, msb;
    int64_t  xmin, xmax;
} coder;

/* Common to src/encode.c:777-872 and src/decode.c:699-763. `enc` selects the
 * encoder-only validation (block size, rsi): the decoder checks neither
 * (decode.c:694-795 never looks at block_size or rsi). */
static int setup(coder *c, const orc_params *p, int enc)
{
    memset(c, 0, sizeof *c);
    c->n = p->bits_per_sample;
    c->J = p->block_size;
    c->rsi = p->rsi;
    c->flags = p->flags;
    if (c->n == 0 || c->n > 32)
        return ORC_CONF_ERROR;                       /* encode.c:777, decode.c:699 */
    if (enc) {
        if (c->flags & ORC_NOT_ENFORCE) {
            if (c->J & 1)
                return ORC_CONF_ERROR;               /* encode.c:780-783 */
        } else if (c->J != 8 && c->J != 16 && c->J != 32 && c->J != 64) {
            return ORC_CONF_ERROR;                   /* encode.c:785-791 */
        }
        if (c->rsi > 4096)
            return ORC_CONF_ERROR;                   /* encode.c:793 */
    }
    if (c->n > 16) {                                 /* encode.c:804-826 */
        c->idl = 5;
        c->B = (c->n <= 24 && (c->flags & ORC_DATA_3BYTE)) ? 3 : 4;
    } else if (c->n > 8) {                           /* encode.c:828-839 */
        c->idl = 4;
        c->B = 2;
    } else {                                         /* encode.c:841-859 */
        if (c->flags & ORC_RESTRICTED) {
            if (c->n <= 2) c->idl = 1;
            else if (c->n <= 4) c->idl = 2;
            else return ORC_CONF_ERROR;
        } else {
            c->idl = 3;
        }
        c->B = 1;
    }
    c->sgn = (c->flags & ORC_DATA_SIGNED) != 0;
    c->pp = (c->flags & ORC_DATA_PREPROCESS) != 0;
    c->msb = (c->flags & ORC_DATA_MSB) != 0;
    if (c->sgn) {                                    /* encode.c:862-865 */
        c->xmax = ((int64_t)1 << (c->n - 1)) - 1;
        c->xmin = -c->xmax - 1;
    } else {                                         /* encode.c:867-869 */
        c->xmin = 0;
        c->xmax = ((int64_t)1 << c->n) - 1;
    }
    c->kmax = (1u << c->idl) - 3;                    /* encode.c:872 */
    return ORC_OK;
}

/* ------------------------------------------------------------------------- */
/* Bit writer: MSB first, grows on demand (encode.c:61-104 semantics)         */
/* ------------------------------------------------------------------------- */

typedef struct {
    uint8_t *buf;
    size_t cap;      /* bytes allocated */
    uint64_t nbits;  /* bits written */
    int oom;
} bitw;

static void bw_reserve(bitw *w, uint64_t morebits)
{
    size_t need = (size_t)((w->nbits + morebits + 7) / 8) + 8;
    if (need <= w->cap || w->oom)
        return;
    size_t ncap = w->cap ? w->cap : 4096;
    while (ncap < need)
        ncap *= 2;
    uint8_t *nb = (uint8_t *)realloc(w->buf, ncap);
    if (!nb) {
        w->oom = 1;
        return;
    }
    memset(nb + w->cap, 0, ncap - w->cap);
    w->buf = nb;
    w->cap = ncap;
}

/* emit(v, len): len-bit big-endian field (encode.c:61-83) */
static void bw_put(bitw *w, uint64_t v, uint32_t len)
{
    bw_reserve(w, len);
    if (w->oom)
        return;
    for (uint32_t i = 0; i < len; i++) {
        uint32_t bit = (uint32_t)((v >> (len - 1 - i)) & 1u);
        if (bit)
            w->buf[w->nbits >> 3] |= (uint8_t)(0x80u >> (w->nbits & 7));
        w->nbits++;
    }
}

/* emitfs(m): m zero bits then a one (encode.c:85-104) */
static void bw_fs(bitw *w, uint64_t m)
{
    bw_reserve(w, m + 1);
    if (w->oom)
        return;
    w->nbits += m;
    w->buf[w->nbits >> 3] |= (uint8_t)(0x80u >> (w->nbits & 7));
    w->nbits++;
}

static void bw_align8(bitw *w)
{
    bw_reserve(w, 8);
    w->nbits = (w->nbits + 7) & ~(uint64_t)7;
}

/* ------------------------------------------------------------------------- */
/* Sample input (encode_accessors.c:61-143)                                   */
/* ------------------------------------------------------------------------- */

static uint32_t load_sample(const coder *c, const uint8_t *p)
{
    uint32_t v = 0;
    if (c->msb)
        for (uint32_t i = 0; i < c->B; i++) v = (v << 8) | p[i];
    else
        for (uint32_t i = 0; i < c->B; i++) v |= (uint32_t)p[i] << (8 * i);
    return v;
}

/* ------------------------------------------------------------------------- */
/* Encoder                                                                    */
/* ------------------------------------------------------------------------- */

/* Unit-delay predictor + mapper over one RSI (encode.c:235-311).
 * raw[]: R samples as loaded; d[]: mapped output; returns nothing.
 * Arithmetic is done in int64 where the reference uses wrapping 32-bit ops;
 * for in-contract inputs (unused high bits zero, README.md:141-144) both
 * agree. */
static void preprocess(const coder *c, const uint32_t *raw, uint32_t *d, uint32_t R)
{
    int64_t prev = raw[0];
    if (c->sgn) {
        int64_t m = (int64_t)1 << (c->n - 1);
        prev = ((int64_t)(raw[0] ^ (uint32_t)m)) - m;      /* encode.c:290-292 */
        if (c->n == 32) prev = (int32_t)(raw[0]);
    }
    d[0] = 0;                                               /* encode.c:253, :291 */
    for (uint32_t i = 1; i < R; i++) {
        int64_t cur = raw[i];
        if (c->sgn) {
            int64_t m = (int64_t)1 << (c->n - 1);
            cur = (c->n == 32) ? (int64_t)(int32_t)raw[i]
                               : ((int64_t)(raw[i] ^ (uint32_t)m)) - m;
        }
        int64_t D;
        if (cur >= prev) {                                  /* encode.c:256-261, :303-308 */
            D = cur - prev;
            d[i] = (D <= prev - c->xmin) ? (uint32_t)(2 * D) : (uint32_t)(cur - c->xmin);
        } else {                                            /* encode.c:262-268, :296-302 */
            D = prev - cur;
            d[i] = (D <= c->xmax - prev) ? (uint32_t)(2 * D - 1) : (uint32_t)(c->xmax - cur);
        }
        prev = cur;
    }
}

/* Split option: argmin plateau [lo,hi] of len(k) over k in 0..kmax and the
 * minimum length (encode.c:313-410; SURVEY App. B1 shows the reference's
 * hill-climb from the previous k returns clamp(k_prev, lo, hi)).
 * Brute force over every k on purpose. */
static uint64_t split_lengths(const coder *c, const uint32_t *blk, uint32_t thisbs,
                              uint32_t *lo, uint32_t *hi)
{
    uint64_t best = UINT64_MAX;
    *lo = *hi = 0;
    for (uint32_t k = 0; k <= c->kmax; k++) {
        uint64_t fs = 0;
        for (uint32_t i = 0; i < c->J; i++)
            fs += (uint64_t)(blk[i] >> k);                  /* encode.c:313-327 */
        uint64_t len = fs + (uint64_t)thisbs * (k + 1);     /* encode.c:375 */
        if (len < best) {
            best = len;
            *lo = *hi = k;
        } else if (len == best) {
            *hi = k;
        }
    }
    return best;
}

/* Second-extension length with the reference's early exit (encode.c:412-434). */
static uint32_t se_length(const coder *c, const uint32_t *blk, uint32_t uncomp_len)
{
    uint64_t len = 1;
    for (uint32_t i = 0; i < c->J; i += 2) {
        uint64_t s = (uint64_t)blk[i] + (uint64_t)blk[i + 1];
        len += s * (s + 1) / 2 + blk[i + 1] + 1;            /* u64, wraps like the reference */
        if (len > uncomp_len)
            return UINT32_MAX;
    }
    return (uint32_t)len;
}

/* Zero-run CDS (encode.c:565-583). run < 0 means ROS. */
static void put_zero_run(const coder *c, bitw *w, int run, int zref, uint32_t refs)
{
    bw_put(w, 0, c->idl + 1);
    if (zref)
        bw_put(w, refs, c->n);
    if (run < 0)
        bw_fs(w, 4);
    else if (run >= 5)
        bw_fs(w, (uint64_t)run);
    else
        bw_fs(w, (uint64_t)run - 1);
}

int orc_encode(const orc_params *p, int honour_pad_rsi,
               const uint8_t *in, size_t in_len,
               uint8_t *out, size_t out_cap, size_t *out_len,
               size_t *in_consumed,
               uint64_t *rsi_bit_offsets, size_t offsets_cap, size_t *n_offsets,
               orc_block_trace *trace, size_t trace_cap, size_t *n_trace)
{
    coder c;
    int st = setup(&c, p, 1);
    if (out_len) *out_len = 0;
    if (in_consumed) *in_consumed = 0;
    if (n_offsets) *n_offsets = 0;
    if (n_trace) *n_trace = 0;
    if (st != ORC_OK)
        return st;

    const uint32_t R = c.rsi * c.J;
    const size_t nsamples = in_len / c.B;                   /* encode.c:673: partial sample bytes stay unconsumed */
    bitw w = {0};
    size_t noff = 0, ntr = 0;
    uint32_t k_prev = 0;                                    /* encode.c:800 memset; never reset afterwards */

    uint32_t *raw = (uint32_t *)malloc(sizeof(uint32_t) * (R ? R : 1));
    uint32_t *d = (uint32_t *)malloc(sizeof(uint32_t) * (R ? R : 1));
    if (!raw || !d) {
        free(raw); free(d);
        return ORC_MEM_ERROR;
    }

    size_t done = 0;
    while (R && done < nsamples) {
        size_t s = nsamples - done;
        if (s > R) s = R;
        uint32_t nblk = (s == R) ? c.rsi : (uint32_t)((s + c.J - 1) / c.J);   /* encode.c:678-680 */
        for (size_t i = 0; i < s; i++)
            raw[i] = load_sample(&c, in + (done + i) * c.B);
        for (size_t i = s; i < R; i++)
            raw[i] = raw[s - 1];                            /* encode.c:681-684 */

        if (rsi_bit_offsets && noff < offsets_cap)
            rsi_bit_offsets[noff] = w.nbits;
        noff++;

        uint32_t refs = 0;
        const uint32_t *blkdata = raw;
        if (c.pp) {
            refs = raw[0];                                  /* encode.c:252, :289 */
            preprocess(&c, raw, d, R);
            blkdata = d;
        }

        int run = 0, zref = 0;
        for (uint32_t bi = 0; bi < nblk; bi++) {
            const uint32_t *blk = blkdata + (size_t)bi * c.J;
            int ref = (c.pp && bi == 0);
            int cut = (bi == nblk - 1) || ((bi + 1) % 64 == 0);   /* encode.c:649 */
            int allzero = 1;
            for (uint32_t i = 0; i < c.J; i++)
                if (blk[i]) { allzero = 0; break; }         /* encode.c:626-628 */

            orc_block_trace tr = {0, 0, 0, 0, 0};
            uint64_t before = w.nbits;

            if (allzero) {
                run++;
                if (run == 1) zref = ref;                   /* encode.c:645-648 */
                if (cut) {
                    put_zero_run(&c, &w, run > 4 ? -1 : run, zref, refs);   /* encode.c:649-654 */
                    run = 0;
                }
                tr.option = 0;
            } else {
                if (run) {                                  /* encode.c:631-638 */
                    /* bits of the pending run are attributed to its last zero block */
                    uint64_t b0 = w.nbits;
                    put_zero_run(&c, &w, run, zref, refs);
                    if (trace && ntr >= 1 && ntr - 1 < trace_cap)
                        trace[ntr - 1].cds_bits = (uint32_t)(w.nbits - b0);
                    before = w.nbits;
                    run = 0;
                }
                uint32_t thisbs = c.J - (uint32_t)ref;
                uint32_t unc = thisbs * c.n;                /* encode.c:270, :746 */
                uint32_t split = UINT32_MAX, lo = 0, hi = 0, k = 0;
                if (c.idl > 1) {                            /* encode.c:595-598 */
                    split = (uint32_t)split_lengths(&c, blk, thisbs, &lo, &hi);
                    k = k_prev < lo ? lo : (k_prev > hi ? hi : k_prev);
                    k_prev = k;                             /* encode.c:407: updated whatever option wins */
                }
                uint32_t se = se_length(&c, blk, unc);
                int opt;                                    /* encode.c:600-611 */
                if (split < unc)
                    opt = (split < se) ? 2 : 1;
                else
                    opt = (unc <= se) ? 3 : 1;
                tr.option = (uint8_t)opt; tr.k = (uint8_t)k; tr.klo = (uint8_t)lo; tr.khi = (uint8_t)hi;

                if (opt == 2) {                             /* encode.c:520-534 */
                    bw_put(&w, k + 1, c.idl);
                    if (ref) bw_put(&w, refs, c.n);
                    for (uint32_t i = (uint32_t)ref; i < c.J; i++)
                        bw_fs(&w, blk[i] >> k);             /* encode.c:118-142 */
                    if (k)
                        for (uint32_t i = (uint32_t)ref; i < c.J; i++)
                            bw_put(&w, blk[i] & ((1u << k) - 1), k);   /* encode.c:144-233 */
                } else if (opt == 1) {                      /* encode.c:547-563 */
                    bw_put(&w, 1, c.idl + 1);
                    if (ref) bw_put(&w, refs, c.n);
                    for (uint32_t i = 0; i < c.J; i += 2) {
                        uint32_t s2 = blk[i] + blk[i + 1];  /* u32 here, encode.c:558 */
                        bw_fs(&w, (uint64_t)(uint32_t)(s2 * (s2 + 1) / 2 + blk[i + 1]));
                    }
                } else {                                    /* encode.c:536-545 */
                    bw_put(&w, (1u << c.idl) - 1, c.idl);
                    bw_put(&w, ref ? refs : blk[0], c.n);
                    for (uint32_t i = 1; i < c.J; i++)
                        bw_put(&w, blk[i], c.n);
                }
            }
            tr.cds_bits = (uint32_t)(w.nbits - before);
            if (trace && ntr < trace_cap) trace[ntr] = tr;
            ntr++;
        }
        if (honour_pad_rsi && (c.flags & ORC_PAD_RSI))
            bw_align8(&w);                                  /* encode.c:499-505 */
        done += s;
    }
    free(raw);
    free(d);

    /* End of stream: pad the last byte with zero bits; an empty input still
     * produces one (zero) byte (encode.c:686-695). */
    size_t nbytes;
    if (w.nbits == 0) {
        bw_reserve(&w, 8);
        nbytes = 1;
    } else {
        nbytes = (size_t)((w.nbits + 7) / 8);
    }
    if (w.oom) {
        free(w.buf);
        return ORC_MEM_ERROR;
    }
    if (in_consumed) *in_consumed = nsamples * c.B;
    if (n_offsets) *n_offsets = noff;
    if (n_trace) *n_trace = ntr;
    size_t ncopy = nbytes < out_cap ? nbytes : out_cap;
    if (ncopy && out) memcpy(out, w.buf, ncopy);
    if (out_len) *out_len = ncopy;
    free(w.buf);
    return nbytes <= out_cap ? ORC_OK : ORC_STREAM_ERROR;   /* encode.c:944-945 */
}

/* ------------------------------------------------------------------------- */
/* Decoder                                                                    */
/* ------------------------------------------------------------------------- */

typedef struct {
    const uint8_t *buf;
    uint64_t nbits;  /* total bits available */
    uint64_t pos;
} bitr;

/* Read len bits; returns 0 when fewer than len bits remain (decode.c:342-353). */
static int br_get(bitr *r, uint32_t len, uint32_t *v)
{
    if (r->pos + len > r->nbits)
        return 0;
    uint64_t x = 0;
    for (uint32_t i = 0; i < len; i++) {
        x = (x << 1) | ((r->buf[r->pos >> 3] >> (7 - (r->pos & 7))) & 1u);
        r->pos++;
    }
    *v = (uint32_t)x;
    return 1;
}

/* Fundamental sequence: count zero bits up to the terminating one
 * (decode.c:361-379). Returns 0 when the terminator is missing. */
static int br_fs(bitr *r, uint32_t *fs)
{
    uint64_t p = r->pos;
    while (p < r->nbits) {
        if ((r->buf[p >> 3] >> (7 - (p & 7))) & 1u) {
            *fs = (uint32_t)(p - r->pos);
            r->pos = p + 1;
            return 1;
        }
        p++;
    }
    return 0;
}

typedef struct {
    const coder *c;
    uint8_t *out;
    size_t max_samples;   /* floor(out_cap / B) */
    size_t nout;          /* samples delivered */
    /* per-RSI post-processing state (decode.c:67-141) */
    size_t rsi_fill;      /* samples of the current RSI parsed so far */
    int64_t last;         /* previous output sample */
} sink;

static void store_sample(sink *s, uint32_t v)
{
    const coder *c = s->c;
    uint8_t *o = s->out + s->nout * c->B;
    if (c->msb)                                              /* decode.c:144-189 */
        for (uint32_t i = 0; i < c->B; i++) o[i] = (uint8_t)(v >> (8 * (c->B - 1 - i)));
    else
        for (uint32_t i = 0; i < c->B; i++) o[i] = (uint8_t)(v >> (8 * i));
    s->nout++;
}

/* Deliver one parsed value of the current RSI. Returns 0 when the output is
 * full (the value is then NOT consumed). */
static int deliver(sink *s, uint32_t dv)
{
    const coder *c = s->c;
    if (s->nout >= s->max_samples)
        return 0;
    if (!c->pp) {
        store_sample(s, dv);                                 /* decode.c:136-139 */
    } else if (s->rsi_fill == 0) {
        int64_t x = dv;
        if (c->sgn) {                                        /* decode.c:78-84 */
            int64_t m = (int64_t)1 << (c->n - 1);
            x = ((int64_t)(dv ^ (uint32_t)m)) - m;
            if (c->n == 32) x = (int32_t)dv;
        }
        s->last = x;
        store_sample(s, (uint32_t)(int32_t)x);
    } else {
        /* Inverse of the mapper, normalised: u = x - xmin in [0,M]
         * (decode.c:89-135 written once for both signednesses). */
        int64_t M = c->xmax - c->xmin;
        int64_t u = s->last - c->xmin;
        int64_t th = u < M - u ? u : M - u;
        int64_t h = ((int64_t)dv + 1) / 2;
        int64_t x;
        if (h <= th)
            x = s->last + ((dv & 1) ? -h : h);
        else if (u <= M - u)
            x = c->xmin + (int64_t)dv;
        else
            x = c->xmax - (int64_t)dv;
        s->last = x;
        store_sample(s, (uint32_t)(int32_t)x);
    }
    s->rsi_fill++;
    return 1;
}

int orc_decode(const orc_params *p,
               const uint8_t *in, size_t in_len,
               uint8_t *out, size_t out_cap, size_t *out_len)
{
    coder c;
    int st = setup(&c, p, 0);
    if (out_len) *out_len = 0;
    if (st != ORC_OK)
        return st;

    const size_t R = (size_t)c.rsi * c.J;
    bitr r = { in, (uint64_t)in_len * 8, 0 };
    sink s = { &c, out, out_cap / c.B, 0, 0, 0 };
    int status = ORC_OK;
    int out_full = 0;
    uint32_t v;

    if (R == 0)
        goto done;

    for (;;) {
        /* RSI start (decode.c:406-410) */
        if (s.rsi_fill == R)
            s.rsi_fill = 0;
        if (s.rsi_fill == 0 && (c.flags & ORC_PAD_RSI))
            r.pos = (r.pos + 7) & ~(uint64_t)7;
        int ref = c.pp && s.rsi_fill == 0;

        uint32_t id;
        if (!br_get(&r, c.idl, &id))
            break;
        if (id == 0) {                                       /* low entropy, decode.c:634-644 */
            uint32_t sel;
            if (!br_get(&r, 1, &sel)) break;
            if (ref) {                                       /* decode.c:618-632 */
                if (s.nout >= s.max_samples) { out_full = 1; break; }
                if (!br_get(&r, c.n, &v)) break;
                deliver(&s, v);
            }
            if (sel == 0) {                                  /* zero run, decode.c:518-558 */
                uint32_t fs;
                if (!br_fs(&r, &fs)) break;
                uint32_t zb = fs + 1;
                if (zb == 5) {                               /* ROS, decode.c:528-530 */
                    uint32_t b = (uint32_t)(s.rsi_fill / c.J);
                    uint32_t a1 = c.rsi - b, a2 = 64 - (b % 64);
                    zb = a1 < a2 ? a1 : a2;
                } else if (zb > 5) {
                    zb--;
                }
                size_t cnt = (size_t)zb * c.J - (size_t)ref;
                size_t room = s.max_samples - s.nout;
                /* fast path of the reference checks the RSI bound only when
                 * the whole run fits the output (decode.c:541-544) */
                size_t room_bytes = out_cap - s.nout * c.B;
                if (room_bytes >= cnt * c.B && R - s.rsi_fill < cnt) {
                    status = ORC_DATA_ERROR;
                    goto done;
                }
                for (size_t i = 0; i < cnt; i++) {
                    if (s.rsi_fill == R) s.rsi_fill = 0;     /* slow path wraps silently (decode.c:504-516) */
                    if (!deliver(&s, 0)) { out_full = 1; break; }
                }
                (void)room;
                if (out_full) break;
            } else {                                         /* second extension, decode.c:589-616 */
                uint32_t i = (uint32_t)ref;
                int stop = 0;
                while (i < c.J) {
                    uint32_t m;
                    if (!br_fs(&r, &m)) { stop = 1; break; }
                    /* s = max{s : s(s+1)/2 <= m}  (decode.c:679-692 as a formula) */
                    uint32_t sv = 0;
                    while ((uint64_t)(sv + 1) * (sv + 2) / 2 <= m) sv++;
                    uint32_t d1 = m - sv * (sv + 1) / 2;
                    if ((i & 1) == 0) {
                        if (!deliver(&s, sv - d1)) { out_full = 1; stop = 1; break; }
                        i++;
                    }
                    if (!deliver(&s, d1)) { out_full = 1; stop = 1; break; }
                    i++;
                }
                if (stop) break;
            }
        } else if (id == (1u << c.idl) - 1) {                /* uncompressed, decode.c:659-677 */
            int stop = 0;
            for (uint32_t i = 0; i < c.J; i++) {
                if (s.nout >= s.max_samples) { out_full = 1; stop = 1; break; }
                if (!br_get(&r, c.n, &v)) { stop = 1; break; }
                deliver(&s, v);
            }
            if (stop) break;
        } else {                                             /* split, decode.c:462-502 */
            uint32_t k = id - 1;
            if (ref) {
                if (s.nout >= s.max_samples) { out_full = 1; break; }
                if (!br_get(&r, c.n, &v)) break;
                deliver(&s, v);
            }
            uint32_t cnt = c.J - (uint32_t)ref;
            uint32_t hi[64 * 64];
            uint32_t *hv = hi;
            uint32_t *dyn = NULL;
            if (cnt > sizeof hi / sizeof hi[0]) {
                dyn = (uint32_t *)malloc(sizeof(uint32_t) * cnt);
                if (!dyn) { status = ORC_MEM_ERROR; goto done; }
                hv = dyn;
            }
            int stop = 0;
            for (uint32_t i = 0; i < cnt; i++) {             /* decode.c:444-460 */
                uint32_t fs;
                if (!br_fs(&r, &fs)) { stop = 1; break; }
                hv[i] = fs << k;
            }
            for (uint32_t i = 0; !stop && i < cnt; i++) {    /* decode.c:423-442 */
                uint32_t lowbits = 0;
                if (s.nout >= s.max_samples) { out_full = 1; stop = 1; break; }
                if (k && !br_get(&r, k, &lowbits)) { stop = 1; break; }
                deliver(&s, hv[i] + lowbits);
            }
            free(dyn);
            if (stop) break;
        }
    }

done:
    if (out_len) *out_len = s.nout * c.B;
    if (status == ORC_OK) {
        size_t left = out_cap - s.nout * c.B;
        if (left > 0 && left < c.B)                          /* decode.c:821-823 */
            status = ORC_MEM_ERROR;
    }
    return status;
}

/* ------------------------------------------------------------------------- */
/* SZIP shim (sz_compat.c)                                                    */
/* ------------------------------------------------------------------------- */

#define SZ_MSB_MASK 16
#define SZ_NN_MASK  32
#define SZ_OUTBUFF_FULL_CODE 2

static uint32_t sz_flags(int mask)                            /* sz_compat.c:12-27 */
{
    uint32_t f = 0;
    if (mask & SZ_MSB_MASK) f |= ORC_DATA_MSB;
    if (mask & SZ_NN_MASK) f |= ORC_DATA_PREPROCESS;
    return f;
}

static int sz_pixel_bytes(int bits)                           /* sz_compat.c:29-37 */
{
    return bits > 16 ? 4 : (bits > 8 ? 2 : 1);
}

int orc_sz_compress(void *dest, size_t *dest_len, const void *src, size_t src_len,
                    int options_mask, int bits_per_pixel,
                    int pixels_per_block, int pixels_per_scanline)
{
    orc_params p;
    p.block_size = (uint32_t)pixels_per_block;
    p.rsi = (uint32_t)((pixels_per_scanline + pixels_per_block - 1) / pixels_per_block);
    p.flags = ORC_NOT_ENFORCE | sz_flags(options_mask);       /* sz_compat.c:125-128 */
    int interleave = bits_per_pixel == 32 || bits_per_pixel == 64;
    p.bits_per_sample = interleave ? 8 : (uint32_t)bits_per_pixel;

    const uint8_t *s8 = (const uint8_t *)src;
    uint8_t *planes = NULL;
    if (interleave) {                                         /* sz_compat.c:39-53, :134-142 */
        size_t ws = (size_t)bits_per_pixel / 8, nw = src_len / ws;
        planes = (uint8_t *)malloc(src_len ? src_len : 1);
        if (!planes) return ORC_MEM_ERROR;
        for (size_t i = 0; i < nw; i++)
            for (size_t j = 0; j < ws; j++)
                planes[j * nw + i] = s8[i * ws + j];
        s8 = planes;
    }
    size_t px = (size_t)sz_pixel_bytes((int)p.bits_per_sample);
    size_t line = (size_t)pixels_per_scanline * px;
    size_t padded_line = (size_t)p.rsi * p.block_size * px;
    size_t scanlines = (src_len / px + (size_t)pixels_per_scanline - 1) / (size_t)pixels_per_scanline;
    size_t pb_size = padded_line * scanlines;                 /* sz_compat.c:152-155 */
    uint8_t *pb = (uint8_t *)malloc(pb_size ? pb_size : 1);
    if (!pb) { free(planes); return ORC_MEM_ERROR; }

    /* add_padding (sz_compat.c:71-94): each scanline is filled up to a whole
     * number of blocks with its last pixel (NN) or zero. */
    size_t i = 0, j = 0;
    uint8_t zero_px[4] = {0, 0, 0, 0};
    while (i < src_len) {
        size_t ls = src_len - i < line ? src_len - i : line;
        memcpy(pb + j, s8 + i, ls);
        j += ls;
        i += ls;
        const uint8_t *fill = (p.flags & ORC_DATA_PREPROCESS) ? s8 + i - px : zero_px;
        size_t ps = padded_line - ls;
        for (size_t k = 0; k < ps; k += px)
            memcpy(pb + j + k, fill, px);
        j += ps;
    }

    size_t outl = 0;
    int st = orc_encode(&p, 0, pb, pb_size, (uint8_t *)dest, *dest_len, &outl,
                        NULL, NULL, 0, NULL, NULL, 0, NULL);
    *dest_len = outl;                                         /* sz_compat.c:175 */
    free(pb);
    free(planes);
    return st == ORC_STREAM_ERROR ? SZ_OUTBUFF_FULL_CODE : st;   /* sz_compat.c:171-174 */
}

int orc_sz_decompress(void *dest, size_t *dest_len, const void *src, size_t src_len,
                      int options_mask, int bits_per_pixel,
                      int pixels_per_block, int pixels_per_scanline)
{
    orc_params p;
    p.block_size = (uint32_t)pixels_per_block;
    p.rsi = (uint32_t)((pixels_per_scanline + pixels_per_block - 1) / pixels_per_block);
    p.flags = sz_flags(options_mask);                         /* sz_compat.c:203 */
    int pad_scanline = pixels_per_scanline % pixels_per_block;
    int deint = bits_per_pixel == 32 || bits_per_pixel == 64;
    int extra = pad_scanline || deint;
    p.bits_per_sample = deint ? 8 : (uint32_t)bits_per_pixel;
    size_t px = (size_t)sz_pixel_bytes((int)p.bits_per_sample);
    size_t padded_line = (size_t)p.rsi * p.block_size * px;
    size_t line = (size_t)pixels_per_scanline * px;
    size_t scanlines = 0, buf_size = *dest_len;
    uint8_t *buf = (uint8_t *)dest;

    if (extra) {                                              /* sz_compat.c:219-237 */
        if (pad_scanline) {
            scanlines = (*dest_len / px + (size_t)pixels_per_scanline - 1) / (size_t)pixels_per_scanline;
            buf_size = padded_line * scanlines;
        }
        buf = (uint8_t *)malloc(buf_size ? buf_size : 1);
        if (!buf) return ORC_MEM_ERROR;
    }
    size_t produced = 0;
    int st = orc_decode(&p, (const uint8_t *)src, src_len, buf, buf_size, &produced);
    if (st != ORC_OK) {
        if (extra) free(buf);
        return st;
    }
    size_t total_out = produced;
    if (pad_scanline) {                                       /* remove_padding, sz_compat.c:96-108 */
        size_t wi = line;
        for (size_t rj = padded_line; rj < produced; rj += padded_line) {
            memmove(buf + wi, buf + rj, line);
            wi += line;
        }
        total_out = scanlines * line;                         /* sz_compat.c:250 */
    }
    if (total_out < *dest_len)
        *dest_len = total_out;                                /* sz_compat.c:255-256 */
    if (deint) {                                              /* sz_compat.c:55-69 */
        size_t ws = (size_t)bits_per_pixel / 8, nw = *dest_len / ws;
        uint8_t *d8 = (uint8_t *)dest;
        for (size_t a = 0; a < nw; a++)
            for (size_t b = 0; b < ws; b++)
                d8[a * ws + b] = buf[b * nw + a];
    } else if (pad_scanline) {
        memcpy(dest, buf, *dest_len);
    }
    if (extra) free(buf);
    return ORC_OK;
}
