/*
 * libaec_api.c -- the libaec stream API (include/libaec.h) in C, on top of the
 * device layer (include/aec_b200.h).
 *
 * Drop-in for the eight entry points of the reference
 * (/root/reference/src/encode.c:773-963, src/decode.c:694-854).  The
 * reference drives two resumable per-sample state machines from these calls;
 * here the calls only do buffer bookkeeping on the host:
 *
 *   encode: input is collected -- on the device -- until at least one whole RSI (or AEC_FLUSH) is
 *           available, whole RSIs are coded on the GPU in one launch with the
 *           (bit phase, k) carry of the stream so far, and the produced bytes
 *           are handed out over as many calls as the caller's windows need.
 *   decode: input is buffered, every attempt decodes as far as the buffered
 *           bits allow (RSI-parallel on the GPU), decoded samples are queued
 *           and handed out; an attempt is only made when new input arrived.
 *
 * Only the concatenation of the output windows is observable, and it is
 * identical to the reference's (README.md:151-159 of the reference).
 * There is no CPU coder in here: without a CUDA device init fails.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/aec_b200.h"
#include "../../include/libaec.h"

#define DEC_AHEAD_BYTES ((size_t)8 << 20)   /* decode-ahead window of the streaming decoder */
#define DEC_DIRECT_BYTES ((size_t)1 << 20)  /* input pieces of this size are decoded straight from the caller's buffer */

struct internal_state {
    int decoder;
    aecb200_ctx *ctx;
    aecb200_params prm;
    size_t B;              /* storage bytes per sample */
    size_t rsi_bytes;      /* bytes of one RSI of input */

    /* pending output (both directions) */
    unsigned char *q;
    size_t qhead, qtail, qcap;

    /* encoder */
    size_t ilen;           /* bytes of whole samples (less than one RSI) accumulated on the device so far */
    aecb200_carry carry;
    uint64_t emitted_bytes;   /* complete stream bytes produced so far */
    int any_samples;
    int flush, finished, flushed;
    int want_offsets;
    size_t *offs;
    size_t noffs, offcap;

    /* decoder */
    unsigned char *cbuf;   /* compressed bytes not yet fully decoded */
    size_t clen, ccap;
    uint64_t cbase_bits;   /* stream bit position of cbuf[0] */
    uint64_t rsi_bit;      /* stream bit where the RSI of the next sample starts */
    size_t rsi_delivered;  /* samples of that RSI already delivered */
    int new_input;
    int oneshot;           /* aec_buffer_decode: no state needed after the call */
    const size_t *doffs;
    size_t ndoffs;
    int data_error;
};

/* ---- context pool: contexts (stream + workspace) outlive streams ---- */
static pthread_mutex_t pool_lock = PTHREAD_MUTEX_INITIALIZER;
static aecb200_ctx *pool[16];
static int pool_n;

/* a stream runs on the device that is current when it is initialised: only contexts of that device
 * are reused */
static aecb200_ctx *ctx_get(void)
{
    aecb200_ctx *c = NULL;
    const int dev = aecb200_current_device();
    pthread_mutex_lock(&pool_lock);
    for (int i = pool_n - 1; i >= 0; i--) {
        if (aecb200_ctx_device(pool[i]) == dev) {
            c = pool[i];
            pool[i] = pool[--pool_n];
            break;
        }
    }
    pthread_mutex_unlock(&pool_lock);
    if (!c && aecb200_ctx_create(&c, dev) != 0) return NULL;
    return c;
}

static void ctx_put(aecb200_ctx *c)
{
    if (!c) return;
    pthread_mutex_lock(&pool_lock);
    if (pool_n < 16) { pool[pool_n++] = c; c = NULL; }
    pthread_mutex_unlock(&pool_lock);
    if (c) aecb200_ctx_destroy(c);
}

/* the same pool for callers outside this file (the SZIP shim, batch calls) */
aecb200_ctx *aecb200_pool_get(void) { return ctx_get(); }
void aecb200_pool_put(aecb200_ctx *c) { ctx_put(c); }

static unsigned bytes_per_sample(const struct aec_stream *s)
{
    unsigned n = s->bits_per_sample;
    if (n > 16) return (n <= 24 && (s->flags & AEC_DATA_3BYTE)) ? 3 : 4;
    return n > 8 ? 2 : 1;
}

/* parameter checks shared with the device layer; mirrors encode.c:777-794,
 * :843-851 and decode.c:699, :739-747 */
static int check_params(const struct aec_stream *s, int enc)
{
    if (s->bits_per_sample == 0 || s->bits_per_sample > 32) return AEC_CONF_ERROR;
    if (enc) {
        if (s->flags & AEC_NOT_ENFORCE) { if (s->block_size & 1) return AEC_CONF_ERROR; }
        else if (s->block_size != 8 && s->block_size != 16 && s->block_size != 32 && s->block_size != 64)
            return AEC_CONF_ERROR;
        if (s->rsi > 4096) return AEC_CONF_ERROR;
    }
    if (s->bits_per_sample <= 8 && (s->flags & AEC_RESTRICTED) && s->bits_per_sample > 4)
        return AEC_CONF_ERROR;
    /* outside the reference's contract (undefined there): refuse */
    if (s->block_size == 0 || s->block_size > 64 || (s->block_size & 1) || s->rsi == 0 || s->rsi > 4096)
        return AEC_CONF_ERROR;
    return AEC_OK;
}

static struct internal_state *state_new(struct aec_stream *strm, int decoder)
{
    struct internal_state *st = (struct internal_state *)calloc(1, sizeof *st);
    if (!st) return NULL;
    st->decoder = decoder;
    st->prm.bits_per_sample = strm->bits_per_sample;
    st->prm.block_size = strm->block_size;
    st->prm.rsi = strm->rsi;
    st->prm.flags = strm->flags;
    st->B = bytes_per_sample(strm);
    st->rsi_bytes = (size_t)strm->rsi * strm->block_size * st->B;
    st->ctx = ctx_get();
    if (!st->ctx) { free(st); return NULL; }
    return st;
}

static void state_free(struct internal_state *st)
{
    if (!st) return;
    ctx_put(st->ctx);
    free(st->q); free(st->offs); free(st->cbuf);
    free(st);
}

static int q_reserve(struct internal_state *st, size_t extra)
{
    if (st->qhead == st->qtail) st->qhead = st->qtail = 0;
    if (st->qtail + extra <= st->qcap) return 0;
    size_t live = st->qtail - st->qhead;
    size_t ncap = live + extra + 64;
    unsigned char *nq = (unsigned char *)malloc(ncap);
    if (!nq) return -1;
    if (live) memcpy(nq, st->q + st->qhead, live);
    free(st->q);
    st->q = nq; st->qcap = ncap; st->qhead = 0; st->qtail = live;
    return 0;
}

/* hand queued bytes to the caller; `gran` = delivery granularity in bytes */
static void q_drain(struct aec_stream *strm, struct internal_state *st, size_t gran)
{
    size_t live = st->qtail - st->qhead;
    size_t n = live < strm->avail_out ? live : strm->avail_out;
    n -= n % gran;
    if (!n) return;
    memcpy(strm->next_out, st->q + st->qhead, n);
    st->qhead += n;
    strm->next_out += n;
    strm->avail_out -= n;
    strm->total_out += n;
}

/* ------------------------------------------------------------------------ */
/* encoder                                                                   */
/* ------------------------------------------------------------------------ */

int aec_encode_init(struct aec_stream *strm)
{
    int rc = check_params(strm, 1);
    if (rc != AEC_OK) return rc;
    struct internal_state *st = state_new(strm, 0);
    if (!st) return AEC_MEM_ERROR;
    strm->state = st;
    strm->total_in = 0;
    strm->total_out = 0;
    return AEC_OK;
}

/* code one piece (whole RSIs, or the rest of the stream when final) */
static int encode_piece(struct aec_stream *strm, struct internal_state *st,
                        const unsigned char *in, size_t len, int final)
{
    if (final && len < st->B && st->carry.bits == 0 && st->any_samples)
        return AEC_OK;          /* stream already ends on a byte boundary: nothing left to emit */
    size_t bound = aecb200_encode_bound(&st->prm, len) + 16;
    size_t nrsi = st->rsi_bytes ? (len / st->B + (st->rsi_bytes / st->B) - 1) / (st->rsi_bytes / st->B) : 0;
    uint64_t *poffs = NULL;
    if (st->want_offsets && nrsi) {
        if (st->noffs + nrsi > st->offcap) {
            size_t ncap = (st->noffs + nrsi) * 2;
            size_t *no = (size_t *)realloc(st->offs, ncap * sizeof(size_t));
            if (!no) return AEC_MEM_ERROR;
            st->offs = no; st->offcap = ncap;
        }
        poffs = (uint64_t *)(st->offs + st->noffs);
    }
    size_t produced = 0, consumed = 0, noff = 0;
    unsigned char *dst;
    int direct = (st->qhead == st->qtail) && strm->avail_out >= bound;
    if (direct) dst = strm->next_out;
    else {
        if (q_reserve(st, bound)) return AEC_MEM_ERROR;
        dst = st->q + st->qtail;
    }
    int rc = aecb200_encode_host_piece(st->ctx, &st->prm, in, len, final, dst, bound, &produced, &consumed,
                                       &st->carry, poffs, nrsi, &noff);
    if (rc != AEC_OK) return rc == AECB200_CUDA_ERROR ? AEC_MEM_ERROR : rc;
    if (poffs) {
        for (size_t i = 0; i < noff; i++) st->offs[st->noffs + i] = (size_t)(poffs[i] + st->emitted_bytes * 8);
        st->noffs += noff;
    }
    if (len >= st->B) st->any_samples = 1;
    st->emitted_bytes += final ? produced : produced;   /* complete bytes (the final one may be padded) */
    if (direct) { strm->next_out += produced; strm->avail_out -= produced; strm->total_out += produced; }
    else st->qtail += produced;
    return AEC_OK;
}

int aec_encode(struct aec_stream *strm, int flush)
{
    struct internal_state *st = strm->state;
    st->flush = flush;
    for (;;) {
        q_drain(strm, st, 1);
        if (st->qhead != st->qtail) break;              /* caller's window is full */
        if (st->finished) { st->flushed = 1; break; }
        size_t whole = (strm->avail_in / st->B) * st->B;
        int rc;
        if (st->ilen == 0 && (whole >= st->rsi_bytes || flush == AEC_FLUSH)) {
            /* code straight from the caller's buffer */
            size_t len = flush == AEC_FLUSH ? whole : (whole / st->rsi_bytes) * st->rsi_bytes;
            rc = encode_piece(strm, st, strm->next_in, len, flush == AEC_FLUSH);
            if (rc != AEC_OK) return rc;
            strm->next_in += len; strm->avail_in -= len; strm->total_in += len;
            if (flush == AEC_FLUSH) st->finished = 1;
            continue;
        }
        /* less than an RSI at hand: the samples accumulate ON THE DEVICE (the context's input stage)
         * until the RSI is complete, then they are coded from there without another upload */
        size_t take = st->rsi_bytes - st->ilen;
        if (take > whole) take = whole;
        if (take) {
            rc = aecb200_ctx_stage_input(st->ctx, st->ilen, strm->next_in, take);
            if (rc != AEC_OK) return rc == AECB200_CUDA_ERROR ? AEC_MEM_ERROR : rc;
            st->ilen += take;
            strm->next_in += take; strm->avail_in -= take; strm->total_in += take;
        }
        if (st->ilen == st->rsi_bytes) {
            rc = encode_piece(strm, st, NULL, st->ilen, 0);
            if (rc != AEC_OK) return rc;
            st->ilen = 0;
            continue;
        }
        if (flush == AEC_FLUSH && strm->avail_in < st->B) {
            rc = encode_piece(strm, st, NULL, st->ilen, 1);
            if (rc != AEC_OK) return rc;
            st->ilen = 0;
            st->finished = 1;
            continue;
        }
        if (!take) break;                               /* need more input */
    }
    return AEC_OK;
}

int aec_encode_end(struct aec_stream *strm)
{
    struct internal_state *st = strm->state;
    int status = AEC_OK;
    if (st->flush == AEC_FLUSH && !st->flushed) status = AEC_STREAM_ERROR;   /* encode.c:944-945 */
    state_free(st);
    strm->state = NULL;
    return status;
}

int aec_buffer_encode(struct aec_stream *strm)
{
    int status = aec_encode_init(strm);
    if (status != AEC_OK) return status;
    status = aec_encode(strm, AEC_FLUSH);
    if (status != AEC_OK) { state_free(strm->state); strm->state = NULL; return status; }
    return aec_encode_end(strm);
}

int aec_encode_enable_offsets(struct aec_stream *strm)
{
    if (!strm->state || strm->state->decoder) return AEC_CONF_ERROR;
    strm->state->want_offsets = 1;
    return AEC_OK;
}

int aec_encode_count_offsets(struct aec_stream *strm, size_t *count)
{
    if (!strm->state || !strm->state->want_offsets || !count) return AEC_CONF_ERROR;
    *count = strm->state->noffs;
    return AEC_OK;
}

int aec_encode_get_offsets(struct aec_stream *strm, size_t *offsets, size_t offsets_count)
{
    struct internal_state *st = strm->state;
    if (!st || !st->want_offsets || !offsets) return AEC_CONF_ERROR;
    if (offsets_count < st->noffs) return AEC_MEM_ERROR;
    memcpy(offsets, st->offs, st->noffs * sizeof(size_t));
    return AEC_OK;
}

/* ------------------------------------------------------------------------ */
/* decoder                                                                   */
/* ------------------------------------------------------------------------ */

int aec_decode_init(struct aec_stream *strm)
{
    int rc = check_params(strm, 0);
    if (rc != AEC_OK) return rc;
    struct internal_state *st = state_new(strm, 1);
    if (!st) return AEC_MEM_ERROR;
    aecb200_ctx_accumulate_next(st->ctx, -1);           /* a pooled context may still hold another stream */
    strm->state = st;
    strm->total_in = 0;
    strm->total_out = 0;
    return AEC_OK;
}

int aec_decode_set_offsets(struct aec_stream *strm, const size_t *offsets, size_t offsets_count)
{
    struct internal_state *st = strm->state;
    if (!st || !st->decoder) return AEC_CONF_ERROR;
    st->doffs = offsets;
    st->ndoffs = offsets_count;
    return AEC_OK;
}

int aec_decode_enable_offsets(struct aec_stream *strm)
{
    if (!strm->state || !strm->state->decoder) return AEC_CONF_ERROR;
    strm->state->want_offsets = 1;
    return AEC_OK;
}

int aec_decode_count_offsets(struct aec_stream *strm, size_t *count)
{
    if (!strm->state || !strm->state->decoder || !strm->state->want_offsets || !count) return AEC_CONF_ERROR;
    *count = strm->state->noffs;
    return AEC_OK;
}

int aec_decode_get_offsets(struct aec_stream *strm, size_t *offsets, size_t offsets_count)
{
    struct internal_state *st = strm->state;
    if (!st || !st->decoder || !st->want_offsets || !offsets) return AEC_CONF_ERROR;
    if (offsets_count < st->noffs) return AEC_MEM_ERROR;
    memcpy(offsets, st->offs, st->noffs * sizeof(size_t));
    return AEC_OK;
}

int aec_decode_range(struct aec_stream *strm, const size_t *rsi_offsets, size_t rsi_offsets_count,
                     size_t pos, size_t size)
{
    struct internal_state *st = strm->state;
    if (!st || !st->decoder || !rsi_offsets) return AEC_CONF_ERROR;
    if (pos % st->B || size % st->B) return AEC_CONF_ERROR;
    if (strm->avail_out < size) return AEC_MEM_ERROR;
    if (size == 0) return AEC_OK;
    const size_t r0 = pos / st->rsi_bytes;             /* RSI that holds the first sample wanted */
    if (r0 >= rsi_offsets_count) return AEC_DATA_ERROR;
    /* the stream from the word that holds that RSI's first bit up to the first bit of the RSI after the
     * last one wanted: nothing else is uploaded or decoded */
    const size_t r1 = (pos + size + st->rsi_bytes - 1) / st->rsi_bytes;
    size_t end_byte = strm->avail_in;
    if (r1 < rsi_offsets_count) {
        end_byte = (size_t)(rsi_offsets[r1] / 8) + 8;
        if (end_byte > strm->avail_in) end_byte = strm->avail_in;
    }
    size_t got = 0, rdel = 0;
    uint64_t rbit = 0;
    int rc = aecb200_decode_host_resume(st->ctx, &st->prm, strm->next_in, end_byte,
                                        (const uint64_t *)rsi_offsets, rsi_offsets_count,
                                        (uint64_t)rsi_offsets[r0], (pos % st->rsi_bytes) / st->B,
                                        strm->next_out, size, &got, &rbit, &rdel);
    if (rc != AEC_OK) return rc == AECB200_CUDA_ERROR ? AEC_MEM_ERROR : rc;
    strm->next_out += got; strm->avail_out -= got; strm->total_out += got;
    return got == size ? AEC_OK : AEC_DATA_ERROR;       /* the stream ends before the range does */
}

static int cbuf_append(struct internal_state *st, const unsigned char *p, size_t n)
{
    if (st->clen + n > st->ccap) {
        size_t ncap = (st->clen + n) * 2 + 64;
        unsigned char *nb = (unsigned char *)realloc(st->cbuf, ncap);
        if (!nb) return -1;
        st->cbuf = nb; st->ccap = ncap;
    }
    memcpy(st->cbuf + st->clen, p, n);
    st->clen += n;
    return 0;
}

/* one decode attempt over `in` (stream bytes starting at stream bit base_bits) */
static int decode_attempt(struct aec_stream *strm, struct internal_state *st,
                          const unsigned char *in, size_t in_len, uint64_t base_bits, int *filled, int accumulate)
{
    size_t want = (strm->avail_out / st->B) * st->B;
    int direct = want >= DEC_AHEAD_BYTES || st->oneshot;
    unsigned char *dst;
    if (direct) dst = strm->next_out;
    else {
        want = DEC_AHEAD_BYTES - DEC_AHEAD_BYTES % st->B;
        if (q_reserve(st, want)) return AEC_MEM_ERROR;
        dst = st->q + st->qtail;
    }
    size_t got = 0, rdel = st->rsi_delivered;
    uint64_t rbit = st->rsi_bit - base_bits;
    /* the caller's index counts from stream bit 0; the device layer wants it relative to in[0] */
    const uint64_t *idx = (const uint64_t *)st->doffs;
    size_t nidx = st->ndoffs;
    uint64_t *rebased = NULL;
    if (idx && base_bits) {
        size_t first = 0;
        while (first < nidx && idx[first] < st->rsi_bit) first++;
        nidx -= first;
        rebased = (uint64_t *)malloc((nidx ? nidx : 1) * sizeof(uint64_t));
        if (!rebased) return AEC_MEM_ERROR;
        for (size_t i = 0; i < nidx; i++) rebased[i] = idx[first + i] - base_bits;
        idx = rebased;
    }
    /* the buffered stream stays on the device between attempts: only new bytes are uploaded */
    if (accumulate) aecb200_ctx_accumulate_next(st->ctx, (long long)(base_bits / 8));
    int rc = aecb200_decode_host_resume(st->ctx, &st->prm, in, in_len, idx, nidx,
                                        st->rsi_bit - base_bits, st->rsi_delivered,
                                        dst, want, &got, &rbit, &rdel);
    free(rebased);
    if (rc == AEC_DATA_ERROR) { st->data_error = 1; return AEC_DATA_ERROR; }
    if (rc != AEC_OK) return rc == AECB200_CUDA_ERROR ? AEC_MEM_ERROR : rc;
    st->rsi_bit = rbit + base_bits;
    st->rsi_delivered = rdel;
    if (st->want_offsets && !st->doffs) {
        /* remember the RSI boundaries the device discovered (an attempt may see some of them again) */
        size_t n = aecb200_ctx_found_offsets(st->ctx, NULL, 0);
        if (n) {
            uint64_t *tmp = (uint64_t *)malloc(n * sizeof(uint64_t));
            if (!tmp) return AEC_MEM_ERROR;
            aecb200_ctx_found_offsets(st->ctx, tmp, n);
            if (st->noffs + n > st->offcap) {
                size_t ncap = (st->noffs + n) * 2;
                size_t *no = (size_t *)realloc(st->offs, ncap * sizeof(size_t));
                if (!no) { free(tmp); return AEC_MEM_ERROR; }
                st->offs = no; st->offcap = ncap;
            }
            for (size_t i = 0; i < n; i++) {
                size_t o = (size_t)(tmp[i] + base_bits);
                if (st->noffs == 0 || o > st->offs[st->noffs - 1]) st->offs[st->noffs++] = o;
            }
            free(tmp);
        }
    }
    *filled = (got == want);
    if (direct) { strm->next_out += got; strm->avail_out -= got; strm->total_out += got; }
    else st->qtail += got;
    return AEC_OK;
}

int aec_decode(struct aec_stream *strm, int flush)
{
    struct internal_state *st = strm->state;
    (void)flush;                                        /* ignored by the reference too (decode.c:797) */
    if (st->data_error) return AEC_DATA_ERROR;
    for (;;) {
        q_drain(strm, st, st->B);
        if (st->qhead != st->qtail) break;              /* caller's window is full */
        int filled = 0, rc;
        if (st->oneshot && st->clen == 0) {
            /* whole stream is in the caller's buffer: decode from it directly */
            if (strm->avail_in == 0 || strm->avail_out < st->B) break;
            rc = decode_attempt(strm, st, strm->next_in, strm->avail_in, 0, &filled, 0);
            if (rc != AEC_OK) return rc;
            strm->total_in += strm->avail_in;
            strm->next_in += strm->avail_in; strm->avail_in = 0;
            break;
        }
        if (st->clen == 0 && strm->avail_in >= DEC_DIRECT_BYTES && strm->avail_out >= st->B) {
            /* nothing buffered and a large piece of the stream at hand: decode from the caller's
             * buffer, then keep only the bytes from the RSI of the next undelivered sample on */
            const size_t n = strm->avail_in;
            rc = decode_attempt(strm, st, strm->next_in, n, st->cbase_bits, &filled, 0);
            if (rc != AEC_OK) return rc;
            size_t keep_from = (size_t)(((st->rsi_bit - st->cbase_bits) >> 5) << 2);
            if (keep_from > n) keep_from = n - n % 4;
            if (cbuf_append(st, strm->next_in + keep_from, n - keep_from)) return AEC_MEM_ERROR;
            st->cbase_bits += (uint64_t)keep_from * 8;
            strm->total_in += n; strm->next_in += n; strm->avail_in = 0;
            st->new_input = filled;
            continue;
        }
        if (strm->avail_in) {
            if (cbuf_append(st, strm->next_in, strm->avail_in)) return AEC_MEM_ERROR;
            strm->total_in += strm->avail_in;
            strm->next_in += strm->avail_in; strm->avail_in = 0;
            st->new_input = 1;
        }
        if (!st->new_input || strm->avail_out < st->B) break;
        rc = decode_attempt(strm, st, st->cbuf, st->clen, st->cbase_bits, &filled, 1);
        if (rc != AEC_OK) return rc;
        st->new_input = filled;                         /* more may be decodable without new input */
        /* forget bytes in front of the current RSI (keep 32-bit alignment of the base) */
        {
            uint64_t keep_from = ((st->rsi_bit - st->cbase_bits) >> 5) << 2;
            if (keep_from > (1u << 20) || keep_from == st->clen) {
                memmove(st->cbuf, st->cbuf + keep_from, st->clen - keep_from);
                st->clen -= keep_from;
                st->cbase_bits += keep_from * 8;
            }
        }
    }
    if (strm->avail_out > 0 && strm->avail_out < st->B) return AEC_MEM_ERROR;   /* decode.c:821-823 */
    return AEC_OK;
}

int aec_decode_end(struct aec_stream *strm)
{
    state_free(strm->state);
    strm->state = NULL;
    return AEC_OK;
}

int aec_buffer_decode(struct aec_stream *strm)
{
    int status = aec_decode_init(strm);
    if (status != AEC_OK) return status;
    strm->state->oneshot = 1;
    status = aec_decode(strm, AEC_FLUSH);
    aec_decode_end(strm);
    return status;
}
