/*
 * aec_runtime.cu -- host runtime behind the C ABI of include/aec_b200.h.
 *
 * Owns the per-context CUDA stream, the device workspace (look-back
 * descriptors, boundary words, staging buffers for host-pointer calls) and the
 * pinned result mailboxes, and turns one C call into the kernel sequence
 *   encode:  memset(descriptors) -> aec_encode_kernel -> aec_encode_fixup_kernel
 *   decode:  [aec_scan_offsets_kernel] -> aec_decode_kernel
 * There is no CPU implementation of the coder in this library: every path goes
 * through the kernels or fails.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>

#include "../../include/aec_b200.h"
#include "aec_device.h"

#define AEC_OK 0
#define AEC_CONF_ERROR (-1)
#define AEC_STREAM_ERROR (-2)
#define AEC_DATA_ERROR (-3)
#define AEC_MEM_ERROR (-4)

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

} // namespace

struct aecb200_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int honour_pad = 0;
    uint64_t launches = 0;
    char err[256] = {0};

    DevBuf grp, rsi_list, desc, pref, headc, tailc, tile_end, tile_kagg, misc, in_stage, out_stage, offs, rsi_count;
    uint64_t tile_limit = 0;             /* next encode codes only this many leading tiles (k repair) */
    bool want_summary = false;
    bool careful_only = false;           /* decode with the lane-per-RSI kernel only (tests) */
    uint64_t *h_res = nullptr;           /* pinned: [0..3] encode result, [4..7] decode result */

    /* bookkeeping of the last enqueued operation */
    uint64_t enc_out_cap_bits = 0;
    bool enc_pending = false;
    bool dec_pending = false;
    uint64_t dec_out_samples = 0;
    uint64_t dec_expect = 0;
    uint32_t dec_B = 1;
};

namespace {

int fail_cuda(aecb200_ctx *c, cudaError_t e, const char *what)
{
    snprintf(c->err, sizeof c->err, "%s: %s", what, cudaGetErrorString(e));
    fprintf(stderr, "aecb200: CUDA failure in %s\n", c->err);     /* never silent: there is no fallback path */
    return AECB200_CUDA_ERROR;
}
#define CK(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail_cuda(ctx, e_, what); } while (0)

uint32_t next_pow2(uint32_t v)
{
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

int make_cfg(aecb200_ctx *ctx, const aecb200_params *p, int enc, AecCfg *c)
{
    if (aec_cfg_init(c, p->bits_per_sample, p->block_size, p->rsi, p->flags, enc, ctx->honour_pad) != 0)
        return AEC_CONF_ERROR;
    if (c->J == 0 || c->J > AEC_MAX_J || (c->J & 1u) || c->rsi == 0 || c->rsi > 4096) {
        /* the reference has undefined behaviour for these; we refuse */
        snprintf(ctx->err, sizeof ctx->err, "unsupported block_size/rsi %u/%u", c->J, c->rsi);
        return AEC_CONF_ERROR;
    }
    return AEC_OK;
}

struct EncGeom { uint64_t nsamples, nrsi, ntiles; uint32_t last_nblk, RP, TB; };

EncGeom enc_geometry(const AecCfg &c, size_t in_bytes)
{
    EncGeom g;
    g.nsamples = in_bytes / c.B;
    g.nrsi = (g.nsamples + c.R - 1) / c.R;
    uint64_t last_s = g.nsamples - (g.nrsi ? (g.nrsi - 1) * (uint64_t)c.R : 0);
    g.last_nblk = (uint32_t)((last_s + c.J - 1) / c.J);
    g.TB = aec_encode_tile_blocks(c.J);
    g.RP = c.rsi <= g.TB ? next_pow2(c.rsi) : ((c.rsi + g.TB - 1) / g.TB) * g.TB;
    g.ntiles = (g.nrsi * (uint64_t)g.RP + g.TB - 1) / g.TB;
    return g;
}

} // namespace

extern "C" {

int aecb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int aecb200_ctx_create(aecb200_ctx **out, int device)
{
    if (!out) return AEC_CONF_ERROR;
    *out = nullptr;
    aecb200_ctx *ctx = new (std::nothrow) aecb200_ctx();
    if (!ctx) return AEC_MEM_ERROR;
    cudaError_t e;
    if (device < 0) {
        e = cudaGetDevice(&device);
        if (e != cudaSuccess) { delete ctx; return AECB200_CUDA_ERROR; }
    }
    ctx->device = device;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) { ctx->own_stream = true; e = cudaMallocHost(&ctx->h_res, 16 * sizeof(uint64_t)); }
    if (e != cudaSuccess) {
        fprintf(stderr, "aecb200: no usable CUDA device (%s); this library has no CPU fallback\n",
                cudaGetErrorString(e));
        if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
        return AECB200_CUDA_ERROR;
    }
    *out = ctx;
    return AEC_OK;
}

void aecb200_ctx_destroy(aecb200_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    ctx->tile_kagg.release(); ctx->pref.release(); ctx->grp.release(); ctx->rsi_list.release();
    ctx->desc.release(); ctx->headc.release(); ctx->tailc.release(); ctx->tile_end.release();
    ctx->misc.release(); ctx->in_stage.release(); ctx->out_stage.release(); ctx->offs.release();
    ctx->rsi_count.release();
    if (ctx->h_res) cudaFreeHost(ctx->h_res);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int aecb200_ctx_set_stream(aecb200_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return AEC_CONF_ERROR;
    if (ctx->own_stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    ctx->stream = (cudaStream_t)cuda_stream;
    return AEC_OK;
}

const char *aecb200_last_error(aecb200_ctx *ctx) { return ctx ? ctx->err : "no context"; }
void aecb200_ctx_set_encode_padding(aecb200_ctx *ctx, int on) { if (ctx) ctx->honour_pad = on ? 1 : 0; }
uint64_t aecb200_ctx_launches(aecb200_ctx *ctx) { return ctx ? ctx->launches : 0; }

size_t aecb200_encode_bound(const aecb200_params *p, size_t in_bytes)
{
    AecCfg c;
    if (aec_cfg_init(&c, p->bits_per_sample, p->block_size, p->rsi, p->flags, 0, 0) != 0 || c.J == 0 || c.rsi == 0)
        return in_bytes + in_bytes / 4 + 1024;
    uint64_t ns = in_bytes / c.B;
    uint64_t nblk = (ns + c.J - 1) / c.J + c.rsi;
    uint64_t bits = nblk * (c.idl + 1ull + (uint64_t)(c.J + 1) * c.n);   /* SURVEY App. A */
    return (size_t)(bits / 8 + nblk / c.rsi + 64);
}

/* ------------------------------------------------------------------------ */
/* device-resident                                                           */
/* ------------------------------------------------------------------------ */

int aecb200_encode_device(aecb200_ctx *ctx, const aecb200_params *p,
                          const void *d_in, size_t in_bytes,
                          void *d_out, size_t out_cap,
                          const aecb200_carry *carry, uint64_t *d_rsi_offsets)
{
    return aecb200_encode_device_indexed(ctx, p, d_in, in_bytes, d_out, out_cap, carry, d_rsi_offsets, nullptr);
}

size_t aecb200_group_index_entries(const aecb200_params *p, size_t in_bytes)
{
    AecCfg c;
    if (aec_cfg_init(&c, p->bits_per_sample, p->block_size, p->rsi, p->flags, 0, 0) != 0 || c.R == 0) return 0;
    size_t ns = in_bytes / c.B;
    return ((ns + c.R - 1) / c.R) * 32;
}

int aecb200_encode_device_indexed(aecb200_ctx *ctx, const aecb200_params *p,
                                  const void *d_in, size_t in_bytes,
                                  void *d_out, size_t out_cap,
                                  const aecb200_carry *carry, uint64_t *d_rsi_offsets,
                                  uint64_t *d_grp_index)
{
    if (!ctx || !p) return AEC_CONF_ERROR;
    AecCfg c;
    int rc = make_cfg(ctx, p, 1, &c);
    if (rc != AEC_OK) return rc;
    if (((uintptr_t)d_out & 3u) != 0) {
        snprintf(ctx->err, sizeof ctx->err, "device output must be 4-byte aligned");
        return AEC_CONF_ERROR;
    }
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    EncGeom g = enc_geometry(c, in_bytes);
    aecb200_carry zero = {0, 0, 0};
    if (!carry) carry = &zero;

    ctx->enc_out_cap_bits = (uint64_t)out_cap * 8ull;
    ctx->enc_pending = true;
    if (g.nsamples == 0) {
        /* nothing to code: the result is the carry itself */
        ctx->h_res[0] = carry->bits; ctx->h_res[1] = carry->k;
        return AEC_OK;
    }

    if (g.ntiles >= 0xFFFFFFF0ull) {
        snprintf(ctx->err, sizeof ctx->err, "input too large for one launch (%llu tiles)", (unsigned long long)g.ntiles);
        return AEC_CONF_ERROR;
    }
    CK(ctx->desc.ensure(g.ntiles * 8), "cudaMalloc(desc)");
    CK(ctx->pref.ensure(g.ntiles * 8), "cudaMalloc(pref)");
    CK(ctx->headc.ensure(g.ntiles * 4), "cudaMalloc(head)");
    CK(ctx->tailc.ensure(g.ntiles * 4), "cudaMalloc(tail)");
    CK(ctx->tile_end.ensure(g.ntiles * 8), "cudaMalloc(tile_end)");
    CK(ctx->tile_kagg.ensure(g.ntiles * 4), "cudaMalloc(tile_kagg)");
    CK(ctx->misc.ensure(256), "cudaMalloc(misc)");
    CK(cudaMemsetAsync(ctx->desc.p, 0, g.ntiles * 8, ctx->stream), "memset(desc)");
    CK(cudaMemsetAsync(ctx->pref.p, 0, g.ntiles * 8, ctx->stream), "memset(pref)");
    CK(cudaMemsetAsync(ctx->misc.p, 0, 256, ctx->stream), "memset(misc)");

    AecEncArgs a;
    memset(&a, 0, sizeof a);
    a.cfg = c;
    a.in = (const uint8_t *)d_in;
    a.nsamples = g.nsamples;
    a.nrsi = g.nrsi;
    a.last_nblk = g.last_nblk;
    a.RP = g.RP;
    a.ntiles = (ctx->tile_limit && ctx->tile_limit < g.ntiles) ? ctx->tile_limit : g.ntiles;
    a.ntiles_total = g.ntiles;
    a.aligned = (((uintptr_t)d_in & 15u) == 0) ? 1u : 0u;
    a.staging_words = aec_encode_staging_words(c);
    a.out_words = (uint32_t *)d_out;
    a.out_cap_words = out_cap / 4;
    a.out_cap_bytes = out_cap;
    a.seed_bits = carry->bits;
    a.seed_k = carry->k;
    a.seed_word = carry->word;
    a.desc = (uint64_t *)ctx->desc.p;
    a.pref = (uint64_t *)ctx->pref.p;
    a.ticket = (uint32_t *)ctx->misc.p;
    a.result = (uint64_t *)((uint8_t *)ctx->misc.p + 64);
    a.head_c = (uint32_t *)ctx->headc.p;
    a.tail_c = (uint32_t *)ctx->tailc.p;
    a.tile_end = (uint64_t *)ctx->tile_end.p;
    a.tile_kagg = (uint32_t *)ctx->tile_kagg.p;
    a.rsi_offsets = d_rsi_offsets;
    a.grp_index = d_grp_index;
    a.grp_G = aec_decode_group_blocks(c);
    const bool repair = a.ntiles < a.ntiles_total;
    CK(aec_encode_launch(a, ctx->num_sms, ctx->stream), "encode launch");
    ctx->launches += 2;
    if (ctx->want_summary && !repair) {
        CK(aec_encode_summary_launch(a, ctx->stream), "summary launch");
        ctx->launches += 1;
    }
    ctx->tile_limit = 0;
    if (!repair)
        CK(cudaMemcpyAsync(ctx->h_res, a.result, 48, cudaMemcpyDeviceToHost, ctx->stream), "memcpy(result)");
    return AEC_OK;
}

int aecb200_encode_finish(aecb200_ctx *ctx, aecb200_carry *end)
{
    if (!ctx || !ctx->enc_pending) return AEC_CONF_ERROR;
    CK(cudaStreamSynchronize(ctx->stream), "encode sync");
    ctx->enc_pending = false;
    if (end) { end->bits = ctx->h_res[0]; end->k = (uint32_t)ctx->h_res[1]; end->word = 0; }
    if (ctx->h_res[0] > ctx->enc_out_cap_bits) return AEC_STREAM_ERROR;
    return AEC_OK;
}

void aecb200_ctx_set_shard_mode(aecb200_ctx *ctx, int on) { if (ctx) ctx->want_summary = on != 0; }

int aecb200_encode_shard_info(aecb200_ctx *ctx, uint32_t *klo, uint32_t *khi, uint64_t *first_const_tile,
                              uint64_t *tail64)
{
    if (!ctx) return AEC_CONF_ERROR;
    if (tail64) *tail64 = ctx->h_res[5];
    if (klo) *klo = (uint32_t)ctx->h_res[2];
    if (khi) *khi = (uint32_t)ctx->h_res[3];
    if (first_const_tile) *first_const_tile = ctx->h_res[4];
    return AEC_OK;
}

void aecb200_ctx_set_tile_limit(aecb200_ctx *ctx, uint64_t ntiles) { if (ctx) ctx->tile_limit = ntiles; }

int aecb200_place_bits_device(aecb200_ctx *ctx, const void *d_src, uint64_t nbits,
                              void *d_dst, size_t dst_cap, uint64_t dst_bit, uint32_t head_or)
{
    if (!ctx) return AEC_CONF_ERROR;
    if ((((uintptr_t)d_src) & 3u) || (((uintptr_t)d_dst) & 3u)) return AEC_CONF_ERROR;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(aec_place_bits_launch((const uint32_t *)d_src, nbits, (uint32_t *)d_dst, dst_bit, dst_cap / 4, head_or, ctx->stream),
       "place launch");
    ctx->launches += 1;
    return AEC_OK;
}

int aecb200_decode_device(aecb200_ctx *ctx, const aecb200_params *p,
                          const void *d_in, size_t in_bytes,
                          const uint64_t *d_rsi_offsets, size_t nrsi,
                          void *d_out, size_t out_bytes)
{
    return aecb200_decode_device_indexed(ctx, p, d_in, in_bytes, d_rsi_offsets, nrsi, nullptr, d_out, out_bytes);
}

int aecb200_decode_device_indexed(aecb200_ctx *ctx, const aecb200_params *p,
                                  const void *d_in, size_t in_bytes,
                                  const uint64_t *d_rsi_offsets, size_t nrsi,
                                  const uint64_t *d_grp_index,
                                  void *d_out, size_t out_bytes)
{
    if (!ctx || !p) return AEC_CONF_ERROR;
    AecCfg c;
    int rc = make_cfg(ctx, p, 0, &c);
    if (rc != AEC_OK) return rc;
    if (((uintptr_t)d_in & 3u) != 0) {
        snprintf(ctx->err, sizeof ctx->err, "device input must be 4-byte aligned");
        return AEC_CONF_ERROR;
    }
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    uint64_t out_samples = out_bytes / c.B;
    uint64_t need_rsi = (out_samples + c.R - 1) / c.R;
    if (need_rsi > nrsi) need_rsi = nrsi;
    ctx->dec_pending = true;
    ctx->dec_out_samples = out_samples;
    ctx->dec_B = c.B;
    CK(ctx->misc.ensure(256), "cudaMalloc(misc)");
    uint64_t *res = (uint64_t *)((uint8_t *)ctx->misc.p + 128);
    /* delivered = min(out_samples, RSIs available * R) unless a lane reports less:
     * lanes that fall short atomicMax the complement of their position into
     * res[0] (zero = nobody fell short), flags go to res[1], res[2] = hand-over count */
    uint64_t avail = need_rsi * (uint64_t)c.R;
    ctx->dec_expect = out_samples < avail ? out_samples : avail;
    CK(cudaMemsetAsync(res, 0, 32, ctx->stream), "memset(result)");
    if (need_rsi) {
        AecDecArgs a;
        memset(&a, 0, sizeof a);
        a.cfg = c;
        a.in_words = (const uint32_t *)d_in;
        a.in_bytes = in_bytes;
        a.rsi_offsets = d_rsi_offsets;
        a.nrsi = need_rsi;
        a.out = (uint8_t *)d_out;
        a.out_samples = out_samples;
        a.out_aligned = (((uintptr_t)d_out & 3u) == 0) ? 1u : 0u;
        a.result = res;
        a.rsi_count = nullptr;
        a.grp_G = aec_decode_group_blocks(c);
        const bool fast = aec_decode_warp_warps(c) != 0 && need_rsi < 0xFFFFFFFFull && !ctx->careful_only;
        if (fast) {
            /* warp-per-RSI kernel from the group index; what it cannot finish goes to the careful kernel */
            CK(ctx->rsi_list.ensure(need_rsi * 4), "cudaMalloc(rsi_list)");
            a.rsi_list = (uint32_t *)ctx->rsi_list.p;
            a.rsi_list_count = (uint32_t *)(res + 2);
            if (!d_grp_index) {
                CK(ctx->grp.ensure(need_rsi * 32 * 8), "cudaMalloc(group index)");
                CK(aec_build_group_index_launch(a, (uint64_t *)ctx->grp.p, ctx->stream), "group index launch");
                ctx->launches += 1;
                d_grp_index = (const uint64_t *)ctx->grp.p;
            }
            a.grp_index = d_grp_index;
            CK(aec_decode_warp_launch(a, ctx->num_sms, ctx->stream), "decode (warp) launch");
            ctx->launches += 1;
        }
        CK(aec_decode_launch(a, ctx->num_sms, ctx->stream), "decode launch");
        ctx->launches += 1;
    }
    CK(cudaMemcpyAsync(&ctx->h_res[4], res, 24, cudaMemcpyDeviceToHost, ctx->stream), "memcpy(result)");
    return AEC_OK;
}

void aecb200_ctx_set_careful_decode(aecb200_ctx *ctx, int on) { if (ctx) ctx->careful_only = on != 0; }

/* RSIs the fast decode kernel handed to the careful kernel in the last finished decode */
uint64_t aecb200_ctx_last_handover(aecb200_ctx *ctx) { return ctx ? (ctx->h_res[6] & 0xFFFFFFFFull) : 0; }

int aecb200_decode_finish(aecb200_ctx *ctx, size_t *out_written)
{
    if (!ctx || !ctx->dec_pending) return AEC_CONF_ERROR;
    CK(cudaStreamSynchronize(ctx->stream), "decode sync");
    ctx->dec_pending = false;
    uint64_t got = ctx->dec_expect;
    if (ctx->h_res[4] != 0 && ~ctx->h_res[4] < got) got = ~ctx->h_res[4];
    if (out_written) *out_written = (size_t)(got * ctx->dec_B);
    if (ctx->h_res[5] & 1ull) return AEC_DATA_ERROR;
    return AEC_OK;
}

int aecb200_scan_offsets_device(aecb200_ctx *ctx, const aecb200_params *p,
                                const void *d_in, size_t in_bytes, uint64_t start_bit,
                                uint64_t *d_rsi_offsets, size_t max_rsi, size_t *found)
{
    if (!ctx || !p) return AEC_CONF_ERROR;
    AecCfg c;
    int rc = make_cfg(ctx, p, 0, &c);
    if (rc != AEC_OK) return rc;
    c.pad = (p->flags & AECF_PAD_RSI) ? 1u : 0u;      /* the decoder always honours it (decode.c:406-408) */
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(ctx->misc.ensure(256), "cudaMalloc(misc)");
    uint64_t *res = (uint64_t *)((uint8_t *)ctx->misc.p + 192);
    if (found) *found = 0;
    if (max_rsi == 0) return AEC_OK;
    CK(aec_scan_offsets_launch(c, (const uint32_t *)d_in, in_bytes, start_bit, d_rsi_offsets, max_rsi, res, ctx->stream),
       "scan launch");
    ctx->launches += 1;
    CK(cudaMemcpyAsync(&ctx->h_res[12], res, 24, cudaMemcpyDeviceToHost, ctx->stream), "memcpy(scan)");
    CK(cudaStreamSynchronize(ctx->stream), "scan sync");
    if (found) *found = (size_t)ctx->h_res[12];
    return (ctx->h_res[13] & 1ull) ? AEC_DATA_ERROR : AEC_OK;
}

/* ------------------------------------------------------------------------ */
/* host buffers                                                              */
/* ------------------------------------------------------------------------ */

static int encode_host_impl(aecb200_ctx *ctx, const aecb200_params *p,
                            const void *in, size_t in_bytes, int final,
                            void *out, size_t out_cap, size_t *out_len, size_t *in_consumed,
                            aecb200_carry *carry,
                            uint64_t *rsi_offsets, size_t offsets_cap, size_t *n_offsets)
{
    AecCfg c;
    int rc = make_cfg(ctx, p, 1, &c);
    if (out_len) *out_len = 0;
    if (in_consumed) *in_consumed = 0;
    if (n_offsets) *n_offsets = 0;
    if (rc != AEC_OK) return rc;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");

    const size_t rsi_bytes = (size_t)c.R * c.B;
    size_t use_bytes = final ? (in_bytes / c.B) * c.B : (in_bytes / rsi_bytes) * rsi_bytes;
    const uint64_t nrsi = (use_bytes / c.B + c.R - 1) / c.R;
    const uint32_t phase = (uint32_t)(carry->bits & 31u);

    uint64_t end_bits = phase;
    uint32_t end_k = carry->k;
    size_t bound = aecb200_encode_bound(p, use_bytes) + 8;
    if (use_bytes) {
        CK(ctx->in_stage.ensure(use_bytes + 16), "cudaMalloc(in)");
        CK(ctx->out_stage.ensure(bound), "cudaMalloc(out)");
        uint64_t *d_offs = nullptr;
        if (rsi_offsets && offsets_cap) {
            CK(ctx->offs.ensure(nrsi * 8), "cudaMalloc(offsets)");
            d_offs = (uint64_t *)ctx->offs.p;
        }
        CK(cudaMemcpyAsync(ctx->in_stage.p, in, use_bytes, cudaMemcpyHostToDevice, ctx->stream), "H2D");
        aecb200_carry seed = {phase, carry->k, carry->word};
        rc = aecb200_encode_device(ctx, p, ctx->in_stage.p, use_bytes, ctx->out_stage.p, ctx->out_stage.cap & ~(size_t)3,
                                   &seed, d_offs);
        if (rc != AEC_OK) return rc;
        aecb200_carry e;
        rc = aecb200_encode_finish(ctx, &e);
        if (rc != AEC_OK) return rc;
        end_bits = e.bits; end_k = e.k;
        if (d_offs) {
            size_t ncopy = nrsi < offsets_cap ? (size_t)nrsi : offsets_cap;
            CK(cudaMemcpyAsync(rsi_offsets, d_offs, ncopy * 8, cudaMemcpyDeviceToHost, ctx->stream), "D2H offsets");
            if (n_offsets) *n_offsets = (size_t)nrsi;
        }
    }
    /* bytes to hand out: complete bytes, plus the padded last one when final
     * (encode.c:686-695; an empty stream still yields one zero byte) */
    size_t nbytes;
    if (final) nbytes = (size_t)((end_bits + 7) / 8);
    else nbytes = (size_t)(end_bits / 8);
    if (final && end_bits == 0) nbytes = 1;
    size_t ncopy = nbytes < out_cap ? nbytes : out_cap;
    uint32_t tailword = 0;
    if (use_bytes) {
        if (ncopy) CK(cudaMemcpyAsync(out, ctx->out_stage.p, ncopy, cudaMemcpyDeviceToHost, ctx->stream), "D2H");
        if (!final && (end_bits & 7u)) {
            CK(cudaMemcpyAsync(&ctx->h_res[10], (uint8_t *)ctx->out_stage.p + (end_bits / 8), 1,
                               cudaMemcpyDeviceToHost, ctx->stream), "D2H tail");
        }
        CK(cudaStreamSynchronize(ctx->stream), "sync");
        if (!final && (end_bits & 7u)) tailword = (uint32_t)(*(uint8_t *)&ctx->h_res[10]) << 24;
    } else {
        /* no whole sample: the stream so far is just the carried partial byte */
        uint8_t b0 = (uint8_t)(carry->word >> 24);
        if (ncopy) ((uint8_t *)out)[0] = b0;
        if (!final) tailword = carry->word;
    }
    if (out_len) *out_len = ncopy;
    if (in_consumed) *in_consumed = use_bytes;
    carry->bits = final ? 0 : (end_bits & 7u);
    carry->k = end_k;
    carry->word = final ? 0 : tailword;
    return nbytes <= out_cap ? AEC_OK : AEC_STREAM_ERROR;
}

int aecb200_encode_host(aecb200_ctx *ctx, const aecb200_params *p,
                        const void *in, size_t in_bytes,
                        void *out, size_t out_cap, size_t *out_len, size_t *in_consumed,
                        uint64_t *rsi_offsets, size_t offsets_cap, size_t *n_offsets)
{
    if (!ctx || !p) return AEC_CONF_ERROR;
    aecb200_carry carry = {0, 0, 0};
    return encode_host_impl(ctx, p, in, in_bytes, 1, out, out_cap, out_len, in_consumed, &carry,
                            rsi_offsets, offsets_cap, n_offsets);
}

int aecb200_encode_host_piece(aecb200_ctx *ctx, const aecb200_params *p,
                              const void *in, size_t in_bytes, int final,
                              void *out, size_t out_cap, size_t *out_len, size_t *in_consumed,
                              aecb200_carry *carry,
                              uint64_t *rsi_offsets, size_t offsets_cap, size_t *n_offsets)
{
    if (!ctx || !p || !carry) return AEC_CONF_ERROR;
    return encode_host_impl(ctx, p, in, in_bytes, final, out, out_cap, out_len, in_consumed, carry,
                            rsi_offsets, offsets_cap, n_offsets);
}

int aecb200_decode_host_resume(aecb200_ctx *ctx, const aecb200_params *p,
                               const void *in, size_t in_bytes,
                               const uint64_t *rsi_offsets, size_t n_offsets,
                               uint64_t start_bit, size_t skip_samples,
                               void *out, size_t out_cap, size_t *out_len,
                               uint64_t *resume_bit, size_t *resume_delivered)
{
    if (!ctx || !p) return AEC_CONF_ERROR;
    AecCfg c;
    int rc = make_cfg(ctx, p, 0, &c);
    if (out_len) *out_len = 0;
    if (resume_bit) *resume_bit = start_bit;
    if (resume_delivered) *resume_delivered = skip_samples;
    if (rc != AEC_OK) return rc;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const uint64_t want_new = out_cap / c.B;                       /* samples the caller can take */
    const uint64_t out_samples = skip_samples + want_new;          /* counted from the RSI at start_bit */
    uint64_t need_rsi = (out_samples + c.R - 1) / c.R;
    if (want_new == 0 || in_bytes * 8ull <= start_bit)
        return AEC_OK;
    /* stage the stream from the 32-bit word that holds start_bit */
    const size_t base_byte = (size_t)((start_bit >> 5) << 2);
    const uint64_t base_bit = (uint64_t)base_byte * 8ull;
    const size_t nbytes = in_bytes - base_byte;
    const size_t in_pad = (nbytes + 3) & ~(size_t)3;
    CK(ctx->in_stage.ensure(in_pad + 16), "cudaMalloc(in)");
    CK(ctx->out_stage.ensure((size_t)(out_samples * c.B) + 16), "cudaMalloc(out)");
    CK(ctx->offs.ensure((need_rsi + 1) * 8), "cudaMalloc(offsets)");
    CK(cudaMemsetAsync((uint8_t *)ctx->in_stage.p + (in_pad - 4), 0, 4, ctx->stream), "memset(in tail)");
    CK(cudaMemcpyAsync(ctx->in_stage.p, (const uint8_t *)in + base_byte, nbytes, cudaMemcpyHostToDevice, ctx->stream), "H2D");
    size_t nrsi = 0;
    uint64_t scan_end = 0;
    uint64_t *h_offs = nullptr;
    if (rsi_offsets) {
        /* caller's index is relative to bit 0 of `in`: entries from the RSI that starts at start_bit */
        size_t first = 0;
        while (first < n_offsets && rsi_offsets[first] < start_bit) first++;
        nrsi = n_offsets - first < need_rsi ? n_offsets - first : (size_t)need_rsi;
        h_offs = (uint64_t *)malloc((nrsi + 1) * sizeof(uint64_t));
        if (!h_offs) return AEC_MEM_ERROR;
        for (size_t i = 0; i < nrsi; i++) h_offs[i] = rsi_offsets[first + i] - base_bit;
        scan_end = (first + nrsi < n_offsets) ? rsi_offsets[first + nrsi] - base_bit : (uint64_t)nbytes * 8ull;
        if (nrsi) {
            cudaError_t e = cudaMemcpyAsync(ctx->offs.p, h_offs, nrsi * 8, cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { free(h_offs); return fail_cuda(ctx, e, "H2D offsets"); }
        }
    } else {
        rc = aecb200_scan_offsets_device(ctx, p, ctx->in_stage.p, nbytes, start_bit - base_bit,
                                         (uint64_t *)ctx->offs.p, (size_t)need_rsi, &nrsi);
        if (rc != AEC_OK && rc != AEC_DATA_ERROR) return rc;
        scan_end = ctx->h_res[14];
        h_offs = (uint64_t *)malloc((nrsi + 1) * sizeof(uint64_t));
        if (!h_offs) return AEC_MEM_ERROR;
        if (nrsi) {
            cudaError_t e = cudaMemcpyAsync(h_offs, ctx->offs.p, nrsi * 8, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { free(h_offs); return fail_cuda(ctx, e, "D2H offsets"); }
        }
    }
    size_t written = 0;
    rc = aecb200_decode_device(ctx, p, ctx->in_stage.p, nbytes, (const uint64_t *)ctx->offs.p, nrsi,
                               ctx->out_stage.p, (size_t)(out_samples * c.B));
    if (rc == AEC_OK) rc = aecb200_decode_finish(ctx, &written);
    if (rc != AEC_OK) { free(h_offs); return rc; }
    uint64_t W = written / c.B;                                    /* samples decoded from start_bit */
    size_t newbytes = W > skip_samples ? (size_t)((W - skip_samples) * c.B) : 0;
    if (newbytes) {
        cudaError_t e = cudaMemcpyAsync(out, (uint8_t *)ctx->out_stage.p + skip_samples * c.B, newbytes,
                                        cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { free(h_offs); return fail_cuda(ctx, e, "D2H"); }
    }
    if (out_len) *out_len = newbytes;
    if (W < skip_samples) W = skip_samples;
    uint64_t full = W / c.R, rem = W % c.R;
    uint64_t rb;
    if (full < nrsi) rb = h_offs[full] + base_bit;
    else rb = scan_end + base_bit;
    if (resume_bit) *resume_bit = rb;
    if (resume_delivered) *resume_delivered = (size_t)rem;
    free(h_offs);
    return AEC_OK;
}

int aecb200_decode_host(aecb200_ctx *ctx, const aecb200_params *p,
                        const void *in, size_t in_bytes,
                        const uint64_t *rsi_offsets, size_t n_offsets,
                        void *out, size_t out_cap, size_t *out_len)
{
    if (!ctx || !p) return AEC_CONF_ERROR;
    AecCfg c;
    int rc = make_cfg(ctx, p, 0, &c);
    if (out_len) *out_len = 0;
    if (rc != AEC_OK) return rc;
    size_t written = 0;
    rc = aecb200_decode_host_resume(ctx, p, in, in_bytes, rsi_offsets, n_offsets, 0, 0,
                                    out, out_cap, &written, nullptr, nullptr);
    if (rc != AEC_OK) return rc;
    if (out_len) *out_len = written;
    size_t left = out_cap - written;
    if (left > 0 && left < c.B) return AEC_MEM_ERROR;               /* decode.c:821-823 */
    return AEC_OK;
}

} /* extern "C" */
