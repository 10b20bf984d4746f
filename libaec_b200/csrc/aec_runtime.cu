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
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

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
    /* same, and a new allocation starts out as zeros (control words the kernels themselves reset) */
    cudaError_t ensure_zeroed(size_t n, cudaStream_t st)
    {
        if (n <= cap) return cudaSuccess;
        cudaError_t e = ensure(n);
        if (e == cudaSuccess) e = cudaMemsetAsync(p, 0, cap, st);
        return e;
    }
    /* same, keeping the first `keep` bytes */
    cudaError_t ensure_preserve(size_t n, size_t keep, cudaStream_t st)
    {
        if (n <= cap) return cudaSuccess;
        void *np = nullptr;
        size_t want = n + n / 2 + 256;
        cudaError_t e = cudaMalloc(&np, want);
        if (e != cudaSuccess) return e;
        if (p && keep) e = cudaMemcpyAsync(np, p, keep, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (p) cudaFree(p);
        p = np; cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

} // namespace

namespace {

/* A few helper threads that copy between pageable caller memory and pinned bounce slots.  cudaMemcpyAsync
 * on pageable memory goes through one staging thread of the driver (8-11 GB/s up, ~20 GB/s down on the
 * round's boxes, profiles/r2_summary.md); four threads move 44 GB/s, which keeps PCIe busy for callers
 * that hand over malloc'd buffers -- every caller of libaec.h.  One copy at a time per process; a
 * second caller that finds the pool busy copies on its own thread. */
class HostCopyPool {
public:
    static HostCopyPool &get() { static HostCopyPool p; return p; }
    void copy(void *dst, const void *src, size_t n)
    {
        if (n < ((size_t)1 << 20) || !busy.try_lock()) { memcpy(dst, src, n); return; }
        {
            std::lock_guard<std::mutex> lk(mu);
            if (th.empty())
                for (int i = 0; i < HELPERS; i++) th.emplace_back([this, i] { worker(i); });
            jdst = (uint8_t *)dst; jsrc = (const uint8_t *)src; jn = n;
            pending = HELPERS;
            gen++;
        }
        cv.notify_all();
        part(HELPERS);                                   /* the caller takes the last part */
        {
            std::unique_lock<std::mutex> lk(mu);
            done_cv.wait(lk, [this] { return pending == 0; });
        }
        busy.unlock();
    }
    ~HostCopyPool()
    {
        { std::lock_guard<std::mutex> lk(mu); stop = true; gen++; }
        cv.notify_all();
        for (std::thread &t : th) t.join();
    }
private:
    static constexpr int HELPERS = 3;
    void part(int i)
    {
        const size_t per = ((jn / (HELPERS + 1)) + 63) & ~(size_t)63;
        const size_t a = per * (size_t)i < jn ? per * (size_t)i : jn;
        const size_t b = (i == HELPERS) ? jn : (a + per < jn ? a + per : jn);
        if (b > a) memcpy(jdst + a, jsrc + a, b - a);
    }
    void worker(int i)
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return gen != seen; });
                seen = gen;
                if (stop) return;
            }
            part(i);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--pending == 0) done_cv.notify_one();
            }
        }
    }
    std::vector<std::thread> th;
    std::mutex mu, busy;
    std::condition_variable cv, done_cv;
    uint8_t *jdst = nullptr; const uint8_t *jsrc = nullptr; size_t jn = 0;
    int pending = 0; uint64_t gen = 0; bool stop = false;
};

constexpr size_t BOUNCE_SLOT = (size_t)8 << 20;

struct Bounce {                 /* two pinned slots per direction, each with the event of its last device copy */
    uint8_t *slot[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool used[4] = {false, false, false, false};
    unsigned turn[2] = {0, 0};
    cudaError_t ensure()
    {
        for (int i = 0; i < 4; i++) {
            if (!slot[i]) { cudaError_t e = cudaMallocHost((void **)&slot[i], BOUNCE_SLOT); if (e != cudaSuccess) return e; }
            if (!ev[i]) { cudaError_t e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming); if (e != cudaSuccess) return e; }
        }
        return cudaSuccess;
    }
    void release()
    {
        for (int i = 0; i < 4; i++) {
            if (ev[i]) { cudaEventSynchronize(ev[i]); cudaEventDestroy(ev[i]); ev[i] = nullptr; }
            if (slot[i]) { cudaFreeHost(slot[i]); slot[i] = nullptr; }
        }
    }
};

bool host_pointer_is_pageable(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

} // namespace

struct aecb200_ctx {
    Bounce bounce;
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int honour_pad = 0;
    uint64_t launches = 0;
    char err[256] = {0};

    DevBuf grp, rsi_list, desc, pref, headc, tailc, tile_end, tile_kagg, misc, in_stage, out_stage, offs, rsi_count, plan;
    void *shard_out = nullptr;           /* device: (bits, klo, khi, tail64) of every shard-mode encode */
    DevBuf raw_stage, out2_stage;        /* SZIP shim: the caller's bytes before / after the byte shuffles */
    /* streaming decode: the compressed bytes received so far stay in HBM (acc_stage holds stream bytes
     * [acc_start, acc_start + acc_len)); a call only uploads what is new */
    DevBuf acc_stage;
    uint64_t acc_start = 0;
    size_t acc_len = 0;
    bool acc_valid = false;
    aecb200_ctx *aux = nullptr;          /* second context (stream + workspace): decodes RSIs already discovered while the scan goes on */
    long long acc_stream_byte0 = -1;     /* >= 0: the next aecb200_decode_host_resume call accumulates; in[0] is this stream byte */
    uint64_t acc_uploaded = 0;           /* bytes sent to the device by accumulating calls (diagnostics) */
    uint64_t staged_uploads = 0;         /* bytes of encoder input accumulated on the device (diagnostics) */
    bool in_stage_ready = false;         /* the next host encode finds its input in in_stage already */
    bool no_bounce = false;              /* pageable host buffers go straight to cudaMemcpyAsync (AECB200_NO_BOUNCE, tests) */
    uint64_t tile_limit = 0;             /* next encode codes only this many leading tiles (k repair) */
    bool want_summary = false;
    bool careful_only = false;           /* decode with the lane-per-RSI kernel only (tests) */
    uint64_t *h_res = nullptr;           /* pinned: [0..3] encode result, [4..7] decode result */
    /* host-pointer calls on large buffers run as a pipeline of pieces: uploads on s_in, kernels on
     * `stream`, downloads on s_out (PCIe carries both directions at once) */
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaStream_t s_idx[4] = {nullptr, nullptr, nullptr, nullptr};   /* group-index builds of the decode pipeline */
    cudaStream_t s_walk = nullptr;                                   /* RSI boundary walk (aec_skim.cu) */
    cudaEvent_t ev_skim[4] = {nullptr, nullptr, nullptr, nullptr};   /* tables ready [0,1], walk done [2,3] per table set */
    cudaEvent_t ev_prog[4] = {nullptr, nullptr, nullptr, nullptr};   /* walk of window i done, ring of four (progress hook) */
    std::vector<cudaEvent_t> ev;
    size_t pipe_piece = (size_t)16 << 20; /* bytes of raw samples per piece; 0 = never pipeline */

    /* RSI boundary discovery for streams without an index (aec_skim.cu) */
    DevBuf skim_tab;
    int scan_mode = 0;                   /* 0 auto, 1 always the one-thread scan, 2 always the parallel tables */
    uint64_t scan_window_bits = 0;        /* 0: by stream size (scan_offsets_impl) */
    uint64_t scan_end = 0, scan_fast = 0;
    bool scan_grp = false;               /* the last scan also wrote the group index */
    uint64_t up_first_bits = 0;          /* un-indexed host decode: stream bits that were uploaded on `stream`; the rest is on s_in */
    int scan_sparse = 1;                 /* RSI lengths at marked chain ends only (AECB200_SCAN_SPARSE=0: every position, as before) */
    int scan_skip8 = -1;                 /* eight-RSI jumps of the walk: -1 by RSI density, 0 never, 1 always (AECB200_SCAN_SKIP8) */
    std::vector<uint64_t> found_offs;    /* RSI offsets the last host decode discovered itself (bits from in[0]) */

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

/* Work runs on the context's device; the caller's current device is put back when the call returns
 * (the reference library has no such side effect, and a multi-GPU host thread relies on it). */
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t enter(int dev)
    {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) return e;
        if (prev != dev) { e = cudaSetDevice(dev); switched = (e == cudaSuccess); }
        return e;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};
#define ENTER_DEVICE() DeviceGuard dg_; CK(dg_.enter(ctx->device), "cudaSetDevice")

/* Host -> device.  Pinned (or small) sources: one asynchronous copy.  Pageable sources: 8 MiB pieces through
 * two pinned slots, filled by the host-copy pool while the previous slot is on its way; returns when the
 * last piece is staged (the caller's buffer is no longer needed), its device copy still in flight on `st`. */
cudaError_t copy_h2d(aecb200_ctx *ctx, void *d_dst, const void *h_src, size_t n, cudaStream_t st)
{
    if (n < ((size_t)2 << 20) || ctx->no_bounce || !host_pointer_is_pageable(h_src))
        return cudaMemcpyAsync(d_dst, h_src, n, cudaMemcpyHostToDevice, st);
    Bounce &b = ctx->bounce;
    cudaError_t e = b.ensure();
    if (e != cudaSuccess) return e;
    for (size_t off = 0; off < n; off += BOUNCE_SLOT) {
        const size_t len = n - off < BOUNCE_SLOT ? n - off : BOUNCE_SLOT;
        const int k = (int)(b.turn[0]++ & 1u);
        if (b.used[k] && (e = cudaEventSynchronize(b.ev[k])) != cudaSuccess) return e;
        HostCopyPool::get().copy(b.slot[k], (const uint8_t *)h_src + off, len);
        if ((e = cudaMemcpyAsync((uint8_t *)d_dst + off, b.slot[k], len, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(b.ev[k], st)) != cudaSuccess) return e;
        b.used[k] = true;
    }
    return cudaSuccess;
}

/* Device -> host.  Pinned (or small) destinations: one asynchronous copy (complete when `st` is synchronised).
 * Pageable destinations: through the two download slots; the data is in place when the call returns. */
cudaError_t copy_d2h(aecb200_ctx *ctx, void *h_dst, const void *d_src, size_t n, cudaStream_t st)
{
    if (n < ((size_t)2 << 20) || ctx->no_bounce || !host_pointer_is_pageable(h_dst))
        return cudaMemcpyAsync(h_dst, d_src, n, cudaMemcpyDeviceToHost, st);
    Bounce &b = ctx->bounce;
    cudaError_t e = b.ensure();
    if (e != cudaSuccess) return e;
    size_t prev_off = 0, prev_len = 0;
    int prev_k = -1;
    for (size_t off = 0; off < n || prev_k >= 0; off += BOUNCE_SLOT) {
        int k = -1;
        size_t len = 0;
        if (off < n) {
            len = n - off < BOUNCE_SLOT ? n - off : BOUNCE_SLOT;
            k = 2 + (int)(b.turn[1]++ & 1u);
            if (b.used[k] && (e = cudaEventSynchronize(b.ev[k])) != cudaSuccess) return e;   /* (always complete: see below) */
            if ((e = cudaMemcpyAsync(b.slot[k], (const uint8_t *)d_src + off, len, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(b.ev[k], st)) != cudaSuccess) return e;
            b.used[k] = true;
        }
        if (prev_k >= 0) {                               /* the piece before: wait for it, hand it to the caller */
            if ((e = cudaEventSynchronize(b.ev[prev_k])) != cudaSuccess) return e;
            HostCopyPool::get().copy((uint8_t *)h_dst + prev_off, b.slot[prev_k], prev_len);
        }
        prev_k = k; prev_off = off; prev_len = len;
    }
    return cudaSuccess;
}

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

/* streams and events of the host-call pipeline (created on first use) */
int pipe_prepare(aecb200_ctx *ctx, size_t nevents)
{
    if (!ctx->s_in) CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking), "cudaStreamCreate(in)");
    if (!ctx->s_out) CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking), "cudaStreamCreate(out)");
    for (cudaStream_t &st : ctx->s_idx)
        if (!st) CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate(index)");
    while (ctx->ev.size() < nevents) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
        ctx->ev.push_back(e);
    }
    return AEC_OK;
}

/* a failed piece must not leave copies in flight on the side streams */
int pipe_abort(aecb200_ctx *ctx, int rc)
{
    cudaStreamSynchronize(ctx->s_in);
    for (cudaStream_t st : ctx->s_idx) if (st) cudaStreamSynchronize(st);
    if (ctx->s_walk) cudaStreamSynchronize(ctx->s_walk);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->s_out);
    return rc;
}

/* Leaves no copy or kernel of a pipeline in flight when a host call returns early (CUDA failure, bad
 * piece): the caller's buffers must not be touched after the call has returned. */
struct PipeGuard {
    aecb200_ctx *ctx;
    bool armed;
    explicit PipeGuard(aecb200_ctx *c) : ctx(c), armed(true) {}
    ~PipeGuard() { if (armed) pipe_abort(ctx, 0); }
};

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
    DeviceGuard dg;
    e = dg.enter(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) { ctx->own_stream = true; e = cudaMallocHost(&ctx->h_res, 32 * sizeof(uint64_t)); }
    if (e != cudaSuccess) {
        fprintf(stderr, "aecb200: no usable CUDA device (%s); this library has no CPU fallback\n",
                cudaGetErrorString(e));
        if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
        return AECB200_CUDA_ERROR;
    }
    /* test hooks: force one of the RSI boundary discovery paths / a small table window */
    if (getenv("AECB200_NO_BOUNCE")) ctx->no_bounce = true;
    if (const char *k = getenv("AECB200_SCAN_SKIP8")) ctx->scan_skip8 = atoi(k);
    if (const char *k = getenv("AECB200_SCAN_SPARSE")) ctx->scan_sparse = atoi(k) != 0;
    if (const char *m = getenv("AECB200_SCAN_MODE")) ctx->scan_mode = atoi(m);
    if (const char *w = getenv("AECB200_SCAN_WINDOW_BITS")) { long long v = atoll(w); if (v > 0) ctx->scan_window_bits = (uint64_t)v; }
    *out = ctx;
    return AEC_OK;
}

void aecb200_ctx_destroy(aecb200_ctx *ctx)
{
    if (!ctx) return;
    DeviceGuard dg;
    dg.enter(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    ctx->tile_kagg.release(); ctx->pref.release(); ctx->grp.release(); ctx->rsi_list.release();
    ctx->desc.release(); ctx->headc.release(); ctx->tailc.release(); ctx->tile_end.release();
    ctx->misc.release(); ctx->in_stage.release(); ctx->out_stage.release(); ctx->offs.release();
    ctx->rsi_count.release(); ctx->skim_tab.release(); ctx->plan.release();
    ctx->raw_stage.release(); ctx->out2_stage.release(); ctx->acc_stage.release();
    if (ctx->aux) { aecb200_ctx_destroy(ctx->aux); ctx->aux = nullptr; }
    ctx->bounce.release();
    if (ctx->h_res) cudaFreeHost(ctx->h_res);
    for (cudaEvent_t e : ctx->ev) cudaEventDestroy(e);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    for (cudaStream_t st : ctx->s_idx) if (st) cudaStreamDestroy(st);
    if (ctx->s_walk) cudaStreamDestroy(ctx->s_walk);
    for (cudaEvent_t e : ctx->ev_skim) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_prog) if (e) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int aecb200_current_device(void)
{
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) return -1;
    return d;
}

int aecb200_ctx_device(aecb200_ctx *ctx) { return ctx ? ctx->device : -1; }

/* make `device` the calling thread's current device (worker threads of the batch calls) */
int aecb200_set_device(int device) { return cudaSetDevice(device) == cudaSuccess ? AEC_OK : AECB200_CUDA_ERROR; }

int aecb200_ctx_set_stream(aecb200_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return AEC_CONF_ERROR;
    DeviceGuard dg;
    dg.enter(ctx->device);
    if (ctx->own_stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    ctx->stream = (cudaStream_t)cuda_stream;
    return AEC_OK;
}

const char *aecb200_last_error(aecb200_ctx *ctx) { return ctx ? ctx->err : "no context"; }
void aecb200_ctx_set_encode_padding(aecb200_ctx *ctx, int on) { if (ctx) ctx->honour_pad = on ? 1 : 0; }
uint64_t aecb200_ctx_launches(aecb200_ctx *ctx) { return ctx ? ctx->launches : 0; }
void aecb200_ctx_set_pipeline_piece(aecb200_ctx *ctx, size_t raw_bytes) { if (ctx) ctx->pipe_piece = raw_bytes; }

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

static int encode_device_impl(aecb200_ctx *ctx, const aecb200_params *p, const void *d_in, size_t in_bytes,
                              void *d_out, size_t out_cap, const aecb200_carry *carry, uint64_t *d_rsi_offsets,
                              uint64_t *d_grp_index, bool planned_repair);

int aecb200_encode_device_indexed(aecb200_ctx *ctx, const aecb200_params *p,
                                  const void *d_in, size_t in_bytes,
                                  void *d_out, size_t out_cap,
                                  const aecb200_carry *carry, uint64_t *d_rsi_offsets,
                                  uint64_t *d_grp_index)
{
    return encode_device_impl(ctx, p, d_in, in_bytes, d_out, out_cap, carry, d_rsi_offsets, d_grp_index, false);
}

/* Shard protocol without the host (aec_b200.h): the k repair of a shard whose incoming k and number of
 * tiles to code again were worked out on the device by aecb200_shard_plan_device. */
int aecb200_encode_repair_device(aecb200_ctx *ctx, const aecb200_params *p, const void *d_in, size_t in_bytes,
                                 void *d_out, size_t out_cap)
{
    return encode_device_impl(ctx, p, d_in, in_bytes, d_out, out_cap, nullptr, nullptr, nullptr, true);
}

static int encode_device_impl(aecb200_ctx *ctx, const aecb200_params *p, const void *d_in, size_t in_bytes,
                              void *d_out, size_t out_cap, const aecb200_carry *carry, uint64_t *d_rsi_offsets,
                              uint64_t *d_grp_index, bool planned_repair)
{
    if (!ctx || !p) return AEC_CONF_ERROR;
    AecCfg c;
    int rc = make_cfg(ctx, p, 1, &c);
    if (rc != AEC_OK) return rc;
    if (((uintptr_t)d_out & 3u) != 0) {
        snprintf(ctx->err, sizeof ctx->err, "device output must be 4-byte aligned");
        return AEC_CONF_ERROR;
    }
    ENTER_DEVICE();
    EncGeom g = enc_geometry(c, in_bytes);
    aecb200_carry zero = {0, 0, 0};
    if (!carry) carry = &zero;

    if (planned_repair) {
        if (g.nsamples == 0) return AEC_OK;
        CK(ctx->plan.ensure_zeroed(PLAN_WORDS * 8, ctx->stream), "cudaMalloc(plan)");
    } else {
        ctx->enc_out_cap_bits = (uint64_t)out_cap * 8ull;
        ctx->enc_pending = true;
        if (g.nsamples == 0) {
            /* nothing to code: the result is the carry itself */
            ctx->h_res[0] = carry->bits; ctx->h_res[1] = carry->k;
            if (ctx->want_summary && ctx->shard_out) {
                /* an empty shard still has to answer the all_gather: zero bits, identity clamp pair */
                const uint64_t ident[4] = {0, 0, c.kmax, 0};
                memcpy(&ctx->h_res[8], ident, sizeof ident);
                CK(cudaMemcpyAsync(ctx->shard_out, &ctx->h_res[8], 32, cudaMemcpyHostToDevice, ctx->stream), "memcpy(shard out)");
            }
            return AEC_OK;
        }
    }

    if (g.ntiles >= 0xFFFFFFF0ull) {
        snprintf(ctx->err, sizeof ctx->err, "input too large for one launch (%llu tiles)", (unsigned long long)g.ntiles);
        return AEC_CONF_ERROR;
    }
    /* tile descriptors, prefixes and the ticket are zero between launches: the fix-up kernel, which
     * runs last, clears what its launch used (no memsets in front of every launch) */
    CK(ctx->desc.ensure_zeroed(g.ntiles * 8, ctx->stream), "cudaMalloc(desc)");
    CK(ctx->pref.ensure_zeroed(g.ntiles * 8, ctx->stream), "cudaMalloc(pref)");
    CK(ctx->headc.ensure(g.ntiles * 4), "cudaMalloc(head)");
    CK(ctx->tailc.ensure(g.ntiles * 4), "cudaMalloc(tail)");
    CK(ctx->tile_end.ensure(g.ntiles * 8), "cudaMalloc(tile_end)");
    CK(ctx->tile_kagg.ensure(g.ntiles * 4), "cudaMalloc(tile_kagg)");
    CK(ctx->misc.ensure_zeroed(256, ctx->stream), "cudaMalloc(misc)");

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
    a.shard_out = (ctx->want_summary && !planned_repair) ? (uint64_t *)ctx->shard_out : nullptr;
    a.dyn = planned_repair ? (const uint64_t *)ctx->plan.p : nullptr;
    const bool repair = a.ntiles < a.ntiles_total || planned_repair;
    if (planned_repair) {
        /* How many leading tiles depend on the incoming k is known on the device only.  Nearly always
         * it is one or two.  This launch covers up to AECB200_REPAIR_TILES tiles and runs when the plan
         * asks for that many or fewer (it returns at once when the incoming k is 0; tiles coded again with
         * the right k come out as they were).  The rare shard that needs more is repaired by the caller,
         * who sees the number in the plan (plan word 1 > AECB200_REPAIR_TILES): aecb200_ctx_set_tile_limit
         * + aecb200_encode_device with the true k, as in the host-driven protocol. */
        const uint64_t few = AECB200_REPAIR_TILES;
        a.ntiles = g.ntiles < few ? g.ntiles : few;
        a.dyn_lo = 0; a.dyn_hi = few;
        CK(aec_encode_launch(a, ctx->num_sms, ctx->stream), "repair launch");
        ctx->launches += 2;
        return AEC_OK;
    }
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
    ENTER_DEVICE();                 /* a caller-supplied stream may be the legacy default stream, which means "of the current device" */
    CK(cudaStreamSynchronize(ctx->stream), "encode sync");
    ctx->enc_pending = false;
    if (end) { end->bits = ctx->h_res[0]; end->k = (uint32_t)ctx->h_res[1]; end->word = 0; }
    if (ctx->h_res[0] > ctx->enc_out_cap_bits) return AEC_STREAM_ERROR;
    return AEC_OK;
}

void aecb200_ctx_set_shard_mode(aecb200_ctx *ctx, int on) { if (ctx) ctx->want_summary = on != 0; }
void aecb200_ctx_set_shard_out(aecb200_ctx *ctx, void *d_info) { if (ctx) ctx->shard_out = d_info; }

int aecb200_shard_plan_device(aecb200_ctx *ctx, const void *d_all, int world, int rank, void *d_plan_out)
{
    if (!ctx || !d_all || world < 1 || rank < 0 || rank >= world) return AEC_CONF_ERROR;
    ENTER_DEVICE();
    CK(ctx->plan.ensure_zeroed(PLAN_WORDS * 8, ctx->stream), "cudaMalloc(plan)");
    CK(ctx->misc.ensure_zeroed(256, ctx->stream), "cudaMalloc(misc)");
    const uint64_t *result = (const uint64_t *)((uint8_t *)ctx->misc.p + 64);
    CK(aec_shard_plan_launch((const uint64_t *)d_all, (uint32_t)world, (uint32_t)rank, result, (uint64_t *)ctx->plan.p,
                             (uint64_t *)d_plan_out, ctx->stream), "plan launch");
    ctx->launches += 1;
    return AEC_OK;
}

int aecb200_place_bits_planned(aecb200_ctx *ctx, const void *d_src, void *d_dst, size_t dst_cap, int global, int last_rank)
{
    if (!ctx || !ctx->plan.p) return AEC_CONF_ERROR;
    if ((((uintptr_t)d_src) & 3u) || (((uintptr_t)d_dst) & 3u)) return AEC_CONF_ERROR;
    ENTER_DEVICE();
    CK(aec_place_bits_planned_launch((const uint32_t *)d_src, (const uint64_t *)ctx->plan.p, (uint32_t *)d_dst, dst_cap / 4,
                                     global ? 1u : 0u, last_rank ? 1u : 0u, ctx->num_sms, ctx->stream), "place launch");
    ctx->launches += 1;
    return AEC_OK;
}

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
    ENTER_DEVICE();
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
    ENTER_DEVICE();
    uint64_t out_samples = out_bytes / c.B;
    uint64_t need_rsi = (out_samples + c.R - 1) / c.R;
    if (need_rsi > nrsi) need_rsi = nrsi;
    ctx->dec_pending = true;
    ctx->dec_out_samples = out_samples;
    ctx->dec_B = c.B;
    CK(ctx->misc.ensure_zeroed(256, ctx->stream), "cudaMalloc(misc)");
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
    ENTER_DEVICE();
    CK(cudaStreamSynchronize(ctx->stream), "decode sync");
    ctx->dec_pending = false;
    uint64_t got = ctx->dec_expect;
    if (ctx->h_res[4] != 0 && ~ctx->h_res[4] < got) got = ~ctx->h_res[4];
    if (out_written) *out_written = (size_t)(got * ctx->dec_B);
    if (ctx->h_res[5] & 1ull) return AEC_DATA_ERROR;
    return AEC_OK;
}

/* Hook of the un-indexed host decode: called between windows with the number of RSIs whose offsets are
 * final, so that they can be decoded and sent back while the discovery goes on. */
struct ScanProgress {
    virtual int consume(uint64_t complete_rsis, cudaEvent_t offsets_ready) = 0;
    virtual ~ScanProgress() {}
};

static int scan_offsets_impl(aecb200_ctx *ctx, const aecb200_params *p,
                             const void *d_in, size_t in_bytes, uint64_t start_bit,
                             uint64_t *d_rsi_offsets, size_t max_rsi, size_t *found, ScanProgress *prog,
                             uint64_t *d_grp = nullptr);

/* Window size of the boundary discovery: a window costs about 0.15 ms whatever its size (launches, every
 * kernel's tail), and the walk of window i overlaps the tables of window i+1 only when there are several: an
 * eighth of the stream, between 2^24 and 2^26 bits (measured on the README workload: 2^25 19.4 ms, 2^26 17.9 ms;
 * config 2, a 190 Mbit stream: 7.7 and 8.1 ms), unless the caller set one. */
static uint64_t scan_window_bits_for(const aecb200_ctx *ctx, uint64_t span_bits)
{
    uint64_t nh = ctx->scan_window_bits;
    if (nh == 0) {
        nh = span_bits / 8ull;
        if (nh < (1ull << 24)) nh = 1ull << 24;
        if (nh > (1ull << 26)) nh = 1ull << 26;
    }
    nh = (nh + 127ull) & ~127ull;
    return nh < 1024 ? 1024 : nh;
}

int aecb200_scan_offsets_device(aecb200_ctx *ctx, const aecb200_params *p,
                                const void *d_in, size_t in_bytes, uint64_t start_bit,
                                uint64_t *d_rsi_offsets, size_t max_rsi, size_t *found)
{
    return scan_offsets_impl(ctx, p, d_in, in_bytes, start_bit, d_rsi_offsets, max_rsi, found, nullptr);
}

/* d_grp (optional, max_rsi * 32 entries): the group index of the RSIs found is written from the chain tables
 * as well (ctx->scan_grp tells whether it was: the one-thread scan of short streams does not); RSIs that
 * had to be skimmed serially are marked SK_GRP_MISSING for aec_build_group_index_launch(only_missing). */
static int scan_offsets_impl(aecb200_ctx *ctx, const aecb200_params *p,
                             const void *d_in, size_t in_bytes, uint64_t start_bit,
                             uint64_t *d_rsi_offsets, size_t max_rsi, size_t *found, ScanProgress *prog,
                             uint64_t *d_grp)
{
    if (!ctx || !p) return AEC_CONF_ERROR;
    ctx->scan_grp = false;
    AecCfg c;
    int rc = make_cfg(ctx, p, 0, &c);
    if (rc != AEC_OK) return rc;
    c.pad = (p->flags & AECF_PAD_RSI) ? 1u : 0u;      /* the decoder always honours it (decode.c:406-408) */
    ENTER_DEVICE();
    CK(ctx->misc.ensure_zeroed(256, ctx->stream), "cudaMalloc(misc)");
    uint64_t *state = (uint64_t *)((uint8_t *)ctx->misc.p + 192);
    if (found) *found = 0;
    ctx->scan_end = start_bit;
    ctx->scan_fast = 0;
    if (max_rsi == 0) return AEC_OK;
    const uint64_t nbits = (uint64_t)in_bytes * 8ull;
    /* walk state: next RSI bit, RSIs found, flags, RSIs taken from the tables */
    uint64_t *h_state = &ctx->h_res[8];
    h_state[0] = start_bit;
    for (int i = 1; i < 8; i++) h_state[i] = 0;         /* [4..7]: first RSI of the window, dense request, list counters */
    CK(cudaMemcpyAsync(state, h_state, 64, cudaMemcpyHostToDevice, ctx->stream), "memcpy(scan state)");
    uint64_t up_first_bits = ctx->up_first_bits;        /* stream bits uploaded on the context's stream; the rest follows on s_in (ev[0]) */
    const uint64_t base = start_bit & ~127ull;          /* windows start 16-byte aligned (bulk copies of the tiles) */
    const bool parallel = ctx->scan_mode == 2 || (ctx->scan_mode == 0 && in_bytes >= 2048);
    if ((!parallel || base >= nbits) && up_first_bits) {
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev[0], 0), "cudaStreamWaitEvent");
        up_first_bits = 0;
    }
    if (!parallel || base >= nbits) {
        /* short streams: one thread skims CDS after CDS (a dozen launches would cost more) */
        uint64_t *res = state;
        CK(aec_scan_offsets_launch(c, (const uint32_t *)d_in, in_bytes, start_bit, d_rsi_offsets, max_rsi, res, ctx->stream),
           "scan launch");
        ctx->launches += 1;
        CK(cudaMemcpyAsync(&ctx->h_res[12], res, 24, cudaMemcpyDeviceToHost, ctx->stream), "memcpy(scan)");
        CK(cudaStreamSynchronize(ctx->stream), "scan sync");
        if (found) *found = (size_t)ctx->h_res[12];
        ctx->scan_end = ctx->h_res[14];
        return (ctx->h_res[13] & 1ull) ? AEC_DATA_ERROR : AEC_OK;
    }
    /* Windows of nh bit positions; the tables of a window reach one worst-case RSI further so that
     * every RSI starting inside the window can be followed to its end. */
    const uint32_t LV = aec_skim_levels(c);
    const uint64_t margin = aec_skim_margin_bits(c);
    const uint64_t span = ((nbits - base) + 31ull) & ~31ull;
    uint64_t nh = scan_window_bits_for(ctx, span);

    const uint64_t nwin = (span + nh - 1) / nh;
    const uint64_t np_max = nh + margin < span ? nh + margin : span;
    if (np_max >= 0x7FFFFFFFull) { snprintf(ctx->err, sizeof ctx->err, "scan window too large"); return AEC_CONF_ERROR; }
    /* two table sets: the walk through window i (one thread, a dependent load per RSI) runs on a side
     * stream next to the table kernels of window i+1 */
    /* Streams of many short RSIs (well-compressed data, small chunks): the one-load-per-RSI walk would take
     * longer than the tables; three more passes give the length of eight RSIs in a row and the walk an
     * eighth of its steps.  Decided from what the caller expects: RSIs asked for per window of stream. */
    /* RSIs expected per 2^25 bits of stream: the walk's cost against the tables' is a matter of RSI density */
    const double rsis_per_window = (double)max_rsi * (double)(1ull << 25) / (double)span;
    /* The walk costs about 0.8 us per RSI while table kernels run next to it, the tables of a window about
     * 0.9 ms (1.15 ms dense), the long-jump passes 0.13 ms over the candidate list: measured break-even around 1000
     * RSIs per window (profiles/r2_summary.md). */
    const bool sparse = ctx->scan_sparse && LV >= aec_skim_sparse_min_levels();
    const bool skip8 = ctx->scan_skip8 == 1 || (ctx->scan_skip8 < 0 && rsis_per_window > (sparse ? 1000.0 : 1500.0));
    const size_t cand_cap = (sparse && skip8) ? (size_t)((np_max / 4u + 63u) & ~63ull) : 0;
    /* the levels, H and R (and two doubling buffers, and the list of candidates) */
    const size_t set_words = (size_t)(LV + 2u + (skip8 ? 2u : 0u)) * (size_t)np_max + cand_cap;
    const int nsets = nwin > 1 ? 2 : 1;
    CK(ctx->skim_tab.ensure(set_words * 4u * (size_t)nsets), "cudaMalloc(skim tables)");
    if (!ctx->s_walk) CK(cudaStreamCreateWithFlags(&ctx->s_walk, cudaStreamNonBlocking), "cudaStreamCreate(walk)");
    for (cudaEvent_t &e : ctx->ev_skim)
        if (!e) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    for (cudaEvent_t &e : ctx->ev_prog)
        if (!e) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    AecSkimArgs a;
    memset(&a, 0, sizeof a);
    a.cfg = c;
    a.in_words = (const uint32_t *)d_in;
    a.nbits = nbits;
    a.LV = LV;
    a.state = state;
    a.offsets = d_rsi_offsets;
    a.max_rsi = max_rsi;
    a.grp_index = d_grp;
    a.grp_G = aec_decode_group_blocks(c);
    ctx->scan_grp = d_grp != nullptr;
    for (uint64_t i = 0; i < nwin; i++) {
        const int k = (int)(i & 1u) % nsets;
        a.T = (uint32_t *)ctx->skim_tab.p + (size_t)k * set_words;
        a.H = a.T + (size_t)LV * (size_t)np_max;
        a.R = a.H + (size_t)np_max;
        a.H8 = skip8 ? a.R + (size_t)np_max : nullptr;
        a.sparse = sparse ? 1u : 0u;
        a.set = (uint32_t)k;
        a.cand_list = cand_cap ? a.H8 + 2u * (size_t)np_max : nullptr;
        a.cand_cap = (uint32_t)cand_cap;
        a.wb = base + i * nh;
        const uint64_t rem = ((nbits - a.wb) + 31ull) & ~31ull;
        a.np = (uint32_t)(nh + margin < rem ? nh + margin : rem);
        a.last = (i + 1 == nwin) ? 1u : 0u;
        a.nh_eff = a.last ? a.np : (uint32_t)nh;
        if (i >= 2) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_skim[2 + k], 0), "cudaStreamWaitEvent");   /* set k is free again */
        if (up_first_bits && a.wb + a.np + 8192ull > up_first_bits) {      /* this window reads beyond the part uploaded first */
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev[0], 0), "cudaStreamWaitEvent");
            up_first_bits = 0;
        }
        CK(aec_skim_window_launch(a, ctx->stream), "skim launch");
        CK(cudaEventRecord(ctx->ev_skim[k], ctx->stream), "cudaEventRecord");
        CK(cudaStreamWaitEvent(ctx->s_walk, ctx->ev_skim[k], 0), "cudaStreamWaitEvent");
        CK(aec_skim_walk_launch(a, ctx->s_walk), "walk launch");
        CK(cudaEventRecord(ctx->ev_skim[2 + k], ctx->s_walk), "cudaEventRecord");
        if (prog) {
            const int s4 = (int)(i & 3u);
            CK(cudaMemcpyAsync(&ctx->h_res[16 + 4 * s4], state, 32, cudaMemcpyDeviceToHost, ctx->s_walk), "memcpy(scan progress)");
            CK(cudaEventRecord(ctx->ev_prog[s4], ctx->s_walk), "cudaEventRecord");
        }
        /* level 0, LV - 1 doubling passes, RSI lengths (+ begin and the candidates' pass), the long jumps, walk (+ fill), group index */
        ctx->launches += 1u + (LV - 1u) + 1u + (sparse ? 2u : 0u) + (skip8 ? (sparse ? 5u : 4u) : 0u) + 1u + (d_grp ? 1u : 0u);
        if (prog && i >= 2) {
            /* The walk through the window two back has finished long ago; two windows of work are queued
             * behind it, so the device stays busy while the host hands that window's RSIs on (a copy to
             * pageable memory holds the host for its whole duration).  Every RSI the walk found except
             * the last has its successor's offset too. */
            const int sp = (int)((i - 2) & 3u);
            CK(cudaEventSynchronize(ctx->ev_prog[sp]), "scan progress sync");
            const uint64_t f = ctx->h_res[16 + 4 * sp + 1];
            if (f > 1) { rc = prog->consume(f - 1, ctx->ev_prog[sp]); if (rc != AEC_OK) return rc; }
        }
    }
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_skim[2 + (int)((nwin - 1) & 1u) % nsets], 0), "cudaStreamWaitEvent");
    CK(cudaMemcpyAsync(&ctx->h_res[12], state, 32, cudaMemcpyDeviceToHost, ctx->stream), "memcpy(scan)");
    CK(cudaStreamSynchronize(ctx->stream), "scan sync");
    if (found) *found = (size_t)ctx->h_res[13];
    ctx->scan_end = ctx->h_res[12];
    ctx->scan_fast = ctx->h_res[15];
    return (ctx->h_res[14] & 2ull) ? AEC_DATA_ERROR : AEC_OK;
}

void aecb200_ctx_set_scan_mode(aecb200_ctx *ctx, int mode, uint64_t window_bits)
{
    if (!ctx) return;
    ctx->scan_mode = mode;
    if (window_bits) ctx->scan_window_bits = window_bits;
}

uint64_t aecb200_ctx_last_scan_fast(aecb200_ctx *ctx) { return ctx ? ctx->scan_fast : 0; }

/* Streaming decode: the next aecb200_decode_host_resume call keeps the compressed bytes it uploads in
 * HBM and sends only bytes it has not seen (in[0] is byte stream_byte0 of the stream, a multiple of 4).
 * stream_byte0 < 0 forgets the accumulated stream. */
void aecb200_ctx_accumulate_next(aecb200_ctx *ctx, long long stream_byte0)
{
    if (!ctx) return;
    if (stream_byte0 < 0) { ctx->acc_valid = false; ctx->acc_len = 0; ctx->acc_stream_byte0 = -1; }
    else ctx->acc_stream_byte0 = stream_byte0;
}
uint64_t aecb200_ctx_accumulated_uploads(aecb200_ctx *ctx) { return ctx ? ctx->acc_uploaded : 0; }

size_t aecb200_ctx_found_offsets(aecb200_ctx *ctx, uint64_t *dst, size_t cap)
{
    if (!ctx) return 0;
    const size_t n = ctx->found_offs.size();
    if (dst) for (size_t i = 0; i < n && i < cap; i++) dst[i] = ctx->found_offs[i];
    return n;
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
    ENTER_DEVICE();

    const bool staged = ctx->in_stage_ready;             /* input already sits in in_stage (SZIP shim) */
    ctx->in_stage_ready = false;
    const size_t rsi_bytes = (size_t)c.R * c.B;
    size_t use_bytes = final ? (in_bytes / c.B) * c.B : (in_bytes / rsi_bytes) * rsi_bytes;
    const uint64_t nrsi = (use_bytes / c.B + c.R - 1) / c.R;
    const uint32_t phase = (uint32_t)(carry->bits & 31u);

    uint64_t end_bits = phase;
    uint32_t end_k = carry->k;
    size_t bound = aecb200_encode_bound(p, use_bytes) + 8;
    /* pieces of whole RSIs; a piece starts at the bit where the one before ended, so the pieces
     * write one contiguous stream into out_stage (the same seeding as AEC_NO_FLUSH streaming) */
    size_t piece = use_bytes;
    if (!staged && ctx->pipe_piece && use_bytes >= 2 * ctx->pipe_piece) {
        const size_t unit = rsi_bytes * 16;             /* pieces start 16-byte aligned (vector loads) */
        piece = ((ctx->pipe_piece + unit - 1) / unit) * unit;
        if ((use_bytes + piece - 1) / piece > 256) piece = (((use_bytes + 255) / 256 + unit - 1) / unit) * unit;
    }
    const size_t npieces = use_bytes ? (use_bytes + piece - 1) / piece : 0;
    size_t copied = 0;                                   /* bytes of the stream already on their way to `out` */
    if (use_bytes) {
        CK(ctx->in_stage.ensure(use_bytes + 16), "cudaMalloc(in)");
        CK(ctx->out_stage.ensure(bound), "cudaMalloc(out)");
        uint64_t *d_offs = nullptr;
        if (rsi_offsets && offsets_cap) {
            CK(ctx->offs.ensure(nrsi * 8), "cudaMalloc(offsets)");
            d_offs = (uint64_t *)ctx->offs.p;
        }
        aecb200_carry seed = {phase, carry->k, carry->word};
        if (npieces == 1) {
            if (!staged) CK(copy_h2d(ctx, ctx->in_stage.p, in, use_bytes, ctx->stream), "H2D");
            rc = aecb200_encode_device(ctx, p, ctx->in_stage.p, use_bytes, ctx->out_stage.p, ctx->out_stage.cap & ~(size_t)3,
                                       &seed, d_offs);
            if (rc != AEC_OK) return rc;
            aecb200_carry e;
            rc = aecb200_encode_finish(ctx, &e);
            if (rc != AEC_OK) return rc;
            end_bits = e.bits; end_k = e.k;
        } else {
            rc = pipe_prepare(ctx, npieces);
            if (rc != AEC_OK) return rc;
            PipeGuard guard(ctx);
            for (size_t i = 0; i < npieces; i++) {       /* all uploads are queued now and run back to back */
                const size_t o = i * piece, nb = (o + piece <= use_bytes) ? piece : use_bytes - o;
                CK(copy_h2d(ctx, (uint8_t *)ctx->in_stage.p + o, (const uint8_t *)in + o, nb, ctx->s_in), "H2D");
                CK(cudaEventRecord(ctx->ev[i], ctx->s_in), "cudaEventRecord");
            }
            for (size_t i = 0; i < npieces; i++) {
                const size_t o = i * piece, nb = (o + piece <= use_bytes) ? piece : use_bytes - o;
                CK(cudaStreamWaitEvent(ctx->stream, ctx->ev[i], 0), "cudaStreamWaitEvent");
                rc = aecb200_encode_device(ctx, p, (uint8_t *)ctx->in_stage.p + o, nb, ctx->out_stage.p,
                                           ctx->out_stage.cap & ~(size_t)3, &seed,
                                           d_offs ? d_offs + (o / rsi_bytes) : nullptr);
                aecb200_carry e = {0, 0, 0};
                if (rc == AEC_OK) rc = aecb200_encode_finish(ctx, &e);
                if (rc != AEC_OK) return rc;
                end_bits = e.bits; end_k = e.k;
                seed.bits = e.bits; seed.k = e.k; seed.word = 0;
                if (i + 1 < npieces) {
                    /* words in front of the stream's last, unfinished one are final: send them out
                     * while the next piece is coded */
                    size_t fin = (size_t)(end_bits >> 5) << 2;
                    if (fin > out_cap) fin = out_cap;
                    if (fin > copied) {
                        CK(copy_d2h(ctx, (uint8_t *)out + copied, (uint8_t *)ctx->out_stage.p + copied, fin - copied, ctx->s_out), "D2H");
                        copied = fin;
                    }
                    if (end_bits & 31u) {
                        /* the next piece completes that word: hand it the bits it already holds */
                        uint8_t *hb = (uint8_t *)&ctx->h_res[11];
                        CK(cudaMemcpyAsync(hb, (uint8_t *)ctx->out_stage.p + ((size_t)(end_bits >> 5) << 2), 4,
                                           cudaMemcpyDeviceToHost, ctx->stream), "D2H word");
                        CK(cudaStreamSynchronize(ctx->stream), "sync");
                        const uint32_t w = ((uint32_t)hb[0] << 24) | ((uint32_t)hb[1] << 16) | ((uint32_t)hb[2] << 8) | hb[3];
                        seed.word = w & ~(0xFFFFFFFFu >> (end_bits & 31u));
                    }
                }
            }
            guard.armed = false;                         /* s_in is drained, s_out is waited for below */
        }
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
        if (ncopy > copied)
            CK(copy_d2h(ctx, (uint8_t *)out + copied, (uint8_t *)ctx->out_stage.p + copied, ncopy - copied, ctx->stream), "D2H");
        if (!final && (end_bits & 7u)) {
            CK(cudaMemcpyAsync(&ctx->h_res[10], (uint8_t *)ctx->out_stage.p + (end_bits / 8), 1,
                               cudaMemcpyDeviceToHost, ctx->stream), "D2H tail");
        }
        CK(cudaStreamSynchronize(ctx->stream), "sync");
        if (npieces > 1) CK(cudaStreamSynchronize(ctx->s_out), "sync(out)");
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

/* AEC_NO_FLUSH encoding accumulates its input on the device: bytes that do not yet make a whole RSI are
 * sent to the context's input stage as they arrive (offset = bytes staged so far); once an RSI is
 * complete (or the stream is flushed) aecb200_encode_host_piece is called with in == NULL and codes
 * the staged bytes without another upload. */
int aecb200_ctx_stage_input(aecb200_ctx *ctx, size_t offset, const void *src, size_t n)
{
    if (!ctx || (!src && n)) return AEC_CONF_ERROR;
    ENTER_DEVICE();
    CK(ctx->in_stage.ensure_preserve(offset + n + 16, offset, ctx->stream), "cudaMalloc(in)");
    if (n) CK(cudaMemcpyAsync((uint8_t *)ctx->in_stage.p + offset, src, n, cudaMemcpyHostToDevice, ctx->stream), "H2D (accumulate)");
    ctx->staged_uploads += n;
    return AEC_OK;
}
uint64_t aecb200_ctx_staged_uploads(aecb200_ctx *ctx) { return ctx ? ctx->staged_uploads : 0; }

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
    if (!in && in_bytes) ctx->in_stage_ready = true;    /* staged by aecb200_ctx_stage_input */
    return encode_host_impl(ctx, p, in, in_bytes, final, out, out_cap, out_len, in_consumed, carry,
                            rsi_offsets, offsets_cap, n_offsets);
}

/* RSIs the boundary discovery has finished with are decoded on the context's second stream/workspace and
 * copied to the caller while the discovery goes on (ScanProgress hook of scan_offsets_impl). */
/* group index entries of RSIs the discovery could not take from its tables (marked SK_GRP_MISSING) */
static int complete_group_index(aecb200_ctx *ctx, cudaStream_t st, const AecCfg &c, const uint8_t *d_stream, size_t nbytes,
                                const uint64_t *d_offs, size_t nrsi, uint64_t *d_grp)
{
    AecDecArgs a;
    memset(&a, 0, sizeof a);
    a.cfg = c;
    a.in_words = (const uint32_t *)d_stream;
    a.in_bytes = nbytes;
    a.rsi_offsets = d_offs;
    a.nrsi = nrsi;
    a.grp_G = aec_decode_group_blocks(c);
    CK(aec_build_group_index_launch(a, d_grp, st, 1), "group index launch");
    ctx->launches += 1;
    return AEC_OK;
}

struct EarlyDecode : ScanProgress {
    AecCfg c;
    bool on = false, failed = false, pending = false;
    aecb200_ctx *ctx = nullptr;
    const aecb200_params *p = nullptr;
    const uint8_t *d_stream = nullptr;
    size_t nbytes = 0, rsi_out = 0, pending_expect = 0;
    uint8_t *out = nullptr;
    uint64_t R = 0, full_rsis = 0, min_rsis = 1, done = 0;
    void finish_pending()
    {
        if (!pending) return;
        size_t got = 0;
        const int rc = aecb200_decode_finish(ctx->aux, &got);
        pending = false;
        if (rc != AEC_OK || got != pending_expect) failed = true;
    }
    int consume(uint64_t complete, cudaEvent_t offsets_ready) override
    {
        if (failed) return AEC_OK;
        if (complete > full_rsis) complete = full_rsis;             /* only RSIs wanted in full */
        if (complete < done + min_rsis) return AEC_OK;
        finish_pending();
        if (failed) return AEC_OK;
        aecb200_ctx *aux = ctx->aux;
        if (cudaStreamWaitEvent(aux->stream, offsets_ready, 0) != cudaSuccess) { failed = true; return AEC_OK; }
        const uint64_t r0 = done, r1 = complete;
        const size_t nb = (size_t)(r1 - r0) * rsi_out;
        uint8_t *d_out = (uint8_t *)ctx->out_stage.p + (size_t)r0 * rsi_out;
        const uint64_t *d_grp = nullptr;
        if (ctx->scan_grp) {
            if (complete_group_index(aux, aux->stream, c, d_stream, nbytes, (const uint64_t *)ctx->offs.p + r0, (size_t)(r1 - r0),
                                     (uint64_t *)ctx->grp.p + r0 * 32) != AEC_OK) { failed = true; return AEC_OK; }
            d_grp = (const uint64_t *)ctx->grp.p + r0 * 32;
        }
        int rc = aecb200_decode_device_indexed(aux, p, d_stream, nbytes, (const uint64_t *)ctx->offs.p + r0, (size_t)(r1 - r0),
                                               d_grp, d_out, nb);
        if (rc != AEC_OK) { failed = true; return AEC_OK; }
        if (copy_d2h(aux, out + (size_t)r0 * rsi_out, d_out, nb, aux->stream) != cudaSuccess) failed = true;
        pending = true; pending_expect = nb; done = r1;
        ctx->launches += 2;
        return AEC_OK;
    }
};

#define AECB200_NOT_PIPELINED 1000
static int decode_host_pipelined(aecb200_ctx *ctx, const aecb200_params *p, const AecCfg &c,
                                 const void *in, size_t in_bytes,
                                 const uint64_t *rsi_offsets, size_t n_offsets,
                                 void *out, size_t out_cap, size_t *out_len);

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
    ENTER_DEVICE();
    const uint64_t want_new = out_cap / c.B;                       /* samples the caller can take */
    const uint64_t out_samples = skip_samples + want_new;          /* counted from the RSI at start_bit */
    uint64_t need_rsi = (out_samples + c.R - 1) / c.R;
    if (want_new == 0 || in_bytes * 8ull <= start_bit)
        return AEC_OK;
    if (rsi_offsets && skip_samples == 0) {
        /* large indexed ranges from the start of an RSI: the three-stream pipeline */
        size_t first = 0;
        while (first < n_offsets && rsi_offsets[first] < start_bit) first++;
        if (first < n_offsets && rsi_offsets[first] == start_bit) {
            size_t written = 0;
            rc = decode_host_pipelined(ctx, p, c, in, in_bytes, rsi_offsets + first, n_offsets - first, out,
                                       (size_t)(want_new * c.B), &written);
            if (rc == AEC_OK) {
                const uint64_t W = written / c.B, full = W / c.R;
                const size_t used = (n_offsets - first) < need_rsi ? (n_offsets - first) : (size_t)need_rsi;
                uint64_t rb;
                if (full < used) rb = rsi_offsets[first + full];
                else rb = (first + used < n_offsets) ? rsi_offsets[first + used] : (uint64_t)in_bytes * 8ull;
                if (out_len) *out_len = written;
                if (resume_bit) *resume_bit = rb;
                if (resume_delivered) *resume_delivered = (size_t)(W % c.R);
                return AEC_OK;
            }
            if (rc != AECB200_NOT_PIPELINED) return rc;
        }
    }
    /* stage the stream from the 32-bit word that holds start_bit */
    const size_t base_byte = (size_t)((start_bit >> 5) << 2);
    const uint64_t base_bit = (uint64_t)base_byte * 8ull;
    const size_t nbytes = in_bytes - base_byte;
    const size_t in_pad = (nbytes + 3) & ~(size_t)3;
    CK(ctx->out_stage.ensure((size_t)(out_samples * c.B) + 16), "cudaMalloc(out)");
    CK(ctx->offs.ensure((need_rsi + 1) * 8), "cudaMalloc(offsets)");
    const uint8_t *d_stream;                                     /* device copy of in[base_byte .. in_bytes) */
    const long long acc0 = ctx->acc_stream_byte0;
    ctx->acc_stream_byte0 = -1;
    if (acc0 >= 0) {
        /* accumulate: bytes of this stream uploaded by earlier calls are still in acc_stage */
        const uint64_t A = (uint64_t)acc0 + base_byte, E = (uint64_t)acc0 + in_bytes;
        if (!ctx->acc_valid || A < ctx->acc_start || ctx->acc_start + ctx->acc_len > E || A > ctx->acc_start + ctx->acc_len ||
            A - ctx->acc_start > ((uint64_t)64 << 20)) {
            ctx->acc_start = A; ctx->acc_len = 0; ctx->acc_valid = true;
        }
        const size_t want_len = (size_t)(E - ctx->acc_start);
        CK(ctx->acc_stage.ensure_preserve(want_len + 16, ctx->acc_len, ctx->stream), "cudaMalloc(stream)");
        if (want_len > ctx->acc_len) {
            CK(cudaMemcpyAsync((uint8_t *)ctx->acc_stage.p + ctx->acc_len,
                               (const uint8_t *)in + (size_t)(ctx->acc_start + ctx->acc_len - (uint64_t)acc0), want_len - ctx->acc_len,
                               cudaMemcpyHostToDevice, ctx->stream), "H2D");
            ctx->acc_uploaded += want_len - ctx->acc_len;
        }
        ctx->acc_len = want_len;
        CK(cudaMemsetAsync((uint8_t *)ctx->acc_stage.p + want_len, 0, 8, ctx->stream), "memset(stream tail)");
        d_stream = (const uint8_t *)ctx->acc_stage.p + (size_t)(A - ctx->acc_start);
    } else {
        CK(ctx->in_stage.ensure(in_pad + 16), "cudaMalloc(in)");
        /* Without an index the discovery starts at the front of the stream: upload what its first two windows
         * read, start, and let the rest of the stream follow on the upload stream meanwhile (the window
         * that first reaches beyond waits for it: scan_offsets_impl). */
        size_t first = nbytes;
        ctx->up_first_bits = 0;
        if (!rsi_offsets && ctx->pipe_piece && nbytes > ((size_t)32 << 20) && pipe_prepare(ctx, 1) == AEC_OK) {
            first = (size_t)((2 * scan_window_bits_for(ctx, (uint64_t)nbytes * 8ull) + aec_skim_margin_bits(c)) / 8ull) + 65536;
            first &= ~(size_t)15;
            if (first >= nbytes) first = nbytes;
        }
        if (first == nbytes) CK(cudaMemsetAsync((uint8_t *)ctx->in_stage.p + (in_pad - 4), 0, 4, ctx->stream), "memset(in tail)");
        CK(copy_h2d(ctx, ctx->in_stage.p, (const uint8_t *)in + base_byte, first, ctx->stream), "H2D");
        if (first < nbytes) {
            CK(cudaMemsetAsync((uint8_t *)ctx->in_stage.p + (in_pad - 4), 0, 4, ctx->s_in), "memset(in tail)");
            CK(copy_h2d(ctx, (uint8_t *)ctx->in_stage.p + first, (const uint8_t *)in + base_byte + first, nbytes - first, ctx->s_in), "H2D");
            CK(cudaEventRecord(ctx->ev[0], ctx->s_in), "cudaEventRecord");
            ctx->up_first_bits = (uint64_t)first * 8ull;
        }
        d_stream = (const uint8_t *)ctx->in_stage.p;
    }
    size_t nrsi = 0;
    uint64_t scan_end = 0;
    uint64_t *h_offs = nullptr;
    bool scan_grp = false;
    EarlyDecode early;
    ctx->found_offs.clear();
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
        /* Large outputs: RSIs whose offsets are final are decoded on a second context and their
         * samples sent back while the discovery of the later ones goes on (the download of a README-size
         * buffer takes a fifth of the discovery time). */
        const size_t rsi_out = (size_t)c.R * c.B;
        if (skip_samples == 0 && acc0 < 0 && ctx->pipe_piece && (size_t)(out_samples * c.B) >= 2 * ctx->pipe_piece &&
            !ctx->careful_only) {
            if (!ctx->aux && aecb200_ctx_create(&ctx->aux, ctx->device) != AEC_OK) ctx->aux = nullptr;
            if (ctx->aux) {
                early.on = true;
                early.ctx = ctx; early.p = p; early.R = c.R;
                early.d_stream = d_stream; early.nbytes = nbytes;
                early.out = (uint8_t *)out; early.rsi_out = rsi_out;
                early.full_rsis = out_samples / c.R;
                early.min_rsis = (ctx->pipe_piece + rsi_out - 1) / rsi_out;
            }
        }
        CK(ctx->grp.ensure((size_t)need_rsi * 32 * 8), "cudaMalloc(group index)");
        early.c = c;
        rc = scan_offsets_impl(ctx, p, d_stream, nbytes, start_bit - base_bit,
                               (uint64_t *)ctx->offs.p, (size_t)need_rsi, &nrsi, early.on ? &early : nullptr,
                               ctx->careful_only ? nullptr : (uint64_t *)ctx->grp.p);
        if (ctx->up_first_bits) {                       /* whatever the scan did not wait for: the decode reads all of it */
            ctx->up_first_bits = 0;
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev[0], 0), "cudaStreamWaitEvent");
            if (ctx->aux) CK(cudaStreamWaitEvent(ctx->aux->stream, ctx->ev[0], 0), "cudaStreamWaitEvent");
        }
        scan_grp = ctx->scan_grp;
        if (early.on) {
            early.finish_pending();
            if (early.failed || rc != AEC_OK) {          /* anything unusual: decode everything on the plain path */
                cudaStreamSynchronize(ctx->aux->stream);
                early.done = 0;
            }
        }
        if (rc != AEC_OK && rc != AEC_DATA_ERROR) return rc;
        scan_end = ctx->scan_end;
        h_offs = (uint64_t *)malloc((nrsi + 1) * sizeof(uint64_t));
        if (!h_offs) return AEC_MEM_ERROR;
        if (nrsi) {
            cudaError_t e = cudaMemcpyAsync(h_offs, ctx->offs.p, nrsi * 8, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { free(h_offs); return fail_cuda(ctx, e, "D2H offsets"); }
        }
        ctx->found_offs.resize(nrsi);
        for (size_t i = 0; i < nrsi; i++) ctx->found_offs[i] = h_offs[i] + base_bit;
    }
    size_t written = 0;
    const uint64_t early_rsis = early.done < nrsi ? early.done : 0;    /* already decoded and on their way to `out` */
    const size_t early_bytes = (size_t)early_rsis * c.R * c.B;
    const uint64_t *d_grp = nullptr;
    if (scan_grp && nrsi > early_rsis) {
        /* the discovery wrote the group index of the RSIs it took from its tables; skim the few others */
        rc = complete_group_index(ctx, ctx->stream, c, d_stream, nbytes, (const uint64_t *)ctx->offs.p + early_rsis,
                                  nrsi - early_rsis, (uint64_t *)ctx->grp.p + early_rsis * 32);
        if (rc != AEC_OK) { free(h_offs); return rc; }
        d_grp = (const uint64_t *)ctx->grp.p + early_rsis * 32;
    }
    rc = aecb200_decode_device_indexed(ctx, p, d_stream, nbytes, (const uint64_t *)ctx->offs.p + early_rsis, nrsi - early_rsis,
                                       d_grp, (uint8_t *)ctx->out_stage.p + early_bytes,
                                       (size_t)(out_samples * c.B) - early_bytes);
    if (rc == AEC_OK) rc = aecb200_decode_finish(ctx, &written);
    if (rc != AEC_OK) { free(h_offs); if (early.on) cudaStreamSynchronize(ctx->aux->stream); return rc; }
    written += early_bytes;
    uint64_t W = written / c.B;                                    /* samples decoded from start_bit */
    /* only RSIs that delivered samples count as discovered (the scan also notes where the zero padding
     * behind the last RSI begins) */
    if (ctx->found_offs.size() > (size_t)((W + c.R - 1) / c.R)) ctx->found_offs.resize((size_t)((W + c.R - 1) / c.R));
    size_t newbytes = W > skip_samples ? (size_t)((W - skip_samples) * c.B) : 0;
    if (newbytes > early_bytes) {
        /* (early pieces only exist with skip_samples == 0) */
        cudaError_t e = copy_d2h(ctx, (uint8_t *)out + early_bytes, (uint8_t *)ctx->out_stage.p + skip_samples * c.B + early_bytes,
                                 newbytes - early_bytes, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { free(h_offs); return fail_cuda(ctx, e, "D2H"); }
    }
    if (early.on) {
        cudaError_t e = cudaStreamSynchronize(ctx->aux->stream);         /* the early pieces have arrived */
        if (e != cudaSuccess) { free(h_offs); return fail_cuda(ctx, e, "D2H (early pieces)"); }
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

/* Whole-buffer decode of a large stream with a caller-supplied RSI offset index, as a pipeline of
 * RSI ranges: a range is decoded as soon as the bytes up to its last bit have arrived, and its
 * samples travel back while the next range is decoded.  Anything but a clean decode of every range
 * (short index, truncated or damaged stream) returns AECB200_NOT_PIPELINED and the caller repeats
 * the call on the one-piece path, which reproduces the reference's partial results. */
static int decode_host_pipelined(aecb200_ctx *ctx, const aecb200_params *p, const AecCfg &c,
                                 const void *in, size_t in_bytes,
                                 const uint64_t *rsi_offsets, size_t n_offsets,
                                 void *out, size_t out_cap, size_t *out_len)
{
    const size_t rsi_bytes = (size_t)c.R * c.B;
    const uint64_t out_samples = out_cap / c.B;
    const uint64_t need_rsi = (out_samples + c.R - 1) / c.R;
    if (!ctx->pipe_piece || !rsi_offsets || need_rsi > n_offsets || need_rsi < 2 ||
        out_samples * c.B < 2 * ctx->pipe_piece || in_bytes < 8)
        return AECB200_NOT_PIPELINED;
    for (uint64_t r = 0; r < need_rsi; r++)                     /* a usable index ascends inside the stream */
        if (rsi_offsets[r] >= (uint64_t)in_bytes * 8ull || (r && rsi_offsets[r] <= rsi_offsets[r - 1]))
            return AECB200_NOT_PIPELINED;
    uint64_t per = (ctx->pipe_piece + rsi_bytes - 1) / rsi_bytes;            /* RSIs per piece */
    if ((need_rsi + per - 1) / per > 240) per = (need_rsi + 239) / 240;
    /* piece i covers RSIs [first[i], first[i+1]) (short leading pieces were tried and lost to the
     * fixed cost of a piece: gpurun e2e_sweep3, profiles/r1_f_summary.md) */
    std::vector<uint64_t> first;
    for (uint64_t r = 0; r < need_rsi; r += per) first.push_back(r);
    first.push_back(need_rsi);
    const size_t npieces = first.size() - 1;
    ENTER_DEVICE();
    int rc = pipe_prepare(ctx, 2 * npieces);
    if (rc != AEC_OK) return rc;
    PipeGuard guard(ctx);
    const size_t in_pad = (in_bytes + 3) & ~(size_t)3;
    CK(ctx->in_stage.ensure(in_pad + 16), "cudaMalloc(in)");
    CK(ctx->out_stage.ensure((size_t)(out_samples * c.B) + 16), "cudaMalloc(out)");
    CK(ctx->offs.ensure((need_rsi + 1) * 8), "cudaMalloc(offsets)");
    CK(cudaMemsetAsync((uint8_t *)ctx->in_stage.p + (in_pad - 4), 0, 4, ctx->s_in), "memset(in tail)");
    CK(cudaMemcpyAsync(ctx->offs.p, rsi_offsets, need_rsi * 8, cudaMemcpyHostToDevice, ctx->s_in), "H2D offsets");
    size_t up = 0;                                              /* bytes of the stream queued for upload */
    for (size_t i = 0; i < npieces; i++) {
        const uint64_t r1 = first[i + 1] < need_rsi ? first[i + 1] : need_rsi;
        /* through the last bit of the range, plus what the readers prefetch beyond it */
        size_t upto = r1 < need_rsi ? (size_t)(rsi_offsets[r1] / 8) + 256 : in_bytes;
        if (upto > in_bytes || i + 1 == npieces) upto = in_bytes;
        if (upto > up) {
            CK(copy_h2d(ctx, (uint8_t *)ctx->in_stage.p + up, (const uint8_t *)in + up, upto - up, ctx->s_in), "H2D");
            up = upto;
        }
        CK(cudaEventRecord(ctx->ev[i], ctx->s_in), "cudaEventRecord");
    }
    /* The warp-per-RSI decoder wants the group index of its RSIs.  Building it (one lane skims one
     * RSI) is latency bound and takes about as long for one piece as for the whole stream, so the
     * builds run ahead of the decoder on side streams of their own, several at a time. */
    const bool fast = aec_decode_warp_warps(c) != 0 && !ctx->careful_only;
    if (fast) {
        CK(ctx->grp.ensure(need_rsi * 32 * 8), "cudaMalloc(group index)");
        for (size_t i = 0; i < npieces; i++) {
            const uint64_t r0 = first[i], r1 = first[i + 1] < need_rsi ? first[i + 1] : need_rsi;
            cudaStream_t si = ctx->s_idx[i & 3];
            AecDecArgs a;
            memset(&a, 0, sizeof a);
            a.cfg = c;
            a.in_words = (const uint32_t *)ctx->in_stage.p;
            a.in_bytes = in_bytes;
            a.rsi_offsets = (const uint64_t *)ctx->offs.p + r0;
            a.nrsi = r1 - r0;
            a.grp_G = aec_decode_group_blocks(c);
            CK(cudaStreamWaitEvent(si, ctx->ev[i], 0), "cudaStreamWaitEvent");
            CK(aec_build_group_index_launch(a, (uint64_t *)ctx->grp.p + r0 * 32, si), "group index launch");
            ctx->launches += 1;
            CK(cudaEventRecord(ctx->ev[npieces + i], si), "cudaEventRecord");
        }
    }
    size_t total = 0;
    for (size_t i = 0; i < npieces; i++) {
        const uint64_t r0 = first[i], r1 = first[i + 1] < need_rsi ? first[i + 1] : need_rsi;
        const size_t o = (size_t)r0 * rsi_bytes;
        size_t nb = (size_t)(r1 - r0) * rsi_bytes;
        if (o + nb > (size_t)(out_samples * c.B)) nb = (size_t)(out_samples * c.B) - o;
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev[fast ? npieces + i : i], 0), "cudaStreamWaitEvent");
        size_t got = 0;
        rc = aecb200_decode_device_indexed(ctx, p, ctx->in_stage.p, in_bytes, (const uint64_t *)ctx->offs.p + r0, (size_t)(r1 - r0),
                                   fast ? (const uint64_t *)ctx->grp.p + r0 * 32 : nullptr,
                                   (uint8_t *)ctx->out_stage.p + o, nb);
        if (rc == AEC_OK) rc = aecb200_decode_finish(ctx, &got);
        if (rc != AEC_OK || got != nb)
            return (rc == AEC_OK || rc == AEC_DATA_ERROR) ? AECB200_NOT_PIPELINED : rc;
        CK(copy_d2h(ctx, (uint8_t *)out + o, (uint8_t *)ctx->out_stage.p + o, nb, ctx->s_out), "D2H");
        total += nb;
    }
    CK(cudaStreamSynchronize(ctx->s_out), "sync(out)");
    guard.armed = false;
    if (out_len) *out_len = total;
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
    rc = aecb200_decode_host_resume(ctx, p, in, in_bytes, rsi_offsets, n_offsets,
                                    (rsi_offsets && n_offsets) ? rsi_offsets[0] : 0, 0,
                                    out, out_cap, &written, nullptr, nullptr);
    if (rc != AEC_OK) return rc;
    if (out_len) *out_len = written;
    size_t left = out_cap - written;
    if (left > 0 && left < c.B) return AEC_MEM_ERROR;               /* decode.c:821-823 */
    return AEC_OK;
}

/* ------------------------------------------------------------------------ */
/* SZIP shim (szlib.h) with the byte shuffles on the device                  */
/* ------------------------------------------------------------------------ */

#define SZ_MSB_OPTION_MASK 16
#define SZ_NN_OPTION_MASK 32
#define SZ_OUTBUFF_FULL 2

namespace {
struct SzGeom {
    aecb200_params prm;
    int planes;
    uint32_t ws, px;
    size_t line, full_line;
};

void sz_geometry(int options_mask, int bits_per_pixel, int pixels_per_block, int pixels_per_scanline, int enc, SzGeom *g)
{
    /* sz_compat.c:12-37, :125-142 / :200-214: only MSB and NN reach the coder */
    g->planes = bits_per_pixel == 32 || bits_per_pixel == 64;
    g->prm.bits_per_sample = g->planes ? 8u : (uint32_t)bits_per_pixel;
    g->prm.block_size = (uint32_t)pixels_per_block;
    g->prm.rsi = pixels_per_block > 0 ? (uint32_t)((pixels_per_scanline + pixels_per_block - 1) / pixels_per_block) : 0u;
    g->prm.flags = (enc ? AECF_NOT_ENFORCE : 0u) | ((options_mask & SZ_MSB_OPTION_MASK) ? AECF_MSB : 0u) |
                   ((options_mask & SZ_NN_OPTION_MASK) ? AECF_PREPROCESS : 0u);
    g->ws = g->planes ? (uint32_t)bits_per_pixel / 8u : 1u;
    g->px = g->prm.bits_per_sample > 16 ? 4u : (g->prm.bits_per_sample > 8 ? 2u : 1u);
    g->line = (size_t)pixels_per_scanline * g->px;
    g->full_line = (size_t)g->prm.rsi * g->prm.block_size * g->px;
}
} // namespace

int aecb200_sz_compress_host(aecb200_ctx *ctx, int options_mask, int bits_per_pixel, int pixels_per_block,
                             int pixels_per_scanline, const void *source, size_t source_len, void *dest, size_t *dest_len)
{
    if (!ctx || !dest_len) return AEC_CONF_ERROR;
    SzGeom g;
    sz_geometry(options_mask, bits_per_pixel, pixels_per_block, pixels_per_scanline, 1, &g);
    AecCfg c;
    int rc = make_cfg(ctx, &g.prm, 1, &c);
    if (rc != AEC_OK || pixels_per_scanline <= 0) return AEC_CONF_ERROR;
    const size_t nlines = (source_len / g.px + (size_t)pixels_per_scanline - 1) / (size_t)pixels_per_scanline;
    const size_t padded_len = g.full_line * nlines;
    const bool shuffle = g.planes || g.full_line != g.line || source_len != padded_len;
    size_t produced = 0, consumed = 0;
    aecb200_carry carry = {0, 0, 0};
    const void *in = source;
    if (shuffle && padded_len) {
        ENTER_DEVICE();
        CK(ctx->raw_stage.ensure(source_len + 16), "cudaMalloc(sz source)");
        CK(ctx->in_stage.ensure(padded_len + 16), "cudaMalloc(in)");
        CK(copy_h2d(ctx, ctx->raw_stage.p, source, source_len, ctx->stream), "H2D");
        CK(aec_sz_pack_launch((const uint8_t *)ctx->raw_stage.p, source_len, (uint8_t *)ctx->in_stage.p, padded_len, g.ws, g.line,
                              g.full_line, g.px, (g.prm.flags & AECF_PREPROCESS) ? 1u : 0u, ctx->stream), "sz pack launch");
        ctx->launches += 1;
        ctx->in_stage_ready = true;
        in = nullptr;
    }
    rc = encode_host_impl(ctx, &g.prm, in, padded_len, 1, dest, *dest_len, &produced, &consumed, &carry, nullptr, 0, nullptr);
    ctx->in_stage_ready = false;
    *dest_len = produced;
    return rc == AEC_STREAM_ERROR ? SZ_OUTBUFF_FULL : rc;
}

int aecb200_sz_decompress_host(aecb200_ctx *ctx, int options_mask, int bits_per_pixel, int pixels_per_block,
                               int pixels_per_scanline, const void *source, size_t source_len, void *dest, size_t *dest_len)
{
    if (!ctx || !dest_len) return AEC_CONF_ERROR;
    SzGeom g;
    sz_geometry(options_mask, bits_per_pixel, pixels_per_block, pixels_per_scanline, 0, &g);
    AecCfg c;
    int rc = make_cfg(ctx, &g.prm, 0, &c);
    if (rc != AEC_OK || pixels_per_scanline <= 0) return AEC_CONF_ERROR;
    const bool ragged = (pixels_per_scanline % pixels_per_block) != 0;
    size_t written = 0;
    if (!ragged && !g.planes) {
        rc = aecb200_decode_host(ctx, &g.prm, source, source_len, nullptr, 0, dest, *dest_len, &written);
        if (rc != AEC_OK) return rc;
        if (written < *dest_len) *dest_len = written;
        return AEC_OK;
    }
    size_t nlines = 0, cap = *dest_len;
    if (ragged) {
        nlines = (*dest_len / g.px + (size_t)pixels_per_scanline - 1) / (size_t)pixels_per_scanline;
        cap = g.full_line * nlines;
    }
    ENTER_DEVICE();
    /* decode into out_stage (the padded, plane-ordered samples stay in HBM) */
    const uint64_t out_samples = cap / c.B;
    const uint64_t need_rsi = (out_samples + c.R - 1) / c.R;
    if (out_samples && source_len) {
        const size_t in_pad = (source_len + 3) & ~(size_t)3;
        CK(ctx->in_stage.ensure(in_pad + 16), "cudaMalloc(in)");
        CK(ctx->out_stage.ensure((size_t)(out_samples * c.B) + 16), "cudaMalloc(out)");
        CK(ctx->offs.ensure((need_rsi + 1) * 8), "cudaMalloc(offsets)");
        CK(cudaMemsetAsync((uint8_t *)ctx->in_stage.p + (in_pad - 4), 0, 4, ctx->stream), "memset(in tail)");
        CK(copy_h2d(ctx, ctx->in_stage.p, source, source_len, ctx->stream), "H2D");
        size_t nrsi = 0;
        rc = aecb200_scan_offsets_device(ctx, &g.prm, ctx->in_stage.p, source_len, 0, (uint64_t *)ctx->offs.p, (size_t)need_rsi, &nrsi);
        if (rc != AEC_OK && rc != AEC_DATA_ERROR) return rc;
        rc = aecb200_decode_device(ctx, &g.prm, ctx->in_stage.p, source_len, (const uint64_t *)ctx->offs.p, nrsi,
                                   ctx->out_stage.p, (size_t)(out_samples * c.B));
        if (rc == AEC_OK) rc = aecb200_decode_finish(ctx, &written);
        if (rc != AEC_OK) return rc;
    }
    if (cap - written > 0 && cap - written < c.B) return AEC_MEM_ERROR;      /* decode.c:821-823 */
    size_t total = written;
    if (ragged) total = nlines * g.line;                                     /* sz_compat.c:243-250 */
    if (total < *dest_len) *dest_len = total;
    const size_t n = *dest_len;
    if (n) {
        CK(ctx->out2_stage.ensure(n + 16), "cudaMalloc(sz out)");
        if (g.planes && n % g.ws)            /* bytes behind the last whole word are not written by the reference either */
            CK(cudaMemsetAsync(ctx->out2_stage.p, 0, n, ctx->stream), "memset(sz out)");
        CK(aec_sz_unpack_launch((const uint8_t *)ctx->out_stage.p, (uint8_t *)ctx->out2_stage.p, n, g.ws,
                                ragged ? g.line : (size_t)1 << 62, ragged ? g.full_line : (size_t)1 << 62, ctx->stream),
           "sz unpack launch");
        ctx->launches += 1;
        CK(copy_d2h(ctx, dest, ctx->out2_stage.p, n, ctx->stream), "D2H");
        CK(cudaStreamSynchronize(ctx->stream), "sync");
    }
    return AEC_OK;
}

} /* extern "C" */
