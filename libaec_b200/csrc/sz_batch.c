/*
 * sz_batch.c -- many SZIP chunks in flight (BASELINE config 3: HDF5-style 4 MiB chunks).
 *
 * The reference codes one chunk per call on one core (sz_compat.c:110-268); an HDF5 filter pipeline
 * that wants throughput runs several such calls side by side.  Here a batch of chunks is spread over a
 * few host threads, each with its own pooled context (CUDA stream + workspace), so that the upload of
 * one chunk, the kernels of another and the download of a third overlap on the device.  Host code only:
 * the work itself is aecb200_sz_compress_host / aecb200_sz_decompress_host.
 */
#include <pthread.h>
#include <stdlib.h>

#include "../../include/aec_b200.h"

struct batch {
    int decompress, n, next;
    pthread_mutex_t lock;
    void *const *dest; size_t *dest_len;
    const void *const *source; const size_t *source_len;
    int mask, bpp, ppb, pps;
    int *status;
    int device;
};

extern int aecb200_set_device(int device);

static void *batch_worker(void *arg)
{
    struct batch *b = (struct batch *)arg;
    aecb200_set_device(b->device);                       /* worker threads start on device 0 */
    aecb200_ctx *ctx = aecb200_pool_get();
    for (;;) {
        pthread_mutex_lock(&b->lock);
        const int i = b->next < b->n ? b->next++ : -1;
        pthread_mutex_unlock(&b->lock);
        if (i < 0) break;
        if (!ctx) { b->status[i] = AECB200_CUDA_ERROR; continue; }
        b->status[i] = b->decompress
            ? aecb200_sz_decompress_host(ctx, b->mask, b->bpp, b->ppb, b->pps, b->source[i], b->source_len[i], b->dest[i], &b->dest_len[i])
            : aecb200_sz_compress_host(ctx, b->mask, b->bpp, b->ppb, b->pps, b->source[i], b->source_len[i], b->dest[i], &b->dest_len[i]);
    }
    aecb200_pool_put(ctx);
    return NULL;
}

static int run_batch(struct batch *b, int threads)
{
    if (b->n <= 0) return 0;
    if (threads <= 0) threads = 4;
    if (threads > b->n) threads = b->n;
    if (threads > 16) threads = 16;
    b->device = aecb200_current_device();
    b->next = 0;
    pthread_mutex_init(&b->lock, NULL);
    pthread_t tid[16];
    int started = 0;
    for (int t = 1; t < threads; t++)
        if (pthread_create(&tid[started], NULL, batch_worker, b) == 0) started++;
    batch_worker(b);                                     /* the calling thread works too */
    for (int t = 0; t < started; t++) pthread_join(tid[t], NULL);
    pthread_mutex_destroy(&b->lock);
    for (int i = 0; i < b->n; i++)
        if (b->status[i] != 0) return b->status[i];
    return 0;
}

int aecb200_sz_compress_batch(int n, void *const *dest, size_t *dest_len, const void *const *source, const size_t *source_len,
                              int options_mask, int bits_per_pixel, int pixels_per_block, int pixels_per_scanline,
                              int *status, int threads)
{
    struct batch b = {0};
    b.decompress = 0; b.n = n; b.dest = dest; b.dest_len = dest_len; b.source = source; b.source_len = source_len;
    b.mask = options_mask; b.bpp = bits_per_pixel; b.ppb = pixels_per_block; b.pps = pixels_per_scanline; b.status = status;
    return run_batch(&b, threads);
}

int aecb200_sz_decompress_batch(int n, void *const *dest, size_t *dest_len, const void *const *source, const size_t *source_len,
                                int options_mask, int bits_per_pixel, int pixels_per_block, int pixels_per_scanline,
                                int *status, int threads)
{
    struct batch b = {0};
    b.decompress = 1; b.n = n; b.dest = dest; b.dest_len = dest_len; b.source = source; b.source_len = source_len;
    b.mask = options_mask; b.bpp = bits_per_pixel; b.ppb = pixels_per_block; b.pps = pixels_per_scanline; b.status = status;
    return run_batch(&b, threads);
}
