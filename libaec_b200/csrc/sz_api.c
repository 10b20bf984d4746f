/*
 * sz_api.c -- SZIP-compatible entry points (include/szlib.h) on top of the device layer.
 *
 * Drop-in for /root/reference/src/sz_compat.c:110-276.  The reference maps SZ options to AEC flags, turns
 * 32/64-bit pixels into byte planes, pads every scanline to a whole number of blocks (so that one
 * scanline == one RSI) with host loops and calls aec_buffer_encode / aec_buffer_decode.  Here the
 * caller's bytes go to the GPU as they are; planes and padding are produced (and undone) by kernels next
 * to the coder (aecb200_sz_compress_host / aecb200_sz_decompress_host, csrc/aec_runtime.cu, aec_sz.cu).
 * This file only borrows a context (CUDA stream + workspace) from the library's pool.
 */
#include <stdlib.h>
#include <string.h>

#include "../../include/aec_b200.h"
#include "../../include/szlib.h"

int SZ_BufftoBuffCompress(void *dest, size_t *destLen, const void *source, size_t sourceLen,
                          SZ_com_t *param)
{
    aecb200_ctx *ctx = aecb200_pool_get();
    if (!ctx) return SZ_MEM_ERROR;                       /* no CUDA device: there is no CPU coder to fall back to */
    int status = aecb200_sz_compress_host(ctx, param->options_mask, param->bits_per_pixel, param->pixels_per_block,
                                          param->pixels_per_scanline, source, sourceLen, dest, destLen);
    aecb200_pool_put(ctx);
    return status == AECB200_CUDA_ERROR ? SZ_MEM_ERROR : status;
}

int SZ_BufftoBuffDecompress(void *dest, size_t *destLen, const void *source, size_t sourceLen,
                            SZ_com_t *param)
{
    aecb200_ctx *ctx = aecb200_pool_get();
    if (!ctx) return SZ_MEM_ERROR;
    int status = aecb200_sz_decompress_host(ctx, param->options_mask, param->bits_per_pixel, param->pixels_per_block,
                                            param->pixels_per_scanline, source, sourceLen, dest, destLen);
    aecb200_pool_put(ctx);
    return status == AECB200_CUDA_ERROR ? SZ_MEM_ERROR : status;
}

int SZ_encoder_enabled(void) { return 1; }

/* netCDF's configure looks for this symbol (reference sz_compat.c:275-276) */
char SZ_Compress(void) { return SZ_OK; }
