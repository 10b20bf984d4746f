/*
 * sz_api.c -- SZIP-compatible entry points (include/szlib.h) on top of the
 * libaec stream API of this library.
 *
 * Drop-in for /root/reference/src/sz_compat.c:110-276.  The shim maps SZ
 * options to AEC flags, turns 32/64-bit pixels into byte planes, pads every
 * scanline to a whole number of blocks (so that one scanline == one RSI) and
 * calls aec_buffer_encode / aec_buffer_decode, which run on the GPU.
 */
#include <stdlib.h>
#include <string.h>

#include "../../include/szlib.h"

static unsigned sz_to_aec_flags(int mask)
{
    /* only MSB and NN influence the coding (reference sz_compat.c:12-27) */
    unsigned f = 0;
    if (mask & SZ_MSB_OPTION_MASK) f |= AEC_DATA_MSB;
    if (mask & SZ_NN_OPTION_MASK) f |= AEC_DATA_PREPROCESS;
    return f;
}

static size_t pixel_bytes(unsigned bits)
{
    return bits > 16 ? 4 : (bits > 8 ? 2 : 1);
}

/* word-interleaved bytes -> byte planes and back (reference sz_compat.c:39-69) */
static void to_planes(unsigned char *dst, const unsigned char *src, size_t n, size_t ws)
{
    size_t nw = n / ws;
    for (size_t w = 0; w < nw; w++)
        for (size_t j = 0; j < ws; j++)
            dst[j * nw + w] = src[w * ws + j];
}

static void from_planes(unsigned char *dst, const unsigned char *src, size_t n, size_t ws)
{
    size_t nw = n / ws;
    for (size_t w = 0; w < nw; w++)
        for (size_t j = 0; j < ws; j++)
            dst[w * ws + j] = src[j * nw + w];
}

int SZ_BufftoBuffCompress(void *dest, size_t *destLen, const void *source, size_t sourceLen,
                          SZ_com_t *param)
{
    struct aec_stream strm;
    memset(&strm, 0, sizeof strm);
    strm.block_size = (unsigned)param->pixels_per_block;
    strm.rsi = (unsigned)((param->pixels_per_scanline + param->pixels_per_block - 1) / param->pixels_per_block);
    strm.flags = AEC_NOT_ENFORCE | sz_to_aec_flags(param->options_mask);
    int planes = param->bits_per_pixel == 32 || param->bits_per_pixel == 64;
    strm.bits_per_sample = planes ? 8u : (unsigned)param->bits_per_pixel;

    const unsigned char *pix = (const unsigned char *)source;
    unsigned char *planebuf = NULL, *padded = NULL;
    int status;
    if (planes) {
        planebuf = (unsigned char *)malloc(sourceLen ? sourceLen : 1);
        if (!planebuf) return SZ_MEM_ERROR;
        to_planes(planebuf, pix, sourceLen, (size_t)param->bits_per_pixel / 8);
        pix = planebuf;
    }
    const size_t px = pixel_bytes(strm.bits_per_sample);
    const size_t line = (size_t)param->pixels_per_scanline * px;
    const size_t full_line = (size_t)strm.rsi * strm.block_size * px;
    const size_t nlines = (sourceLen / px + (size_t)param->pixels_per_scanline - 1) / (size_t)param->pixels_per_scanline;
    const size_t padded_len = full_line * nlines;

    const unsigned char *enc_in = pix;
    if (full_line != line || sourceLen != padded_len) {
        /* fill each scanline up to full_line with its last pixel (NN) or zeros */
        padded = (unsigned char *)malloc(padded_len ? padded_len : 1);
        if (!padded) { free(planebuf); return SZ_MEM_ERROR; }
        size_t rd = 0, wr = 0;
        while (rd < sourceLen) {
            size_t take = sourceLen - rd < line ? sourceLen - rd : line;
            memcpy(padded + wr, pix + rd, take);
            rd += take; wr += take;
            size_t fill = full_line - take;
            if (strm.flags & AEC_DATA_PREPROCESS) {
                for (size_t k = 0; k < fill; k += px) memcpy(padded + wr + k, pix + rd - px, px);
            } else {
                memset(padded + wr, 0, fill);
            }
            wr += fill;
        }
        enc_in = padded;
    }
    strm.next_in = enc_in;
    strm.avail_in = padded_len;
    strm.next_out = (unsigned char *)dest;
    strm.avail_out = *destLen;
    status = aec_buffer_encode(&strm);
    if (status == AEC_STREAM_ERROR) status = SZ_OUTBUFF_FULL;
    *destLen = strm.total_out;
    free(padded);
    free(planebuf);
    return status;
}

int SZ_BufftoBuffDecompress(void *dest, size_t *destLen, const void *source, size_t sourceLen,
                            SZ_com_t *param)
{
    struct aec_stream strm;
    memset(&strm, 0, sizeof strm);
    strm.block_size = (unsigned)param->pixels_per_block;
    strm.rsi = (unsigned)((param->pixels_per_scanline + param->pixels_per_block - 1) / param->pixels_per_block);
    strm.flags = sz_to_aec_flags(param->options_mask);
    int planes = param->bits_per_pixel == 32 || param->bits_per_pixel == 64;
    int ragged = param->pixels_per_scanline % param->pixels_per_block;
    strm.bits_per_sample = planes ? 8u : (unsigned)param->bits_per_pixel;
    const size_t px = pixel_bytes(strm.bits_per_sample);
    const size_t line = (size_t)param->pixels_per_scanline * px;
    const size_t full_line = (size_t)strm.rsi * strm.block_size * px;

    unsigned char *tmp = NULL;
    size_t nlines = 0, cap = *destLen;
    if (ragged) {
        nlines = (*destLen / px + (size_t)param->pixels_per_scanline - 1) / (size_t)param->pixels_per_scanline;
        cap = full_line * nlines;
    }
    if (ragged || planes) {
        tmp = (unsigned char *)malloc(cap ? cap : 1);
        if (!tmp) return SZ_MEM_ERROR;
    }
    strm.next_in = (const unsigned char *)source;
    strm.avail_in = sourceLen;
    strm.next_out = tmp ? tmp : (unsigned char *)dest;
    strm.avail_out = cap;
    int status = aec_buffer_decode(&strm);
    if (status != AEC_OK) { free(tmp); return status; }

    size_t total = strm.total_out;
    if (ragged) {
        /* squeeze the per-scanline padding out again */
        size_t wr = line;
        for (size_t rd = full_line; rd < strm.total_out; rd += full_line) {
            memmove(tmp + wr, tmp + rd, line);
            wr += line;
        }
        total = nlines * line;
    }
    if (total < *destLen) *destLen = total;
    if (planes) from_planes((unsigned char *)dest, tmp, *destLen, (size_t)param->bits_per_pixel / 8);
    else if (ragged) memcpy(dest, tmp, *destLen);
    free(tmp);
    return SZ_OK;
}

int SZ_encoder_enabled(void) { return 1; }

/* netCDF's configure looks for this symbol (reference sz_compat.c:275-276) */
char SZ_Compress(void) { return SZ_OK; }
