/*
 * aec_sz.cu -- the byte shuffles of the SZIP shim on the device.
 *
 * /root/reference/src/sz_compat.c:39-108 turns 32/64-bit pixels into byte planes (interleave_buffer /
 * deinterleave_buffer) and fills every scanline up to a whole number of blocks (add_padding /
 * remove_padding) with host loops around aec_buffer_encode / aec_buffer_decode.  Here the caller's bytes
 * go to HBM as they are and one kernel writes the coder's input (planes + padded scanlines) from them;
 * after decoding one kernel undoes both on the way to the buffer that is copied back.  One thread per
 * destination byte: these passes are a few percent of the PCIe time of a chunk.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "aec_device.h"

namespace {

/* byte t of the plane-ordered buffer (dest[j * nw + w] = src[w * ws + j], sz_compat.c:39-53) */
__device__ __forceinline__ uint8_t plane_byte(const uint8_t *src, uint64_t t, uint64_t nw, uint32_t ws)
{
    if (ws == 1u) return src[t];
    if (t >= nw * ws) return 0u;                        /* bytes behind the last whole word: unspecified in the reference */
    const uint64_t j = t / nw, w = t - j * nw;
    return src[w * ws + j];
}

__global__ void aec_sz_pack_kernel(const uint8_t *src, uint64_t src_len, uint8_t *dst, uint64_t padded_len,
                                   uint32_t ws, uint64_t line, uint64_t full_line, uint32_t px, uint32_t nn)
{
    const uint64_t nw = ws > 1u ? src_len / ws : 0;
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < padded_len; o += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t l = o / full_line, x = o - l * full_line;
        const uint64_t start = l * line;
        uint64_t avail = src_len > start ? src_len - start : 0;
        if (avail > line) avail = line;
        uint8_t v = 0;
        if (x < avail) v = plane_byte(src, start + x, nw, ws);
        else if (nn && avail >= px) v = plane_byte(src, start + avail - px + (x - avail) % px, nw, ws);   /* last pixel again (sz_compat.c:71-94) */
        dst[o] = v;
    }
}

/* decoded samples (padded scanlines, plane order) -> the caller's layout; n = bytes wanted */
__global__ void aec_sz_unpack_kernel(const uint8_t *src, uint8_t *dst, uint64_t n, uint32_t ws, uint64_t line, uint64_t full_line)
{
    const uint64_t nw = ws > 1u ? n / ws : 0;
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t t = o;                                 /* index in the squeezed, plane-ordered buffer */
        if (ws > 1u) {
            if (o >= nw * ws) continue;                 /* sz_compat.c:55-69 leaves these untouched */
            const uint64_t w = o / ws, j = o - w * ws;
            t = j * nw + w;
        }
        const uint64_t l = t / line, x = t - l * line;  /* remove_padding, sz_compat.c:96-108 */
        dst[o] = src[l * full_line + x];
    }
}

} // namespace

cudaError_t aec_sz_pack_launch(const uint8_t *src, uint64_t src_len, uint8_t *dst, uint64_t padded_len, uint32_t ws,
                               uint64_t line, uint64_t full_line, uint32_t px, uint32_t nn, cudaStream_t st)
{
    if (padded_len == 0) return cudaSuccess;
    const uint64_t blocks = (padded_len + 255) / 256;
    aec_sz_pack_kernel<<<(unsigned)(blocks > 148 * 32 ? 148 * 32 : blocks), 256, 0, st>>>(src, src_len, dst, padded_len, ws, line,
                                                                                         full_line, px, nn);
    return cudaGetLastError();
}

cudaError_t aec_sz_unpack_launch(const uint8_t *src, uint8_t *dst, uint64_t n, uint32_t ws, uint64_t line, uint64_t full_line,
                                 cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const uint64_t blocks = (n + 255) / 256;
    aec_sz_unpack_kernel<<<(unsigned)(blocks > 148 * 32 ? 148 * 32 : blocks), 256, 0, st>>>(src, dst, n, ws, line, full_line);
    return cudaGetLastError();
}
