/*
 * aec_decode.cu -- RSI-parallel CCSDS 121.0-B-2 decoder for sm_100a.
 *
 * Replaces the reference's decoder state machine
 * (/root/reference/src/decode.c:402-677 m_id .. m_uncomp and the FLUSH/put_*
 * post-processing at :67-189) for streams whose RSI start offsets are known
 * (from our encoder's index, or from aec_scan_offsets below):
 *
 *   lane  = one RSI.  It walks its RSI block by block with a position-addressed
 *           bit reader (aec_decode_core.cuh), undoes the predictor (a true
 *           recurrence, so it stays inside the lane) and leaves the block's
 *           samples in a padded shared-memory row;
 *   warp  = 32 consecutive RSIs in lock-step over the block index; after every
 *           block the warp stores its 32 rows cooperatively with 32-bit words
 *           in the sample layout the caller asked for (decode.c:144-189).
 *
 * aec_scan_offsets is the slow path for foreign streams that carry no index:
 * one thread skims CDS after CDS (ids, FS terminators by popcount) and records
 * where every RSI starts.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "aec_decode_core.cuh"
#include "aec_device.h"

namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr int DEC_WARPS = 4;

/* Four consecutive samples of a row -> the 32-bit words of their storage bytes. */
template <int B>
__device__ __forceinline__ void store_group(uint8_t *out, uint64_t sample_idx, const uint32_t *s, uint32_t msb)
{
    /* sample_idx is a multiple of 4 / B' such that the byte address is 4-byte aligned */
    if (B == 4) {
        uint32_t v = s[0];
        reinterpret_cast<uint32_t *>(out)[sample_idx] = msb ? __byte_perm(v, 0, 0x0123) : v;
    } else if (B == 2) {
        uint32_t a = s[0] & 0xFFFFu, b = s[1] & 0xFFFFu;
        uint32_t v = msb ? (__byte_perm(a, b, 0x4501)) : (a | (b << 16));
        reinterpret_cast<uint32_t *>(out)[sample_idx >> 1] = v;
    } else if (B == 1) {
        uint32_t v = (s[0] & 0xFFu) | ((s[1] & 0xFFu) << 8) | ((s[2] & 0xFFu) << 16) | (s[3] << 24);
        reinterpret_cast<uint32_t *>(out)[sample_idx >> 2] = v;
    } else {
        uint32_t a = s[0] & 0xFFFFFFu, b = s[1] & 0xFFFFFFu, cc = s[2] & 0xFFFFFFu, d = s[3] & 0xFFFFFFu;
        if (msb) {
            a = __byte_perm(a, 0, 0x4012); b = __byte_perm(b, 0, 0x4012);
            cc = __byte_perm(cc, 0, 0x4012); d = __byte_perm(d, 0, 0x4012);
        }
        uint32_t *o = reinterpret_cast<uint32_t *>(out) + (sample_idx >> 2) * 3;
        o[0] = a | (b << 24);
        o[1] = (b >> 8) | (cc << 16);
        o[2] = (cc >> 16) | (d << 8);
    }
}

template <int JT, int B>
__global__ void __launch_bounds__(DEC_WARPS * 32)
aec_decode_kernel(const AecDecArgs a)
{
    const AecCfg &c = a.cfg;
    const uint32_t J = JT ? (uint32_t)JT : c.J;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t stride = J + 1;                       /* padded row, conflict free */
    extern __shared__ uint32_t rows[];
    uint32_t *wrows = rows + (size_t)warp * 32u * stride;
    uint32_t *row = wrows + (size_t)lane * stride;

    const uint64_t rsi0 = ((uint64_t)blockIdx.x * DEC_WARPS + warp) * 32ull;   /* first RSI of this warp */
    if (rsi0 >= a.nrsi) return;
    const uint64_t r = rsi0 + lane;
    const bool have = r < a.nrsi;

    /* how many samples this RSI has to deliver */
    uint64_t limit = 0;
    if (have) {
        uint64_t startS = r * (uint64_t)c.R;
        limit = a.out_samples > startS ? a.out_samples - startS : 0;
        if (limit > c.R) limit = c.R;
    }
    BitRd br;
    br.init(a.in_words, (a.in_bytes + 3) >> 2, a.in_bytes * 8ull);
    RsiDec st; st.pos = have ? a.rsi_offsets[r] : 0; st.zero_left = 0; st.status = DEC_OK;
    uint32_t u_prev = 0;
    uint32_t delivered = 0;
    bool active = have && limit > 0;

    const uint32_t nblocks = c.rsi;
    for (uint32_t b = 0; b < nblocks; b++) {
        uint32_t cnt = 0;
        if (active) {
            cnt = aec_decode_block<JT>(c, br, st, b, row);
            uint64_t room = limit - delivered;
            if (cnt > room) cnt = (uint32_t)room;
            aec_unmap_row(c, row, cnt, (c.pp && b == 0) ? 1u : 0u, &u_prev);
        }
        __syncwarp();
        /* ---- cooperative store of the 32 rows of this block index ---- */
        const uint32_t fullmask = __ballot_sync(FULL, cnt == J);
        const uint32_t partmask = __ballot_sync(FULL, cnt != J && cnt != 0);
        if (JT != 0 && a.out_aligned && partmask == 0) {
            /* rows are full or empty: flat index over 32*J samples, one 32-bit store per lane-step */
            constexpr int SPG = (B == 4) ? 1 : ((B == 2) ? 2 : 4);       /* samples per 32-bit group */
            constexpr int GPR = (JT ? JT : 4) / SPG;                     /* groups per row */
#pragma unroll 4
            for (int g = (int)lane; g < 32 * GPR; g += 32) {
                int rowi = g / GPR, gi = g % GPR;
                if (!((fullmask >> rowi) & 1u)) continue;
                const uint32_t *src = wrows + (size_t)rowi * stride + gi * SPG;
                uint32_t s[4] = {src[0], SPG > 1 ? src[1] : 0u, SPG > 2 ? src[2] : 0u, SPG > 2 ? src[3] : 0u};
                uint64_t sidx = (rsi0 + rowi) * (uint64_t)c.R + (uint64_t)b * J + (uint64_t)gi * SPG;
                store_group<B>(a.out, sidx, s, c.msb);
            }
        } else if (cnt) {
            /* generic: each lane stores its own row bytewise */
            uint64_t sidx = r * (uint64_t)c.R + (uint64_t)b * J;
            for (uint32_t i = 0; i < cnt; i++)
                aec_store_sample(a.out + (sidx + i) * c.B, row[i], c.B, c.msb);
        }
        __syncwarp();
        delivered += cnt;
        if (active && (cnt < J || delivered >= limit)) active = false;
        if (!__any_sync(FULL, active)) break;
    }
    if (have) {
        if (a.rsi_count) a.rsi_count[r] = delivered;
        if (delivered < limit)
            atomicMax(reinterpret_cast<unsigned long long *>(&a.result[0]),
                      ~(unsigned long long)(r * (uint64_t)c.R + delivered));
        if (st.status == DEC_ERROR)
            atomicOr(reinterpret_cast<unsigned long long *>(&a.result[1]), 1ull);
    }
}

__global__ void aec_scan_offsets_kernel(const AecCfg c, const uint32_t *in_words, uint64_t in_bytes,
                                        uint64_t start_bit, uint64_t *offsets, uint64_t max_rsi,
                                        uint64_t *result)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    BitRd br;
    br.init(in_words, (in_bytes + 3) >> 2, in_bytes * 8ull);
    RsiDec st; st.pos = start_bit; st.zero_left = 0; st.status = DEC_OK;
    uint64_t found = 0;
    for (uint64_t r = 0; r < max_rsi; r++) {
        if (c.pad) st.pos = (st.pos + 7ull) & ~7ull;
        uint64_t start = st.pos;
        st.zero_left = 0;
        if (start >= br.nbits) break;
        offsets[found++] = start;          /* even a truncated RSI may still deliver leading samples */
        for (uint32_t b = 0; b < c.rsi; b++)
            if (!aec_skim_block(c, br, st, b)) break;
        if (st.status != DEC_OK) break;
    }
    result[0] = found;
    result[1] = (st.status == DEC_ERROR) ? 1ull : 0ull;
    result[2] = st.pos;
}

template <int JT, int B>
cudaError_t launch_dec(const AecDecArgs &a, cudaStream_t st)
{
    auto kern = aec_decode_kernel<JT, B>;
    uint32_t J = JT ? (uint32_t)JT : a.cfg.J;
    uint32_t smem = DEC_WARPS * 32u * (J + 1u) * 4u;
    static uint32_t attr = 48 * 1024;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = smem;
    }
    uint64_t nwarps = (a.nrsi + 31) / 32;
    uint64_t grid = (nwarps + DEC_WARPS - 1) / DEC_WARPS;
    if (grid == 0) return cudaSuccess;
    kern<<<(unsigned)grid, DEC_WARPS * 32, smem, st>>>(a);
    return cudaGetLastError();
}

template <int JT>
cudaError_t launch_dec_j(const AecDecArgs &a, cudaStream_t st)
{
    switch (a.cfg.B) {
    case 1: return launch_dec<JT, 1>(a, st);
    case 2: return launch_dec<JT, 2>(a, st);
    case 3: return launch_dec<JT, 3>(a, st);
    default: return launch_dec<JT, 4>(a, st);
    }
}

} // namespace

cudaError_t aec_decode_launch(const AecDecArgs &a, int num_sms, cudaStream_t st)
{
    (void)num_sms;
    switch (a.cfg.J) {
    case 8:  return launch_dec_j<8>(a, st);
    case 16: return launch_dec_j<16>(a, st);
    case 32: return launch_dec_j<32>(a, st);
    case 64: return launch_dec_j<64>(a, st);
    default: return launch_dec_j<0>(a, st);
    }
}

cudaError_t aec_scan_offsets_launch(const AecCfg &c, const uint32_t *in_words, uint64_t in_bytes,
                                    uint64_t start_bit, uint64_t *offsets, uint64_t max_rsi,
                                    uint64_t *result, cudaStream_t st)
{
    aec_scan_offsets_kernel<<<1, 32, 0, st>>>(c, in_words, in_bytes, start_bit, offsets, max_rsi, result);
    return cudaGetLastError();
}
