/*
 * aec_device.h -- internal interface between the host runtime (aec_runtime.cu)
 * and the CUDA kernels.  Not installed; the public C ABI is include/aec_b200.h.
 */
#ifndef AEC_DEVICE_H
#define AEC_DEVICE_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>

#include "aec_core.cuh"

/* Opt-in to more than the default dynamic shared memory of a kernel.  cudaFuncSetAttribute applies
 * to the current device only, so the size already granted is remembered per device (one instance of
 * this struct per kernel instantiation); raising it is serialised by a process-wide mutex because
 * contexts of several host threads share the kernels. */
struct AecSmemOptIn {
    static constexpr int MAXDEV = 64;
    std::atomic<uint32_t> have[MAXDEV];            /* zero-initialised (static storage): nothing granted yet */
    template <class K>
    cudaError_t ensure(K kern, uint32_t bytes, uint32_t dflt)
    {
        if (bytes <= dflt) return cudaSuccess;
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        const bool tracked = dev >= 0 && dev < MAXDEV;
        if (tracked && bytes <= have[dev].load(std::memory_order_acquire)) return cudaSuccess;
        static std::mutex mu;
        std::lock_guard<std::mutex> lock(mu);
        if (tracked && bytes <= have[dev].load(std::memory_order_relaxed)) return cudaSuccess;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e == cudaSuccess && tracked) have[dev].store(bytes, std::memory_order_release);
        return e;
    }
};

/* Arguments of one encode launch (passed by value). */
struct AecEncArgs {
    AecCfg cfg;
    const uint8_t *in;          /* raw samples, device */
    uint64_t nsamples;          /* whole samples in `in` */
    uint64_t nrsi;              /* RSIs to code (last may be short) */
    uint32_t last_nblk;         /* coded blocks of the last RSI */
    uint32_t RP;                /* block slots per RSI (power of two <= TB, or multiple of TB) */
    uint64_t ntiles;            /* tiles the main kernel codes (a prefix of the shard when repairing) */
    uint64_t ntiles_total;      /* tiles of the whole launch geometry (fix-up covers all of them) */
    uint32_t aligned;           /* `in` is 16-byte aligned */
    uint32_t staging_words;     /* dynamic shared memory in words */
    uint32_t *out_words;        /* output stream, 4-byte aligned, device */
    uint64_t out_cap_words;     /* floor(capacity / 4) */
    uint64_t out_cap_bytes;
    uint64_t seed_bits;         /* bit offset in out where this launch starts */
    uint32_t seed_k;            /* k carried in from the stream so far */
    uint32_t seed_word;         /* content of the partial word at seed_bits (bits already there) */
    /* workspace, device */
    uint64_t *desc;             /* [ntiles], zeroed: per-tile aggregates, written by the worker CTAs */
    uint64_t *pref;             /* [ntiles], zeroed: per-tile exclusive prefixes, written by the scanner */
    uint32_t *ticket;           /* zeroed */
    uint32_t *head_c, *tail_c;  /* [ntiles] partial boundary words */
    uint64_t *tile_end;         /* [ntiles] absolute end bit of each tile */
    uint32_t *tile_kagg;        /* [ntiles] clamp pair (lo | hi<<8) of each tile, independent of the seed */
    uint64_t *rsi_offsets;      /* optional [nrsi] absolute start bit of each RSI */
    uint64_t *grp_index;        /* optional [nrsi*32] group index for the warp-per-RSI decoder (see AecDecArgs) */
    uint32_t grp_G;             /* blocks per group = ceil(rsi / 32) */
    /* filled by aec_encode_launch */
    uint32_t tpr;               /* tiles per RSI when RP >= TB, else 0 */
    uint32_t rp_shift;          /* log2(RP) when RP < TB */
    uint32_t tile_rsi_shift;    /* log2(TB / RP) when RP < TB */
    uint32_t grp_magic;         /* floor(2^32 / grp_G) + 1: b / grp_G == umulhi(b, grp_magic) for b < 4096; 0 when grp_G == 1 */
    uint64_t *result;           /* [0] end bit, [1] k after the last block, [2..5] shard summary (lo, hi, first constant tile, last 64 bits) */
    /* multi-GPU shards, everything decided on the device (no host round trip inside a step): */
    uint64_t *shard_out;        /* optional [4]: (bits, klo, khi, last 64 bits) of this shard, what the ranks all_gather */
    const uint64_t *dyn;        /* optional (k repair planned on the device): [0] incoming k, replaces seed_k; [1] leading tiles
                                 * that have to be coded again: the launch does nothing unless dyn_lo < dyn[1] <= dyn_hi */
    uint64_t dyn_lo, dyn_hi;
};

/* what aec_shard_plan_kernel leaves for the repair and placement launches of a shard */
enum { PLAN_K_IN = 0, PLAN_REPAIR_TILES = 1, PLAN_BIT_OFFSET = 2, PLAN_HEAD_OR = 3, PLAN_TOTAL_BITS = 4, PLAN_MY_BITS = 5,
       PLAN_WORDS = 8 };

/* Arguments of one decode launch. */
struct AecDecArgs {
    AecCfg cfg;
    const uint32_t *in_words;   /* compressed stream, 4-byte aligned, device */
    uint64_t in_bytes;
    const uint64_t *rsi_offsets;/* [nrsi] start bit of each RSI */
    uint64_t nrsi;
    uint8_t *out;               /* decoded samples, device */
    uint64_t out_samples;       /* samples wanted in total */
    uint32_t out_aligned;       /* out is 16-byte aligned */
    uint64_t *result;           /* [0] ~(first sample position that fell short) or 0, [1] status flags */
    uint32_t *rsi_count;        /* [nrsi] samples each RSI delivered */
    /* Group index: entry [r*32 + l] describes the G = ceil(rsi/32) blocks lane l of the
     * warp that decodes RSI r owns: bits 55..0 = absolute bit offset of the first CDS
     * to parse, bits 63..56 = leading blocks that still belong to a zero run started
     * in an earlier group (they are zero and have no CDS of their own). */
    const uint64_t *grp_index;
    uint32_t grp_G;
    /* careful (lane-per-RSI) kernel only: decode the RSIs listed here instead of 0..nrsi */
    uint32_t *rsi_list;         /* [nrsi] */
    uint32_t *rsi_list_count;
};

/* Arguments of the parallel RSI-boundary discovery over one window of the stream (aec_skim.cu). */
struct AecSkimArgs {
    AecCfg cfg;                 /* cfg.pad: RSIs start on byte boundaries (the decoder always honours AEC_PAD_RSI) */
    const uint32_t *in_words;   /* compressed stream, 4-byte aligned, device */
    uint64_t nbits;             /* stream bits (in_bytes * 8) */
    uint64_t wb;                /* first bit of the window (multiple of 32) */
    uint32_t np;                /* bit positions of the window that get table entries */
    uint32_t nh_eff;            /* RSIs that start before wb + nh_eff are walked in this window */
    uint32_t last;              /* last window of the stream */
    uint32_t LV;                /* levels of the CDS chain tables (level j = 2^j CDSs) */
    uint32_t la_words;          /* filled by the launcher: words staged beyond a tile */
    uint32_t bulk;              /* filled by the launcher: tiles can be staged with cp.async.bulk (16-byte aligned) */
    uint32_t *T;                /* [LV][np] */
    uint32_t *H;                /* [np] RSI lengths */
    uint32_t *R;                /* [np] entries of a first CDS of an RSI (the one with the reference sample) */
    uint32_t *H8;               /* optional: two buffers of [np] behind each other; the second ends up holding the length of
                                 * eight RSIs in a row (streams of many short RSIs: the walk then takes an eighth of the steps) */
    uint32_t sparse;            /* RSI lengths for marked chain ends only (LV >= 4; state[5] turns it off on the device) */
    uint32_t set;               /* table set of this window (0/1): which list counter it uses */
    uint32_t *cand_list;        /* optional [cand_cap]: positions that have an RSI length (what the long-jump passes run over) */
    uint32_t cand_cap;
    uint64_t *grp_index;        /* optional [max_rsi * 32]: group index of the RSIs found (AecDecArgs::grp_index) */
    uint32_t grp_G;             /* blocks per group = ceil(rsi / 32) */
    uint64_t *state;            /* [0] next RSI bit, [1] RSIs found, [2] flags (1 ended, 2 data error), [3] RSIs taken from the tables,
                                 * [4] RSIs found before the current window's walk, [5] bit 0: dense tables from now on,
                                 * [6], [7] per table set: candidates listed; bit 63: this window's tables are dense */
    uint64_t *offsets;          /* [max_rsi] */
    uint64_t max_rsi;
};
uint32_t aec_skim_levels(const AecCfg &c);
uint32_t aec_skim_sparse_min_levels(void);  /* sparse candidates need this many levels */
uint64_t aec_skim_margin_bits(const AecCfg &c);
/* level-0 tables, doubling and RSI lengths of one window; then the walk through it */
cudaError_t aec_skim_window_launch(const AecSkimArgs &a, cudaStream_t st);
cudaError_t aec_skim_walk_launch(const AecSkimArgs &a, cudaStream_t st);

/* SZIP shim on the device (aec_sz.cu): caller's bytes -> byte planes + padded scanlines, and back */
cudaError_t aec_sz_pack_launch(const uint8_t *src, uint64_t src_len, uint8_t *dst, uint64_t padded_len, uint32_t ws,
                               uint64_t line, uint64_t full_line, uint32_t px, uint32_t nn, cudaStream_t st);
cudaError_t aec_sz_unpack_launch(const uint8_t *src, uint8_t *dst, uint64_t n, uint32_t ws, uint64_t line, uint64_t full_line,
                                 cudaStream_t st);

uint32_t aec_encode_tile_blocks(uint32_t J);
uint32_t aec_encode_staging_words(const AecCfg &c);
cudaError_t aec_encode_launch(const AecEncArgs &a, int num_sms, cudaStream_t st);
/* result[2..4] = clamp pair of the whole launch and the first tile after which k no longer depends on the seed */
cudaError_t aec_encode_summary_launch(const AecEncArgs &a, cudaStream_t st);
/* copy nbits bits from src (bit 0 = MSB of word 0) to dst starting at bit dst_bit; dst words are private to the caller */
cudaError_t aec_place_bits_launch(const uint32_t *src, uint64_t nbits, uint32_t *dst, uint64_t dst_bit, uint64_t dst_cap_words, uint32_t head_or, cudaStream_t st);
/* plan[PLAN_*] from the gathered (bits, klo, khi, tail64) of all shards; result[4] = this shard's first constant tile */
cudaError_t aec_shard_plan_launch(const uint64_t *all, uint32_t world, uint32_t rank, const uint64_t *result, uint64_t *plan,
                                  uint64_t *plan_copy, cudaStream_t st);
/* placement with bit offset, length and head bits read from the plan; global != 0: dst is the base of the whole
 * stream (possibly a peer GPU's buffer) and only the words this shard owns are written, else dst[0] is the word
 * that holds the shard's first bit */
cudaError_t aec_place_bits_planned_launch(const uint32_t *src, const uint64_t *plan, uint32_t *dst, uint64_t dst_cap_words,
                                          uint32_t global, uint32_t last_rank, int num_sms, cudaStream_t st);

cudaError_t aec_decode_launch(const AecDecArgs &a, int num_sms, cudaStream_t st);
uint32_t aec_decode_group_blocks(const AecCfg &c);     /* G = ceil(rsi / 32) */
uint32_t aec_decode_warp_warps(const AecCfg &c);       /* 0 when the fast kernel cannot be used */
/* fast path: one warp per RSI from the group index; RSIs it cannot finish are appended to rsi_list */
cudaError_t aec_decode_warp_launch(const AecDecArgs &a, int num_sms, cudaStream_t st);
/* build the group index of RSIs whose start offsets are known (one lane skims one RSI); only_missing: leave RSIs
 * alone whose first entry is not SK_GRP_MISSING (the boundary discovery has written theirs from its tables) */
cudaError_t aec_build_group_index_launch(const AecDecArgs &a, uint64_t *grp_index, cudaStream_t st, int only_missing = 0);
/* Sequential RSI-boundary scan for streams without an offset index:
 * fills offsets[0..max_rsi) and result[0] = RSIs found, result[1] = status. */
cudaError_t aec_scan_offsets_launch(const AecCfg &c, const uint32_t *in_words, uint64_t in_bytes,
                                    uint64_t start_bit, uint64_t *offsets, uint64_t max_rsi,
                                    uint64_t *result, cudaStream_t st);

#endif
