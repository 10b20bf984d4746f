/*
 * aec_b200.h -- C ABI of the B200 AEC device layer.
 *
 * This is the thin boundary the C host code (libaec.h / szlib.h entry points in
 * libaec_b200/csrc/libaec_api.c and sz_api.c) uses to reach the CUDA kernels,
 * and the boundary a foreign-language binding (ctypes, cgo, JNI ...) would bind
 * for the hot path.  Plain pointers and sizes only.
 *
 * What each entry point replaces in the reference (/root/reference/src):
 *   aecb200_encode_host    the work of aec_buffer_encode  (encode.c:950-963) = aec_encode_init
 *                          (:773-907) + aec_encode(AEC_FLUSH) (:909-936) + aec_encode_end (:938-948)
 *   aecb200_decode_host    the work of aec_buffer_decode  (decode.c:843-854)
 *   aecb200_encode_device / aecb200_decode_device
 *                          the same on buffers already resident in HBM (no reference
 *                          counterpart: the reference only knows host pointers)
 *   aecb200_scan_offsets_* RSI boundary discovery; the reference discovers boundaries
 *                          implicitly by decoding sequentially (decode.c:402-421 m_id)
 *
 * Return values are the libaec codes (libaec.h): 0 AEC_OK, -1 AEC_CONF_ERROR,
 * -2 AEC_STREAM_ERROR, -3 AEC_DATA_ERROR, -4 AEC_MEM_ERROR; -100 = CUDA failure
 * (see aecb200_last_error).  There is no CPU fallback: without a usable CUDA
 * device every call fails with -100.
 */
#ifndef AEC_B200_H
#define AEC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AECB200_CUDA_ERROR (-100)

typedef struct aecb200_ctx aecb200_ctx;

/* same four fields as struct aec_stream's coding parameters (libaec.h:84-97) */
typedef struct {
    uint32_t bits_per_sample;
    uint32_t block_size;
    uint32_t rsi;
    uint32_t flags;
} aecb200_params;

/* Carry between consecutive encode launches of one stream (AEC_NO_FLUSH
 * streaming, multi-GPU shards): where the next bit goes, the bits already in
 * the partially filled 32-bit word there, and the split position k of the last
 * coded block (the reference keeps both in struct internal_state:
 * encode.h:128-133 `bits`, :145 `k`). */
typedef struct {
    uint64_t bits;      /* bit offset of the next free bit in the output buffer */
    uint32_t k;
    uint32_t word;      /* big-endian content of the 32-bit word containing `bits` */
} aecb200_carry;

int  aecb200_device_count(void);
/* The calling thread's current CUDA device (-1: none usable) and the device a context is bound to.
 * Every aecb200_* call runs on its context's device and restores the caller's current device. */
int  aecb200_current_device(void);
int  aecb200_ctx_device(aecb200_ctx *ctx);
int  aecb200_set_device(int device);
/* device < 0: current device.  One context = one CUDA stream + workspace;
 * use one context per thread. */
int  aecb200_ctx_create(aecb200_ctx **ctx, int device);
void aecb200_ctx_destroy(aecb200_ctx *ctx);
/* Run all work of this context on an existing CUDA stream (cudaStream_t). */
int  aecb200_ctx_set_stream(aecb200_ctx *ctx, void *cuda_stream);
const char *aecb200_last_error(aecb200_ctx *ctx);
/* Honour AEC_PAD_RSI when encoding (the reference does so only when built with
 * -DENABLE_RSI_PADDING, encode.c:499-505; default off like the stock build). */
void aecb200_ctx_set_encode_padding(aecb200_ctx *ctx, int on);
/* Number of kernels this context has launched so far. */
uint64_t aecb200_ctx_launches(aecb200_ctx *ctx);
/* Host-pointer calls (aecb200_encode_host*, aecb200_decode_host with an offset index; what
 * aec_buffer_encode / aec_buffer_decode of the reference, encode.c:929-963 / decode.c:831-854, turn
 * into) move buffers of at least two pieces as a pipeline: pieces of about `raw_bytes` of samples
 * (whole RSIs) are uploaded, coded and downloaded on three streams so that both PCIe directions
 * and the kernels overlap.  The bytes produced do not depend on the piece size.  Default 16 MiB;
 * 0 = always one piece. */
void aecb200_ctx_set_pipeline_piece(aecb200_ctx *ctx, size_t raw_bytes);

/* Upper bound of the compressed size of in_bytes of input. */
size_t aecb200_encode_bound(const aecb200_params *p, size_t in_bytes);
/* Device workspace the context will hold for an input of in_bytes. */

/* ---- device-resident buffers ------------------------------------------- */

/* Enqueue the encode of d_in[0..in_bytes) into d_out starting at carry->bits.
 * d_out must be 4-byte aligned; d_in should be 16-byte aligned (unaligned input
 * takes a slower bytewise load path).  d_rsi_offsets (optional, device,
 * ceil(samples / (rsi*block_size)) entries) receives the start bit of every RSI.
 * Asynchronous: results are collected by aecb200_encode_finish. */
int aecb200_encode_device(aecb200_ctx *ctx, const aecb200_params *p,
                          const void *d_in, size_t in_bytes,
                          void *d_out, size_t out_cap,
                          const aecb200_carry *carry,
                          uint64_t *d_rsi_offsets);
/* Same, and also records the group index the warp-per-RSI decoder reads
 * (d_grp_index: aecb200_group_index_entries() uint64 entries, 32 per RSI: for
 * each group of ceil(rsi/32) blocks the bit offset of its first CDS and the
 * number of leading blocks that belong to a zero run started earlier).  No
 * reference counterpart. */
int aecb200_encode_device_indexed(aecb200_ctx *ctx, const aecb200_params *p,
                                  const void *d_in, size_t in_bytes,
                                  void *d_out, size_t out_cap,
                                  const aecb200_carry *carry,
                                  uint64_t *d_rsi_offsets, uint64_t *d_grp_index);
size_t aecb200_group_index_entries(const aecb200_params *p, size_t in_bytes);

/* Wait for the last enqueued encode; returns the carry after it (end bit, k;
 * `word` is not filled).  AEC_STREAM_ERROR when the stream did not fit out_cap. */
int aecb200_encode_finish(aecb200_ctx *ctx, aecb200_carry *end);

/* ---- multi-GPU shards: a shard is a contiguous range of whole RSIs
 * coded on its own GPU from carry {0,0,0}.  In shard mode every encode also
 * reports the shard's k clamp pair [klo,khi] (its outgoing k is
 * clamp(k_in, klo, khi), independent of the seed) and a tile index after
 * which k no longer depends on k_in.  After the ranks have exchanged
 * (bits, klo, khi) a rank whose true k_in differs from 0 re-codes its first
 * first_const_tile+1 tiles (aecb200_ctx_set_tile_limit + aecb200_encode_device
 * with the true k in the carry), then moves its stream to its bit offset in the
 * global stream with aecb200_place_bits_device. */
void aecb200_ctx_set_shard_mode(aecb200_ctx *ctx, int on);
/* tail64: the shard's last 64 bits, right-aligned (the next shard completes the word both share) */
int  aecb200_encode_shard_info(aecb200_ctx *ctx, uint32_t *klo, uint32_t *khi, uint64_t *first_const_tile,
                               uint64_t *tail64);
void aecb200_ctx_set_tile_limit(aecb200_ctx *ctx, uint64_t ntiles);
/* d_dst[dst_bit ..) = d_src[0 .. nbits) (bit 0 = MSB of byte 0); whole destination
 * words are written, bits after the range are zero and the dst_bit % 32 bits in
 * front of it are taken from head_or (the predecessor shard's tail).  Asynchronous. */
int  aecb200_place_bits_device(aecb200_ctx *ctx, const void *d_src, uint64_t nbits,
                               void *d_dst, size_t dst_cap, uint64_t dst_bit, uint32_t head_or);

/* The same protocol without the host inside a step (every call below only enqueues):
 *   aecb200_ctx_set_shard_out   every shard-mode encode also writes (bits, klo, khi, tail64) -- four uint64,
 *                               what the ranks all_gather -- to this device address;
 *   aecb200_shard_plan_device   from the gathered 4 x world uint64 a one-thread kernel works out this rank's
 *                               bit offset in the global stream, incoming k, predecessor bits of the shared
 *                               word and how many leading tiles depend on k (d_plan_out, optional: eight
 *                               uint64 = k_in, tiles to code again, bit offset, head bits, total bits, own bits);
 *   aecb200_encode_repair_device  codes those tiles again with the true k (nothing when k_in is 0);
 *   aecb200_place_bits_planned  moves the shard to its bit phase; global != 0: d_dst is the base of the
 *                               whole stream (e.g. a peer GPU's buffer mapped over NVLink) and only the words
 *                               the shard owns are written, so that the placement is the stitch. */
#define AECB200_REPAIR_TILES 64   /* tiles aecb200_encode_repair_device can code again; a plan that asks for more
                                   * (plan word 1) is finished by the caller through the host-driven repair */
void aecb200_ctx_set_shard_out(aecb200_ctx *ctx, void *d_info);
int  aecb200_shard_plan_device(aecb200_ctx *ctx, const void *d_all, int world, int rank, void *d_plan_out);
int  aecb200_encode_repair_device(aecb200_ctx *ctx, const aecb200_params *p, const void *d_in, size_t in_bytes,
                                  void *d_out, size_t out_cap);
int  aecb200_place_bits_planned(aecb200_ctx *ctx, const void *d_src, void *d_dst, size_t dst_cap, int global, int last_rank);

/* Enqueue the decode of out_bytes/bytes_per_sample samples from the stream at
 * d_in using the RSI start offsets d_rsi_offsets[0..nrsi).  Asynchronous. */
int aecb200_decode_device(aecb200_ctx *ctx, const aecb200_params *p,
                          const void *d_in, size_t in_bytes,
                          const uint64_t *d_rsi_offsets, size_t nrsi,
                          void *d_out, size_t out_bytes);
/* Same with the encoder's group index (NULL: it is rebuilt on the device by
 * skimming every RSI from its start offset). */
int aecb200_decode_device_indexed(aecb200_ctx *ctx, const aecb200_params *p,
                                  const void *d_in, size_t in_bytes,
                                  const uint64_t *d_rsi_offsets, size_t nrsi,
                                  const uint64_t *d_grp_index,
                                  void *d_out, size_t out_bytes);
/* Force the lane-per-RSI ("careful") decode kernel for everything (testing). */
void aecb200_ctx_set_careful_decode(aecb200_ctx *ctx, int on);
/* RSIs the fast kernel handed to the careful kernel in the last finished decode (diagnostics). */
uint64_t aecb200_ctx_last_handover(aecb200_ctx *ctx);
/* Wait for the last enqueued decode; *out_written = bytes of samples delivered. */
int aecb200_decode_finish(aecb200_ctx *ctx, size_t *out_written);

/* Discover the RSI start offsets of a stream without an index (the only kind of stream a caller of
 * libaec.h 0.3.4 can hand over; the reference finds them by decoding sequentially, decode.c:402-421).
 * Streams of at least 2 KiB are skimmed in parallel: per-bit-position CDS tables, pointer doubling,
 * one table look-up per RSI (aec_skim.cu); shorter ones by a single thread.
 * d_rsi_offsets must hold max_rsi entries.  Synchronous. */
int aecb200_scan_offsets_device(aecb200_ctx *ctx, const aecb200_params *p,
                                const void *d_in, size_t in_bytes, uint64_t start_bit,
                                uint64_t *d_rsi_offsets, size_t max_rsi, size_t *found);

/* mode 0: choose by stream size, 1: always the one-thread scan, 2: always the parallel tables;
 * window_bits (0 = keep): stream bits whose tables are held at a time (default 2^25). */
void aecb200_ctx_set_scan_mode(aecb200_ctx *ctx, int mode, uint64_t window_bits);
/* RSIs of the last scan whose length came from the tables (the rest were skimmed serially). */
uint64_t aecb200_ctx_last_scan_fast(aecb200_ctx *ctx);
/* AEC_NO_FLUSH decoding accumulates the stream on the device: after aecb200_ctx_accumulate_next(ctx, b)
 * the next aecb200_decode_host_resume call treats in[0] as byte b (a multiple of 4) of one stream whose
 * earlier bytes, uploaded by earlier such calls, are still in HBM, and uploads only what is new.
 * b < 0 starts a new stream.  aecb200_ctx_accumulated_uploads: bytes uploaded that way so far. */
void aecb200_ctx_accumulate_next(aecb200_ctx *ctx, long long stream_byte0);
uint64_t aecb200_ctx_accumulated_uploads(aecb200_ctx *ctx);
/* RSI start offsets the last aecb200_decode_host / _resume call without an index discovered (bits from
 * in[0]); returns their number, copies at most cap of them. */
size_t aecb200_ctx_found_offsets(aecb200_ctx *ctx, uint64_t *dst, size_t cap);

/* ---- host buffers (what the libaec.h entry points call) ------------------ */

/* Whole-buffer encode: stage to HBM, encode, copy back.  *out_len = bytes
 * produced (<= out_cap), *in_consumed = bytes of whole samples consumed.
 * rsi_offsets (optional, host) receives up to offsets_cap RSI start bits. */
int aecb200_encode_host(aecb200_ctx *ctx, const aecb200_params *p,
                        const void *in, size_t in_bytes,
                        void *out, size_t out_cap, size_t *out_len, size_t *in_consumed,
                        uint64_t *rsi_offsets, size_t offsets_cap, size_t *n_offsets);
/* Streaming piece: codes the whole RSIs in `in` (everything, including a short
 * last RSI and the final byte padding, when `final`), continuing the stream
 * described by *carry (bits in 0..7: how many bits of the stream's last byte
 * are already in use; word: that byte in the top 8 bits; k).  out[0] is the
 * byte that contains the carried bits.  *out_len = complete bytes produced
 * (all bytes when final); *carry is updated for the next piece.  RSI offsets
 * are relative to bit 0 of out[0]. */
int aecb200_encode_host_piece(aecb200_ctx *ctx, const aecb200_params *p,
                              const void *in, size_t in_bytes, int final,
                              void *out, size_t out_cap, size_t *out_len, size_t *in_consumed,
                              aecb200_carry *carry,
                              uint64_t *rsi_offsets, size_t offsets_cap, size_t *n_offsets);

/* AEC_NO_FLUSH encoding accumulates on the device: aecb200_ctx_stage_input appends bytes that do not yet
 * make a whole RSI to the context's input stage (offset = bytes staged so far); aecb200_encode_host_piece
 * with in == NULL then codes in_bytes staged bytes without another upload. */
int aecb200_ctx_stage_input(aecb200_ctx *ctx, size_t offset, const void *src, size_t n);
uint64_t aecb200_ctx_staged_uploads(aecb200_ctx *ctx);

/* Whole-buffer decode.  rsi_offsets == NULL: discover RSI boundaries on the
 * device first (sequential, slow).  *out_len = bytes delivered. */
int aecb200_decode_host(aecb200_ctx *ctx, const aecb200_params *p,
                        const void *in, size_t in_bytes,
                        const uint64_t *rsi_offsets, size_t n_offsets,
                        void *out, size_t out_cap, size_t *out_len);

/* Streaming form of the above: decode starts at `start_bit`, which must be the
 * first bit of an RSI, and drops the first skip_samples samples of that RSI
 * (already delivered earlier).  On return *resume_bit / *resume_delivered tell
 * where the next call has to start: the first bit of the RSI holding the next
 * undelivered sample and how many of its samples were delivered so far. */
int aecb200_decode_host_resume(aecb200_ctx *ctx, const aecb200_params *p,
                               const void *in, size_t in_bytes,
                               const uint64_t *rsi_offsets, size_t n_offsets,
                               uint64_t start_bit, size_t skip_samples,
                               void *out, size_t out_cap, size_t *out_len,
                               uint64_t *resume_bit, size_t *resume_delivered);

/* ---- SZIP shim on the device (what libsz's SZ_BufftoBuffCompress / SZ_BufftoBuffDecompress call;
 * reference: sz_compat.c:110-268).  The byte-plane (de)interleave of 32/64-bit pixels and the scanline
 * padding run as kernels between the copies and the coder instead of host loops.  Return values are
 * the szlib.h codes (SZ_OK 0, SZ_OUTBUFF_FULL 2, SZ_PARAM_ERROR -1, SZ_MEM_ERROR -4) or an AEC_* /
 * AECB200_CUDA_ERROR code from the coder.  *dest_len: in = capacity, out = bytes produced. */
int aecb200_sz_compress_host(aecb200_ctx *ctx, int options_mask, int bits_per_pixel, int pixels_per_block,
                             int pixels_per_scanline, const void *source, size_t source_len, void *dest, size_t *dest_len);
int aecb200_sz_decompress_host(aecb200_ctx *ctx, int options_mask, int bits_per_pixel, int pixels_per_block,
                               int pixels_per_scanline, const void *source, size_t source_len, void *dest, size_t *dest_len);
/* Many chunks in flight (HDF5 chunk pipelines): chunk i goes through its own pooled context and CUDA
 * stream on one of `threads` host threads (0 = a default), so uploads, kernels and downloads of different
 * chunks overlap.  status[i] receives what the single call would have returned; the function returns the
 * first non-zero status (0 when every chunk succeeded). */
int aecb200_sz_compress_batch(int n, void *const *dest, size_t *dest_len, const void *const *source, const size_t *source_len,
                              int options_mask, int bits_per_pixel, int pixels_per_block, int pixels_per_scanline,
                              int *status, int threads);
int aecb200_sz_decompress_batch(int n, void *const *dest, size_t *dest_len, const void *const *source, const size_t *source_len,
                                int options_mask, int bits_per_pixel, int pixels_per_block, int pixels_per_scanline,
                                int *status, int threads);
/* contexts (stream + workspace) of the calling thread's current device, kept between calls */
aecb200_ctx *aecb200_pool_get(void);
void aecb200_pool_put(aecb200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* AEC_B200_H */
