/*
 * aec_oracle.h -- CPU oracle for the CCSDS 121.0-B-2 adaptive entropy coder.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C, whole-buffer restatement of
 * what reference libaec 0.3.4 computes (src/encode.c, src/decode.c,
 * src/encode_accessors.c, src/sz_compat.c).  It exists so the CUDA path can be
 * checked bit for bit.  Nothing under libaec_b200/ may include, link or load
 * it; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs do.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks it against
 *   - the compiled, unmodified reference (oracle/_ref/libaec_ref*.so) on
 *     randomized parameter sets (differential),
 *   - the reference's golden vector data/typical.rz in both directions,
 *   - committed fixtures in tests/golden/ generated from the reference.
 */
#ifndef AEC_ORACLE_H
#define AEC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

/* Same numeric values as the reference's public header (src/libaec.h:102-133). */
#define ORC_DATA_SIGNED     1
#define ORC_DATA_3BYTE      2
#define ORC_DATA_MSB        4
#define ORC_DATA_PREPROCESS 8
#define ORC_RESTRICTED      16
#define ORC_PAD_RSI         32
#define ORC_NOT_ENFORCE     64

#define ORC_OK            0
#define ORC_CONF_ERROR   (-1)
#define ORC_STREAM_ERROR (-2)
#define ORC_DATA_ERROR   (-3)
#define ORC_MEM_ERROR    (-4)

typedef struct {
    uint32_t bits_per_sample;
    uint32_t block_size;
    uint32_t rsi;
    uint32_t flags;
} orc_params;

/* Per-block record produced by the encoder when a trace buffer is given
 * (used by the tests to compare option choices, not only bytes). */
typedef struct {
    uint8_t  option;     /* 0 zero-run member, 1 SE, 2 split, 3 uncompressed */
    uint8_t  k;          /* split position when option == 2 */
    uint8_t  klo, khi;   /* argmin plateau of the split length */
    uint32_t cds_bits;   /* bits this block contributed (0 for non-owner zero blocks) */
} orc_block_trace;

/*
 * Whole-buffer encode == aec_buffer_encode (encode.c:950-963).
 *   honour_pad_rsi: 0 mimics the default build (AEC_PAD_RSI ignored on encode),
 *                   1 mimics -DENABLE_RSI_PADDING (encode.c:499-505).
 *   rsi_bit_offsets/offsets_cap: optional; receives the bit offset at which
 *                   each RSI starts (no reference counterpart: SURVEY D1).
 *   in_consumed:    bytes of input consumed (whole samples only, encode.c:673).
 * Returns ORC_OK, ORC_CONF_ERROR or ORC_STREAM_ERROR (output too small; out_len
 * then holds the bytes that fitted).
 */
int orc_encode(const orc_params *p, int honour_pad_rsi,
               const uint8_t *in, size_t in_len,
               uint8_t *out, size_t out_cap, size_t *out_len,
               size_t *in_consumed,
               uint64_t *rsi_bit_offsets, size_t offsets_cap, size_t *n_offsets,
               orc_block_trace *trace, size_t trace_cap, size_t *n_trace);

/*
 * Whole-buffer decode == aec_buffer_decode (decode.c:843-854).
 * Decodes until out_cap/bytes_per_sample samples are produced or the input is
 * exhausted.  Returns ORC_OK, ORC_CONF_ERROR, ORC_DATA_ERROR or ORC_MEM_ERROR.
 */
int orc_decode(const orc_params *p,
               const uint8_t *in, size_t in_len,
               uint8_t *out, size_t out_cap, size_t *out_len);

/* SZIP shim == SZ_BufftoBuffCompress / SZ_BufftoBuffDecompress
 * (sz_compat.c:110-183, :185-268).  Return codes as szlib.h:14-19. */
int orc_sz_compress(void *dest, size_t *dest_len, const void *src, size_t src_len,
                    int options_mask, int bits_per_pixel,
                    int pixels_per_block, int pixels_per_scanline);
int orc_sz_decompress(void *dest, size_t *dest_len, const void *src, size_t src_len,
                      int options_mask, int bits_per_pixel,
                      int pixels_per_block, int pixels_per_scanline);

#endif
