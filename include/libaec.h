/*
 * libaec.h -- public stream API of the B200 AEC library.
 *
 * ABI-compatible with the header of reference libaec 0.3.4
 * (/root/reference/src/libaec.h:67-166): same struct layout, same flag and
 * return-code values, same eight entry points, so a program compiled against
 * the reference header links and runs against this library unchanged.  The
 * coding itself runs on the GPU (see aec_b200.h); next_in/next_out are host
 * pointers exactly as in the reference.
 *
 * The offset-index calls at the end are an extension modelled on later
 * upstream libaec releases; reference 0.3.4 has no counterpart.
 */
#ifndef LIBAEC_H
#define LIBAEC_H 1

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

struct internal_state;

/* zlib-style stream descriptor (reference: libaec.h:67-97). The caller owns
 * the struct and both buffers; the library owns `state` between init and end. */
struct aec_stream {
    const unsigned char *next_in;   /* next input byte */
    size_t avail_in;                /* bytes available at next_in */
    size_t total_in;                /* input bytes consumed so far */
    unsigned char *next_out;        /* where the next output byte goes */
    size_t avail_out;               /* free bytes at next_out */
    size_t total_out;               /* output bytes produced so far */
    unsigned int bits_per_sample;   /* sample resolution, 1..32 */
    unsigned int block_size;        /* samples per block: 8, 16, 32 or 64 */
    unsigned int rsi;               /* blocks per reference sample interval, <= 4096 */
    unsigned int flags;             /* AEC_DATA_* | AEC_RESTRICTED | ... */
    struct internal_state *state;
};

/* sample description flags (reference: libaec.h:102-124) */
#define AEC_DATA_SIGNED     1   /* samples are two's complement */
#define AEC_DATA_3BYTE      2   /* 17..24 bit samples occupy 3 bytes instead of 4 */
#define AEC_DATA_MSB        4   /* most significant byte first (default: LSB first) */
#define AEC_DATA_PREPROCESS 8   /* unit-delay predictor + mapper */
#define AEC_RESTRICTED      16  /* restricted option set, bits_per_sample <= 4 */
#define AEC_PAD_RSI         32  /* RSIs start on byte boundaries */
#define AEC_NOT_ENFORCE     64  /* allow any even block size */

/* return codes (reference: libaec.h:129-133) */
#define AEC_OK            0
#define AEC_CONF_ERROR   (-1)
#define AEC_STREAM_ERROR (-2)
#define AEC_DATA_ERROR   (-3)
#define AEC_MEM_ERROR    (-4)

/* flush modes (reference: libaec.h:139-149) */
#define AEC_NO_FLUSH 0   /* more input will follow */
#define AEC_FLUSH    1   /* finish the stream; call until avail_out stays > 0 */

int aec_encode_init(struct aec_stream *strm);
int aec_encode(struct aec_stream *strm, int flush);
int aec_encode_end(struct aec_stream *strm);

int aec_decode_init(struct aec_stream *strm);
int aec_decode(struct aec_stream *strm, int flush);
int aec_decode_end(struct aec_stream *strm);

/* init + code(AEC_FLUSH) + end on one buffer */
int aec_buffer_encode(struct aec_stream *strm);
int aec_buffer_decode(struct aec_stream *strm);

/* ---- extension: RSI offset index (bit offsets of every RSI start) ---- */

/* Record offsets while encoding (call after aec_encode_init). */
int aec_encode_enable_offsets(struct aec_stream *strm);
int aec_encode_count_offsets(struct aec_stream *strm, size_t *count);
int aec_encode_get_offsets(struct aec_stream *strm, size_t *offsets, size_t offsets_count);
/* Hand a previously recorded index to the decoder (call after aec_decode_init
 * and before the first aec_decode): every RSI is then decoded in parallel
 * instead of discovering the boundaries sequentially. */
int aec_decode_set_offsets(struct aec_stream *strm, const size_t *offsets, size_t offsets_count);
/* Offsets a decoder found on its own while decoding a stream without an index:
 * enable after aec_decode_init, read after the stream has been decoded and
 * before aec_decode_end. */
int aec_decode_enable_offsets(struct aec_stream *strm);
int aec_decode_count_offsets(struct aec_stream *strm, size_t *count);
int aec_decode_get_offsets(struct aec_stream *strm, size_t *offsets, size_t offsets_count);
/* Random access: decode `size` bytes of samples starting at byte `pos` of the
 * uncompressed data.  strm->next_in / avail_in describe the WHOLE compressed
 * stream, next_out / avail_out the destination (avail_out >= size); pos and
 * size are multiples of the sample size.  Only the RSIs that hold the range are
 * read and decoded.  (Same signature as aec_decode_range of later libaec
 * releases; not in the 0.3.4 reference) */
int aec_decode_range(struct aec_stream *strm, const size_t *rsi_offsets, size_t rsi_offsets_count,
                     size_t pos, size_t size);

#ifdef __cplusplus
}
#endif
#endif /* LIBAEC_H */
