"""libaec_b200 -- B200-native CCSDS 121.0-B-2 (libaec-compatible) coder.

The product is the native library ``libaec_b200/lib/libaec.so.0`` (CUDA kernels
for sm_100a behind the libaec.h C API and the aec_b200.h device C ABI) plus
``libsz.so.2``.  This package is the thin Python host mirror of that interface
(ctypes; torch only for device memory, streams and torch.distributed).
"""
from .api import (AEC_DATA_3BYTE, AEC_DATA_MSB, AEC_DATA_PREPROCESS, AEC_DATA_SIGNED,  # noqa: F401
                  AEC_NOT_ENFORCE, AEC_PAD_RSI, AEC_RESTRICTED, AEC_OK, AEC_CONF_ERROR,
                  AEC_STREAM_ERROR, AEC_DATA_ERROR, AEC_MEM_ERROR, AEC_FLUSH, AEC_NO_FLUSH,
                  Params, AecStream, Carry, DeviceCodec, Encoder, Decoder, buffer_encode, buffer_decode, decode_range, buffer_decode_discover,
                  encode_bound, sz_compress, sz_decompress, sz_compress_batch, sz_decompress_batch, load_library)
