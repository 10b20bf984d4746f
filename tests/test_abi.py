"""CPU checks of the drop-in boundary: the native library loads and exports
every symbol include/*.h declares, the struct layout matches the reference
ABI, parameter validation answers like the reference, and -- with no GPU --
the product path fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest

import libaec_b200 as L
from libaec_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"^\s*#.*$", "", txt, flags=re.M)          # macros are not symbols
    return set(re.findall(r"\b(aec_[a-z_]+|aecb200_[a-z_]+|SZ_[A-Za-z_]+)\s*\(", txt))


def test_library_exports_every_declared_symbol():
    lib = L.load_library()
    for name in sorted(_declared("libaec.h") | _declared("aec_b200.h")):
        assert hasattr(lib, name), name
    sz = api.load_sz_library()
    for name in sorted(_declared("szlib.h")):
        assert hasattr(sz, name), name
    assert set(api.LIBAEC_SYMBOLS) <= _declared("libaec.h")
    assert set(api.DEVICE_SYMBOLS) <= _declared("aec_b200.h")


def test_struct_layout_matches_reference_abi():
    # 72-byte LP64 struct (reference libaec.h:67-97)
    assert C.sizeof(L.AecStream) == 72
    assert L.AecStream.next_out.offset == 24 and L.AecStream.bits_per_sample.offset == 48
    assert L.AecStream.state.offset == 64


def test_parameter_validation_without_device():
    lib = L.load_library()
    bad = [(0, 16, 128, 0), (33, 16, 128, 0), (8, 12, 128, 0), (8, 13, 128, L.AEC_NOT_ENFORCE),
           (8, 16, 4097, 0), (5, 16, 16, L.AEC_RESTRICTED)]
    for n, J, rsi, flags in bad:
        s = L.AecStream()
        s.bits_per_sample, s.block_size, s.rsi, s.flags = n, J, rsi, flags
        assert lib.aec_encode_init(C.byref(s)) == L.AEC_CONF_ERROR, (n, J, rsi, flags)
    for n in (0, 33):
        s = L.AecStream()
        s.bits_per_sample, s.block_size, s.rsi, s.flags = n, 16, 128, 0
        assert lib.aec_decode_init(C.byref(s)) == L.AEC_CONF_ERROR


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path cannot be observed")
    with pytest.raises(RuntimeError):
        L.DeviceCodec()
    res = L.buffer_encode(L.Params(8, 16, 16, 0), bytes(range(64)))
    assert res["status"] != L.AEC_OK and res["out"].size == 0


def test_product_never_touches_the_oracle():
    """Nothing under libaec_b200/ may import or load oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "libaec_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "aec_oracle" not in txt, f
