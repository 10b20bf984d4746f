"""T0 acceptance: the reference's OWN test programs and CLI, compiled
unchanged from /root/reference by oracle/Makefile into oracle/_ref/ and linked against the
sonames libaec.so.0 / libsz.so.2, run against OUR libraries through LD_LIBRARY_PATH.

check_code_options is not run: its "small buffers" pass feeds one sample and takes one byte
per call (about 10^8 library calls) and is replaced by the windowed streaming tests in
test_gpu_parity.py; its large-buffer pass is restated there as well."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from libaec_b200 import datagen

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
LIB = os.path.join(ROOT, "libaec_b200", "lib")


def _env():
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = LIB + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    return env


def _need(name):
    path = os.path.join(REF, name)
    if not os.path.exists(path):
        pytest.skip("reference binary %s not built (oracle/Makefile needs /root/reference)" % name)
    return path


def _uses_our_library(binary):
    out = subprocess.run(["ldd", binary], env=_env(), capture_output=True, text=True).stdout
    return "libaec_b200/lib/libaec.so.0" in out


@pytest.mark.parametrize("name", ["check_buffer_sizes", "check_long_fs"])
def test_reference_test_program(name):
    binary = _need(name)
    assert _uses_our_library(binary)
    r = subprocess.run([binary], env=_env(), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "FAIL" not in r.stdout


def test_reference_check_szcomp(tmp_path):
    binary = _need("check_szcomp")
    assert _uses_our_library(binary)
    data = tmp_path / "input.dat"
    datagen.generate("c2", 1 << 18).tofile(str(data))          # any file works
    r = subprocess.run([binary, str(data)], env=_env(), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "differ" not in r.stderr


def test_reference_cli_round_trip(tmp_path):
    """src/aec.c streams the file through aec_encode/aec_decode(AEC_NO_FLUSH) in 10 Mi-sample
    chunks: encode with our library, compare with the oracle's stream, decode back."""
    from oracle import pyoracle as po
    binary = _need("aec")
    assert _uses_our_library(binary)
    raw = datagen.generate("c1", (12 << 20) // 4 + 12345)        # more than one CLI chunk
    src, rz, back = tmp_path / "in.raw", tmp_path / "out.rz", tmp_path / "back.raw"
    raw.tofile(str(src))
    args = ["-n", "32", "-j", "16", "-r", "128", "-s"]            # preprocessing is the CLI default (-N disables)
    r = subprocess.run([binary] + args + [str(src), str(rz)], env=_env(), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    want = po.orc_encode(po.Params(32, 16, 128, po.AEC_DATA_SIGNED | po.AEC_DATA_PREPROCESS), raw)
    got = np.fromfile(str(rz), dtype=np.uint8)
    assert np.array_equal(got, want["out"])
    r = subprocess.run([binary, "-d"] + args + [str(rz), str(back)], env=_env(), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    dec = np.fromfile(str(back), dtype=np.uint8)
    assert hashlib.sha256(dec[: raw.size].tobytes()).hexdigest() == hashlib.sha256(raw.tobytes()).hexdigest()
