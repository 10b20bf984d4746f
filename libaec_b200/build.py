"""Build the native libraries in-tree (they travel to the GPU box with the snapshot).

    python -m libaec_b200.build            # build if sources are newer
    python -m libaec_b200.build --force

Outputs (git-ignored, *.so):
    libaec_b200/lib/libaec.so.0  (+ libaec.so)   CUDA kernels + runtime + libaec.h API
    libaec_b200/lib/libsz.so.2   (+ libsz.so)    szlib.h shim, links libaec
    tests/_build/libaec_cpumodel.so              CPU harness around the host+device block code (tests only)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib")
OBJ = os.path.join(PKG, "lib", "obj")
TESTBUILD = os.path.join(ROOT, "tests", "_build")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVFLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--use_fast_math"]

CU_SOURCES = ["aec_encode.cu", "aec_decode.cu", "aec_skim.cu", "aec_sz.cu", "aec_runtime.cu"]
HEADERS = ["aec_core.cuh", "aec_decode_core.cuh", "aec_skim_core.cuh", "aec_device.h",
           os.path.join(ROOT, "include", "aec_b200.h"), os.path.join(ROOT, "include", "libaec.h"),
           os.path.join(ROOT, "include", "szlib.h")]


def _newer(srcs, target) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build step failed: " + cmd[0])
    return r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> dict:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(TESTBUILD, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    libaec = os.path.join(LIB, "libaec.so.0")
    libsz = os.path.join(LIB, "libsz.so.2")
    model = os.path.join(TESTBUILD, "libaec_cpumodel.so")

    jobs = []
    objs = []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer([s] + hdrs, o):
            jobs.append([NVCC] + ARCH + NVFLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
    for cname in ("libaec_api", "sz_batch"):
        c_o = os.path.join(OBJ, cname + ".o")
        c_c = os.path.join(CSRC, cname + ".c")
        objs.append(c_o)
        if force or _newer([c_c] + hdrs, c_o):
            jobs.append(["gcc", "-O2", "-fPIC", "-std=c99", "-Wall", "-D_POSIX_C_SOURCE=200809L", "-c", c_c, "-o", c_o])
    with ThreadPoolExecutor(max_workers=4) as ex:
        outs = list(ex.map(_run, jobs))
    if verbose:
        for o in outs:
            sys.stderr.write(o)
    if force or jobs or not os.path.exists(libaec):
        _run([NVCC] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-Xlinker", "-soname,libaec.so.0",
                              "-o", libaec] + objs + ["-lpthread"])
        link = os.path.join(LIB, "libaec.so")
        if os.path.lexists(link):
            os.remove(link)
        os.symlink("libaec.so.0", link)
    sz_c = os.path.join(CSRC, "sz_api.c")
    if force or _newer([sz_c, libaec] + hdrs, libsz):
        _run(["gcc", "-O2", "-fPIC", "-std=c99", "-Wall", "-shared", "-Wl,-soname,libsz.so.2",
              "-Wl,-rpath,$ORIGIN", "-o", libsz, sz_c, "-L" + LIB, "-laec"])
        link = os.path.join(LIB, "libsz.so")
        if os.path.lexists(link):
            os.remove(link)
        os.symlink("libsz.so.2", link)
    cm = os.path.join(CSRC, "cpu_model.cpp")
    if force or _newer([cm] + hdrs, model):
        _run(["g++", "-O2", "-fPIC", "-shared", "-std=c++14", "-Wall", "-o", model, cm])
    return {"libaec": libaec, "libsz": libsz, "cpumodel": model}


if __name__ == "__main__":
    out = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    for k, v in out.items():
        print(k, v)
