"""Regenerate the committed golden vectors from the UNMODIFIED reference.

Run in a container that has /root/reference (after ``make -C oracle``):

    python tests/golden/make_golden.py

Writes tests/golden/ref_vectors.npz: for each seeded case of tests/cases.py the
parameters, the raw input, the reference's compressed stream (default build
and, for AEC_PAD_RSI cases, the -DENABLE_RSI_PADDING build) and the reference's
decoded output for three output sizes.  tests/golden/typical.rz is a verbatim
copy of the reference's own golden file /root/reference/data/typical.rz
(sha256 16a7f994...3f6068; decodes with -n16 -j64 -r256 -m to sha256
e6e1bf68...6df896).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from cases import random_case  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

N_CASES = 240


def main():
    assert po.ref_available(), "build the reference first: make -C oracle"
    rec = {}
    meta = []
    for seed in range(N_CASES):
        pad = seed % 3 == 2
        p, raw = random_case(10_000 + seed, allow_pad=pad, max_samples=1500)
        pad_build = pad and bool(p.flags & po.AEC_PAD_RSI)
        enc = po.ref_encode(p, raw, pad_rsi_build=pad_build)
        B = p.bytes_per_sample
        ns = len(raw) // B
        sizes = [ns * B, (ns // 2) * B, ns * B + 64 * B]
        outs = [po.ref_decode(p, enc["out"], s) for s in sizes]
        meta.append([p.bits_per_sample, p.block_size, p.rsi, p.flags, int(pad_build),
                     enc["status"], enc["total_in"]] + sizes + [o["status"] for o in outs])
        rec[f"raw{seed}"] = np.asarray(raw, dtype=np.uint8)
        rec[f"enc{seed}"] = enc["out"]
        for j, o in enumerate(outs):
            rec[f"dec{seed}_{j}"] = o["out"]
    rec["meta"] = np.array(meta, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **rec)
    print("wrote", N_CASES, "cases")


if __name__ == "__main__":
    main()
