"""Generates tests/golden/dq_refcpu.npz by EXECUTING THE REFERENCE (run in the build container,
where /root/reference exists):

  * `ref_cpu`   = output of the reference's own CPU simulator, unmodified
                  (cpu::DequantizeAndUnpackWeight3567_v2<ushort,B>, /root/reference/src/cpp_simulate.cc:568-691,
                  compiled by oracle/Makefile into oracle/_ref/libxbit_refcpu.so) on the packed inputs stored
                  beside it;
  * `ref_ints`  = the same function with scales == 1.0 and qzeros == 0, i.e. the reference's unpacked
                  integers (0..255 are exact in fp16);
  * `ref_zints` = the same function with qweight == 0 and scales == -1.0, i.e. the reference's unpacked
                  zero points.

The reference ships no golden vectors of its own; these are outputs of its code.
A second file, dq_refgpu.npz, holds outputs of the reference's GPU kernels and is produced on a
B200 by tests/golden/make_golden_gpu.py.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from xbitops_b200 import synth  # noqa: E402

CASES = [(bits, g, K, N) for bits in range(2, 9) for (g, K, N) in ((32, 96, 16), (128, 256, 24), (64, 200, 16))]


def main():
    rc = O.RefCpu()
    out = {}
    for (bits, g, K, N) in CASES:
        tag = f"b{bits}_g{g}_K{K}_N{N}"
        mode = "bits" if g == 64 else "gptq"
        qw, s, qz, _ = synth.make_inputs(K, N, bits, g, seed=1000 + bits, scale_mode=mode)
        out[tag + "_qweight"] = qw
        out[tag + "_scales"] = s.view(np.uint16)
        out[tag + "_qzeros"] = qz
        out[tag + "_ref_cpu"] = rc.dequant(qw, s, qz, g, bits, K).view(np.uint16)
        out[tag + "_ref_ints"] = rc.dequant(qw, np.ones_like(s), np.zeros_like(qz), g, bits, K).astype(np.uint8)
        zi = rc.dequant(np.zeros_like(qw), -np.ones_like(s), qz, g, bits, K).astype(np.float32)
        out[tag + "_ref_zints"] = zi[::g][: (K + g - 1) // g].astype(np.uint8)   # one row per group
    path = os.path.join(ROOT, "tests", "golden", "dq_refcpu.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(CASES), "cases")


if __name__ == "__main__":
    main()
