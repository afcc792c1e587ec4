"""Generates tests/golden/dq_refgpu.npz ON A B200 by executing the UNMODIFIED reference GPU kernels
(the reference extension built by oracle/build_ref_gpu.sh; DequantizeAndUnpackWeight248 /
DequantizeAndUnpackWeight3567_v2, /root/reference/src/cu/unpack_weight_2_to_7.cu:44-85,219-330, and gemv,
src/cu/gemv_w4a16_pt.cu:35-145).  Inputs are the seeded tensors stored beside the outputs.

    gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/dq_refgpu.npz'   # then copy into tests/golden/

bits == 6: only rows 0..31 are recorded as golden (the reference's GPU b=6 path reads the wrong
word rows beyond the first 32-row block)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from xbitops_b200 import synth  # noqa: E402

CASES = [(bits, g, K, N, bias) for bits in range(2, 9) for (g, K, N) in ((32, 96, 32), (128, 256, 64)) for bias in (0, 1)]
GEMV_CASES = [(1024, 64, 0), (1024, 128, 1), (4096, 64, 0)]      # (K, N, bias): b=4, g=128, M=1


def main():
    ref = O.load_ref_gpu()
    assert ref is not None, "build the reference extension first (oracle/build_ref_gpu.sh)"
    dev = torch.device("cuda:0")
    out = {}
    for (bits, g, K, N, bias) in CASES:
        tag = f"b{bits}_g{g}_K{K}_N{N}_z{bias}"
        qw, s, qz, _ = synth.make_inputs(K, N, bits, g, seed=2000 + bits)
        # one padding group row: the reference reads scale/zero row (k0+32)/g for the last block
        s_pad = np.concatenate([s, np.ones((1, N), np.float16)], 0)
        qz_pad = np.concatenate([qz, np.zeros((1, qz.shape[1]), np.int32)], 0)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        r = ref.dequant(t(qw), t(s_pad.view(np.int16)).view(torch.float16), t(qz_pad), g, bits, K, bias)
        torch.cuda.synchronize()
        r = r.cpu().numpy().view(np.uint16)
        out[tag + "_qweight"], out[tag + "_scales"], out[tag + "_qzeros"] = qw, s.view(np.uint16), qz
        out[tag + "_ref_gpu"] = r[:32] if bits == 6 else r
    for (K, N, bias) in GEMV_CASES:
        tag = f"gemv_K{K}_N{N}_z{bias}"
        qw, s, qz, a = synth.make_inputs(K, N, 4, 128, seed=3000 + N)
        t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
        r = ref.gemv(t(a.view(np.int16)).view(torch.float16), t(qw), t(s.view(np.int16)).view(torch.float16), t(qz),
                     128, 4, K, bias)
        torch.cuda.synchronize()
        out[tag + "_qweight"], out[tag + "_scales"], out[tag + "_qzeros"] = qw, s.view(np.uint16), qz
        out[tag + "_a"] = a.view(np.uint16)
        out[tag + "_ref_gpu"] = r.cpu().numpy().view(np.uint16)
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "dq_refgpu.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
