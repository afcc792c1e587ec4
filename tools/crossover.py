"""BASELINE.json configs[4]: A16W4 skinny GEMM, M = 1..16 on 8192x8192 -- kernel family crossover.
us/call (CUDA graph, rotating weights > L2) and fraction of the measured HBM peak per family."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from xbitops_b200 import capi, synth  # noqa: E402
import sweep  # noqa: E402

lib = capi.load()
PEAK = 6549.8
K = N = 8192
R, qw, sc, qz, a, out, nbytes1 = sweep.make(K, N)
print(f"== skinny GEMM {K}x{N}, bits 4, g128 (roofline {nbytes1/PEAK/1e3:.2f} us at M=1)")
for M in (1, 2, 3, 4, 6, 8, 12, 16):
    nb = synth.gemv_bytes(K, N, 4, 128, M)
    row = f"   M={M:2d}:"
    for fam, name in ((capi.GEMV_SIMT, "simt"), (capi.GEMV_MMA, "mma.sync"), (capi.GEMV_PERSIST, "persist")):
        if (fam == capi.GEMV_SIMT and M > 2) or (fam == capi.GEMV_PERSIST and M > 8):
            continue

        def fn(i):
            j = i % R
            rc = lib.xbit_gemv_f16_ex(a.data_ptr(), qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(), out[j].data_ptr(),
                                      M, K, N, 4, 128, 0, N, None, 0, fam | capi.GEMV_FLAG_STATIC_WEIGHTS,
                                      torch.cuda.current_stream().cuda_stream)
            assert rc == 0, capi.last_error()
        us = sweep.time_graph(fn, R)
        row += f"  {name} {us:6.2f}us {nb/us/1e3/PEAK*100:3.0f}%"
    print(row, flush=True)
