"""Developer probe: ring depth and feed ceiling (consumers skipping the math) for the W4 cluster kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xbitops_b200 import capi  # noqa: E402
import sweep  # noqa: E402  (same directory)

lib = capi.load()
PEAK = 6549.8
for (K, N) in ((4096, 11008), (8192, 8192), (8192, 28672), (28672, 8192)):
    R, qw, sc, qz, a, out, nbytes = sweep.make(K, N)
    print(f"== {K}x{N} {nbytes/1e6:.1f} MB roofline {nbytes/PEAK/1e3:.2f} us")
    for skip in (0, 1):
        row = f"   skip_math={skip}:"
        for ring in (2, 3, 4, 5, 6, 8):
            capi.set_option("XBIT_GEMV_RING", int(str(ring)))
            capi.set_option("XBIT_GEMV_DEBUG_SKIP", int(str(skip)))

            def fn(i):
                j = i % R
                rc = lib.xbit_gemv_f16_ex(a.data_ptr(), qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(), out[j].data_ptr(),
                                          1, K, N, 4, 128, 0, N, None, 0, capi.GEMV_MMA | capi.GEMV_FLAG_STATIC_WEIGHTS,
                                          torch.cuda.current_stream().cuda_stream)
                assert rc == 0, capi.last_error()
            us = sweep.time_graph(fn, R)
            row += f"  r{ring} {us:6.2f}us {nbytes/us/1e3/PEAK*100:3.0f}%"
        print(row, flush=True)
    del qw, sc, qz, out
