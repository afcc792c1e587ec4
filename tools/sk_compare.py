"""Developer sweep: cluster split-K schedule vs the persistent stream-K schedule.
    python tools/sk_compare.py [K N]..."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xbitops_b200 import capi  # noqa: E402
from sweep import make, time_graph, PEAK, WS  # noqa: E402

lib = capi.load()


def main():
    shapes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)] or \
        [(4096, 4096), (4096, 11008), (11008, 4096), (8192, 8192), (8192, 28672), (28672, 8192)]
    for (K, N) in shapes:
        R, qw, sc, qz, a, out, nbytes = make(K, N)
        print(f"== {K}x{N} {nbytes/1e6:.1f} MB R={R} roofline {nbytes/PEAK/1e3:.2f} us")
        for fam, name in ((capi.GEMV_MMA, "mma"),):
            row = f"   {name}:"
            for label, sk, ring in (("cluster", 0, 0), ("stream-K", 1, 0), ("stream-K ring 4", 1, 4), ("stream-K ring 6", 1, 6)):
                capi.set_option("XBIT_GEMV_STREAMK", int(str(sk)))
                capi.set_option("XBIT_GEMV_RING", int(str(ring)))

                def fn(i):
                    j = i % R
                    rc = lib.xbit_gemv_f16_ex(a.data_ptr(), qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(), out[j].data_ptr(),
                                              1, K, N, 4, 128, 0, N, WS.data_ptr(), WS.numel(), fam | capi.GEMV_FLAG_STATIC_WEIGHTS,
                                              torch.cuda.current_stream().cuda_stream)
                    assert rc == 0, capi.last_error()
                try:
                    us = time_graph(fn, R)
                    row += f"  {label} {us:6.2f}us {nbytes/us/1e3/PEAK*100:3.0f}%"
                except AssertionError as e:
                    row += f"  {label} n/a"
            print(row, flush=True)
        capi.set_option("XBIT_GEMV_STREAMK", int("0"))
        capi.set_option("XBIT_GEMV_RING", int("0"))
        del qw, sc, qz, out


if __name__ == "__main__":
    main()
