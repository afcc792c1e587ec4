"""Developer timing: A16W8 (8-bit weights, groupsize 128) GEMV, persistent kernel (AUTO) against the generic kernel,
us/call over rotating weight sets > L2 in one CUDA graph (the bench protocol).    python tools/pw8.py [--m M] [K N]..."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from xbitops_b200 import capi, synth  # noqa: E402
import sweep  # noqa: E402

lib = capi.load()
dev = torch.device("cuda:0")


def main():
    argv = sys.argv[1:]
    M, B = 1, 8
    while argv and argv[0] in ("--m", "--bits"):
        if argv[0] == "--m":
            M = int(argv[1])
        else:
            B = int(argv[1])
        argv = argv[2:]
    shapes = [(int(argv[i]), int(argv[i + 1])) for i in range(0, len(argv) - 1, 2)] or [(4096, 4096), (4096, 11008), (11008, 4096), (8192, 8192)]
    for (K, N) in shapes:
        nbytes = synth.gemv_bytes(K, N, B, 128, M)
        R = max(2, (1 << 30) // nbytes + 1)
        G = K // 128
        qw = torch.randint(-2**31, 2**31 - 1, (R, K * B // 32, N), dtype=torch.int32, device=dev)
        sc = (torch.rand((R, G, N), device=dev) * 0.018 + 0.002).to(torch.float16)
        qz = torch.randint(-2**31, 2**31 - 1, (R, G, N * B // 32), dtype=torch.int32, device=dev)
        a = torch.randn((M, K), device=dev, dtype=torch.float16)
        out = torch.empty((R, M, N), device=dev, dtype=torch.float16)
        print(f"== W{B} {K}x{N} M={M} {nbytes/1e6:.1f} MB R={R} roofline {nbytes/sweep.PEAK/1e3:.2f} us")
        for fam, name in ((capi.GEMV_AUTO, "AUTO (persistent, integer math)"), (capi.GEMV_GENERIC, "generic kernel")):
            def fn(i):
                j = i % R
                rc = lib.xbit_gemv_f16_ex(a.data_ptr(), qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(), out[j].data_ptr(),
                                          M, K, N, B, 128, 0, N, sweep.WS.data_ptr(), sweep.WS.numel(),
                                          fam | capi.GEMV_FLAG_STATIC_WEIGHTS, torch.cuda.current_stream().cuda_stream)
                assert rc == 0, capi.last_error()
            us = sweep.time_graph(fn, R)
            print(f"   {name:34s} {us:7.2f} us  {nbytes/us/1e3/sweep.PEAK*100:3.0f}%", flush=True)
        del qw, sc, qz, out


if __name__ == "__main__":
    main()
