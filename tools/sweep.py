"""Developer sweep: GEMV us/call for the planner knob XBIT_GEMV_SPLITS x family, rotating weights > L2,
CUDA-graph timed.  Not the bench; used to pick the planner heuristics.
    python tools/sweep.py [K N]...   ;  --one K N FAMILY [M]: just launch a few calls (for ncu)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xbitops_b200 import capi, synth  # noqa: E402

dev = torch.device("cuda:0")
lib = capi.load()
WS = torch.zeros(max(256, lib.xbit_gemv_workspace_bytes(16, 0, 0, 4, 128)), dtype=torch.uint8, device=dev)
PEAK = 6549.8


def make(K, N, R=None, M=16):
    g, bits = 128, 4
    nbytes = synth.gemv_bytes(K, N, bits, g)
    R = R or max(2, (1 << 30) // nbytes + 1)
    G = K // g
    qw = torch.randint(-2**31, 2**31 - 1, (R, K // 8, N), dtype=torch.int32, device=dev)
    sc = (torch.rand((R, G, N), device=dev) * 0.018 + 0.002).to(torch.float16)
    qz = torch.randint(-2**31, 2**31 - 1, (R, G, N // 8), dtype=torch.int32, device=dev)
    a = torch.randn((M, K), device=dev, dtype=torch.float16)
    out = torch.empty((R, M, N), device=dev, dtype=torch.float16)
    return R, qw, sc, qz, a, out, nbytes


def time_graph(fn, calls, reps=15, warm=3):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(0)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(calls):
            fn(i)
    for _ in range(warm):
        g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / calls)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        K, N, fam = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
        M = int(sys.argv[5]) if len(sys.argv) > 5 else 1
        R, qw, sc, qz, a, out, nbytes = make(K, N, R=8)
        for i in range(8):
            rc = lib.xbit_gemv_f16_ex(a.data_ptr(), qw[i % R].data_ptr(), sc[i % R].data_ptr(), qz[i % R].data_ptr(),
                                      out[i % R].data_ptr(), M, K, N, 4, 128, 0, N, WS.data_ptr(), WS.numel(), fam,
                                      torch.cuda.current_stream().cuda_stream)
            assert rc == 0, capi.last_error()
        torch.cuda.synchronize()
        print("ok")
        return
    shapes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)] or \
        [(4096, 4096), (4096, 11008), (11008, 4096), (8192, 8192), (8192, 28672), (28672, 8192)]
    for (K, N) in shapes:
        R, qw, sc, qz, a, out, nbytes = make(K, N)
        print(f"== {K}x{N} {nbytes/1e6:.1f} MB R={R} roofline {nbytes/PEAK/1e3:.2f} us")
        capi.set_option("XBIT_GEMV_STREAMK", int("0"))
        fams = ((capi.GEMV_SIMT, "simt      ", 0), (capi.GEMV_MMA, "mma       ", 0))
        if os.environ.get("SWEEP_MMA_ONLY"):
            fams = fams[1:2]
        ring_env = os.environ.get("SWEEP_RING", "0")
        for fam, name, hyb in fams:
            capi.set_option("XBIT_GEMV_RING", int(ring_env))
            for wc in (0, 2, 4, 8):
                row = f"   {name} wc={wc if wc else 'A'}:"
                for splits in ((0,) if wc == 0 else (1, 2, 3, 4, 5, 6, 7, 8)):
                    capi.set_option("XBIT_GEMV_SPLITS", int(str(splits)))
                    capi.set_option("XBIT_GEMV_WC", int(str(wc)))
                    flags = capi.GEMV_FLAG_STATIC_WEIGHTS

                    def fn(i):
                        j = i % R
                        rc = lib.xbit_gemv_f16_ex(a.data_ptr(), qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(),
                                                  out[j].data_ptr(), 1, K, N, 4, 128, 0, N, None, 0,
                                                  fam | flags, torch.cuda.current_stream().cuda_stream)
                        assert rc == 0, capi.last_error()
                    try:
                        us = time_graph(fn, R)
                        row += f"  s{splits if splits else 'A'} {us:6.2f}us {nbytes/us/1e3/PEAK*100:3.0f}%"
                    except AssertionError:
                        row += f"  s{splits} n/a"
                print(row, flush=True)
        capi.set_option("XBIT_GEMV_WC", int("0"))
        capi.set_option("XBIT_GEMV_SPLITS", int("0"))
        capi.set_option("XBIT_GEMV_RING", int("0"))
        capi.set_option("XBIT_GEMV_WC", int("0"))
        capi.set_option("XBIT_GEMV_SPLITS", int("0"))
        del qw, sc, qz, out


if __name__ == "__main__":
    main()
