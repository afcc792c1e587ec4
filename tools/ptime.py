"""Developer timing: the persistent W4 schedule (XBIT_GEMV_PERSIST) against the cluster split-K kernel (XBIT_GEMV_MMA),
us/call with rotating weights > L2 in one CUDA graph (the bench protocol), for the knobs of plan_w4p.
    python tools/ptime.py [--m M] [K N]...        env: PTIME_VARIANTS="fine ring grid [warps];..." e.g. "1 0 0;0 0 0;1 4 0 16"
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from xbitops_b200 import capi, synth  # noqa: E402
import sweep  # noqa: E402

lib = capi.load()
PEAK = sweep.PEAK
WS = sweep.WS


def main():
    argv = sys.argv[1:]
    M = 1
    if argv and argv[0] == "--m":
        M = int(argv[1])
        argv = argv[2:]
    shapes = [(int(argv[i]), int(argv[i + 1])) for i in range(0, len(argv) - 1, 2)] or \
        [(4096, 4096), (4096, 11008), (11008, 4096), (8192, 8192), (8192, 28672), (28672, 8192)]
    variants = [tuple(int(x) for x in v.split()) for v in os.environ.get("PTIME_VARIANTS", "1 0 0;0 0 0").split(";")]
    for (K, N) in shapes:
        R, qw, sc, qz, a, out, _ = sweep.make(K, N, M=max(M, 1))
        nbytes = synth.gemv_bytes(K, N, 4, 128, M)
        print(f"== {K}x{N} M={M} {nbytes/1e6:.1f} MB R={R} roofline {nbytes/PEAK/1e3:.2f} us")

        def run(fam, ws=True):
            def fn(i):
                j = i % R
                rc = lib.xbit_gemv_f16_ex(a.data_ptr(), qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(), out[j].data_ptr(),
                                          M, K, N, 4, 128, 0, N, WS.data_ptr() if ws else None, WS.numel() if ws else 0,
                                          fam | capi.GEMV_FLAG_STATIC_WEIGHTS, torch.cuda.current_stream().cuda_stream)
                assert rc == 0, capi.last_error()
            return sweep.time_graph(fn, R)

        capi.set_option("XBIT_GEMV_STREAMK", int("0"))
        us = run(capi.GEMV_MMA, ws=False)
        print(f"   cluster split-K (round 1)      {us:6.2f} us  {nbytes/us/1e3/PEAK*100:3.0f}%", flush=True)
        for var in variants:
            fine, ring, grid = var[:3]
            warps = var[3] if len(var) > 3 else 0
            capi.set_option("XBIT_W4P_FINE", int(str(fine)))
            capi.set_option("XBIT_W4P_RING", int(str(ring)))
            capi.set_option("XBIT_W4P_GRID", int(str(grid)))
            capi.set_option("XBIT_W4P_WARPS", int(str(warps)))
            try:
                us = run(capi.GEMV_PERSIST)
                print(f"   persist fine={fine} ring={ring or 'A'} grid={grid or 'A'} warps={warps or 'A'}   {us:6.2f} us  {nbytes/us/1e3/PEAK*100:3.0f}%", flush=True)
            except AssertionError as ex:
                print(f"   persist fine={fine} ring={ring} grid={grid} warps={warps}: {ex}")
        for k in ("XBIT_W4P_FINE", "XBIT_W4P_RING", "XBIT_W4P_GRID", "XBIT_W4P_WARPS"):
            capi.set_option(k)
        capi.set_option("XBIT_GEMV_STREAMK")
        us = run(capi.GEMV_AUTO)
        print(f"   AUTO (what the op runs)        {us:6.2f} us  {nbytes/us/1e3/PEAK*100:3.0f}%   family {lib.xbit_gemv_pick_family(M, K, N, 4, 128)}", flush=True)
        # correctness spot check against a @ dequant
        import xbitops_b200 as X
        w = X.dequant(qw[0], sc[0], qz[0], 128, 4, K, 0)
        truth = a[:M].double() @ w.double()
        y = X.gemv(a[:M], qw[0], sc[0], qz[0], 128, 4, K, 0, family=capi.GEMV_PERSIST)
        print(f"   persist normalised err {float((y.double() - truth).abs().max() / truth.abs().max()):.3e}")
        del qw, sc, qz, out


if __name__ == "__main__":
    main()
