"""Developer check of the tcgen05 GEMV family against the fp64 truth (small -> large), then timing."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xbitops_b200 as X  # noqa: E402
from xbitops_b200 import capi, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

dev = torch.device("cuda:0")
co = O.COracle()
FAM = capi.GEMV_TCGEN05
stage = sys.argv[1] if len(sys.argv) > 1 else "all"


def case(M, K, N, bias, seed=0):
    qw, s, qz, a = synth.make_inputs(K, N, 4, 128, M=M, seed=seed)
    w = co.dequant(qw, s, qz, 128, 4, K, bias)
    y64 = a.astype(np.float64) @ w.astype(np.float64)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
    got = X.gemv(d(a.view(np.int16)).view(torch.float16), d(qw), d(s.view(np.int16)).view(torch.float16), d(qz), 128, 4, K, bias, family=FAM)
    torch.cuda.synchronize()
    e = float(np.abs(got.cpu().numpy().astype(np.float64) - y64).max() / np.abs(y64).max())
    print(f"  M={M} K={K} N={N} bias={bias}: err {e:.3e}", flush=True)
    return e


if stage in ("small", "all"):
    errs = [case(1, 256, 128, 0), case(1, 512, 128, 1), case(1, 1024, 256, 0), case(2, 512, 128, 0), case(3, 1024, 384, 1), case(8, 2048, 512, 0)]
    print("small worst", max(errs))
if stage in ("big", "all"):
    errs = [case(1, 4096, 4096, 1), case(1, 11008, 4096, 0), case(1, 4096, 11008, 0), case(5, 8192, 8192, 1), case(1, 8192, 28672, 0), case(1, 4096, 160, 0)]
    print("big worst", max(errs))
if stage in ("time", "all"):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import sweep
    lib = capi.load()
    for (K, N) in ((4096, 4096), (4096, 11008), (11008, 4096), (8192, 8192), (8192, 28672), (28672, 8192)):
        R, qw, sc, qz, a, out, nbytes = sweep.make(K, N)
        row = f"== {K}x{N}: roofline {nbytes/6549.8/1e3:.2f} us |"
        for fam, name in ((capi.GEMV_MMA, "mma"), (capi.GEMV_TCGEN05, "tc5")):
            for ring in ((4,) if fam == capi.GEMV_MMA else (3, 4, 5, 6)):
                capi.set_option("XBIT_GEMV_RING", int(str(ring)))

                def fn(i):
                    j = i % R
                    rc = lib.xbit_gemv_f16_ex(a.data_ptr(), qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(), out[j].data_ptr(),
                                              1, K, N, 4, 128, 0, N, None, 0, fam | capi.GEMV_FLAG_STATIC_WEIGHTS,
                                              torch.cuda.current_stream().cuda_stream)
                    assert rc == 0, capi.last_error()
                us = sweep.time_graph(fn, R)
                row += f"  {name} r{ring} {us:6.2f}us {nbytes/us/1e3/6549.8*100:3.0f}%"
        print(row, flush=True)
        del qw, sc, qz, out
