"""Developer timing of xbit_gemv_f16_multi: P matrices sharing one activation vector in one launch against P separate
calls (rotating weight sets > L2, one CUDA graph, the bench protocol).
    python tools/pmulti.py [K N P]...          default: QKV of Llama-2-7B (4096 4096 3) and gate + up (4096 11008 2)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from xbitops_b200 import capi, synth  # noqa: E402
import sweep  # noqa: E402

lib = capi.load()
PEAK, WS = sweep.PEAK, sweep.WS


def run(K, N, P):
    R, qw, sc, qz, a, out, _ = sweep.make(K, N, M=1)
    R -= R % P
    nbytes = synth.gemv_bytes(K, N, 4, 128, 1)
    st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    probs = []
    for j in range(0, R, P):
        arr = (capi.GemvProblem * P)()
        for i in range(P):
            arr[i] = capi.GemvProblem(qw[j + i].data_ptr(), sc[j + i].data_ptr(), qz[j + i].data_ptr(), out[j + i].data_ptr(), N, N)
        probs.append(arr)

    def fused(i):
        rc = lib.xbit_gemv_f16_multi(a.data_ptr(), ctypes.cast(probs[i % len(probs)], ctypes.c_void_p), P, 1, K, 4, 128, 0, WS.data_ptr(), WS.numel(),
                                     capi.GEMV_AUTO | capi.GEMV_FLAG_STATIC_WEIGHTS, st())
        assert rc == 0, capi.last_error()

    def separate(i):
        j = i % R
        rc = lib.xbit_gemv_f16_ex(a.data_ptr(), qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(), out[j].data_ptr(), 1, K, N, 4, 128, 0, N,
                                  WS.data_ptr(), WS.numel(), capi.GEMV_AUTO | capi.GEMV_FLAG_STATIC_WEIGHTS, st())
        assert rc == 0, capi.last_error()

    us_f = sweep.time_graph(fused, len(probs))
    us_s = sweep.time_graph(separate, R)
    print(f"== {P} x {K}x{N}: fused {us_f:6.2f} us per launch = {us_f / P:5.2f} us per matrix ({nbytes * P / us_f / 1e3 / PEAK * 100:3.0f}% of peak); "
          f"separate {us_s:5.2f} us per call ({nbytes / us_s / 1e3 / PEAK * 100:3.0f}%)", flush=True)


def main():
    argv = sys.argv[1:]
    cases = [(int(argv[i]), int(argv[i + 1]), int(argv[i + 2])) for i in range(0, len(argv) - 2, 3)] or [(4096, 4096, 3), (4096, 11008, 2), (4096, 4096, 4), (8192, 8192, 3)]
    for (K, N, P) in cases:
        run(K, N, P)


if __name__ == "__main__":
    main()
