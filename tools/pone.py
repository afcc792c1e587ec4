"""Developer helper for ncu: a few calls of one family on one shape (rotating weight sets).
    python tools/pone.py K N FAMILY [M] [CALLS]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from xbitops_b200 import capi  # noqa: E402
import sweep  # noqa: E402

lib = capi.load()
fam = int(sys.argv[3])
M = int(sys.argv[4]) if len(sys.argv) > 4 else 1
calls = int(sys.argv[5]) if len(sys.argv) > 5 else 8
# PONE_SHAPES="K,N;K,N;...": several shapes in one process (one ncu invocation), `calls` launches each
shapes = [tuple(int(x) for x in s.split(",")) for s in os.environ["PONE_SHAPES"].split(";")] if os.environ.get("PONE_SHAPES") \
    else [(int(sys.argv[1]), int(sys.argv[2]))]
for K, N in shapes:
    R, qw, sc, qz, a, out, nbytes = sweep.make(K, N, R=8, M=M)
    for i in range(calls):
        rc = lib.xbit_gemv_f16_ex(a.data_ptr(), qw[i % R].data_ptr(), sc[i % R].data_ptr(), qz[i % R].data_ptr(), out[i % R].data_ptr(),
                                  M, K, N, 4, 128, 0, N, sweep.WS.data_ptr(), sweep.WS.numel(), fam | capi.GEMV_FLAG_STATIC_WEIGHTS,
                                  torch.cuda.current_stream().cuda_stream)
        assert rc == 0, capi.last_error()
    torch.cuda.synchronize()
    del qw, sc, qz, out
print("ok")
