"""Developer helper for ncu: a few A16W8 / A16W2 calls (AUTO) on one shape.    python tools/pone8.py K N BITS [CALLS]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from xbitops_b200 import capi  # noqa: E402
import sweep  # noqa: E402

lib = capi.load()
dev = torch.device("cuda:0")
K, N, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
calls = int(sys.argv[4]) if len(sys.argv) > 4 else 6
R = 8
qw = torch.randint(-2**31, 2**31 - 1, (R, K * B // 32, N), dtype=torch.int32, device=dev)
sc = (torch.rand((R, K // 128, N), device=dev) * 0.018 + 0.002).to(torch.float16)
qz = torch.randint(-2**31, 2**31 - 1, (R, K // 128, N * B // 32), dtype=torch.int32, device=dev)
a = torch.randn((1, K), device=dev, dtype=torch.float16)
out = torch.empty((R, 1, N), device=dev, dtype=torch.float16)
for i in range(calls):
    j = i % R
    rc = lib.xbit_gemv_f16_ex(a.data_ptr(), qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(), out[j].data_ptr(), 1, K, N, B, 128, 0, N,
                              sweep.WS.data_ptr(), sweep.WS.numel(), capi.GEMV_AUTO | capi.GEMV_FLAG_STATIC_WEIGHTS, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, capi.last_error()
torch.cuda.synchronize()
print("ok")
