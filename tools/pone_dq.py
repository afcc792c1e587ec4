"""Developer helper for ncu: a few dequant calls of one (bits, groupsize) on 4096 x 11008 (rotating buffers).
    python tools/pone_dq.py BITS GROUPSIZE [CALLS]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xbitops_b200 import capi  # noqa: E402

lib = capi.load()
b, g = int(sys.argv[1]), int(sys.argv[2])
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 6
K, N, R = 4096, 11008, 3
dev = torch.device("cuda:0")
qw = torch.randint(-2**31, 2**31 - 1, (R, (K * b + 31) // 32, N), dtype=torch.int32, device=dev)
qz = torch.randint(-2**31, 2**31 - 1, (R, K // g, (N * b + 31) // 32), dtype=torch.int32, device=dev)
sc = (torch.rand((R, K // g, N), device=dev) * 0.018 + 0.002).to(torch.float16)
out = torch.empty((R, K, N), device=dev, dtype=torch.float16)
for i in range(calls):
    j = i % R
    rc = lib.xbit_dequant_f16(qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(), out[j].data_ptr(), K, N, b, g, 1, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, capi.last_error()
torch.cuda.synchronize()
print("ok")
