"""Developer timing of the dequant kernel (4096 x 11008, rotating buffers, CUDA graph): us and fraction of the copy peak.
    python tools/pdq.py [BITS GROUPSIZE]...     env XBIT_DQ_SMEM_KB = occupancy cap under test"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from xbitops_b200 import capi, synth  # noqa: E402
import sweep  # noqa: E402

lib = capi.load()
K, N, R = 4096, 11008, 3
dev = torch.device("cuda:0")
argv = sys.argv[1:]
cases = [(int(argv[i]), int(argv[i + 1])) for i in range(0, len(argv) - 1, 2)] or [(4, 128), (3, 128), (8, 32), (2, 128)]
out = torch.empty((R, K, N), device=dev, dtype=torch.float16)
row = f"cap={os.environ.get('XBIT_DQ_SMEM_KB', '0'):>3s} KB:"
for (b, g) in cases:
    qw = torch.randint(-2**31, 2**31 - 1, (R, (K * b + 31) // 32, N), dtype=torch.int32, device=dev)
    qz = torch.randint(-2**31, 2**31 - 1, (R, K // g, (N * b + 31) // 32), dtype=torch.int32, device=dev)
    sc = (torch.rand((R, K // g, N), device=dev) * 0.018 + 0.002).to(torch.float16)

    def fn(i):
        j = i % R
        rc = lib.xbit_dequant_f16(qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(), out[j].data_ptr(), K, N, b, g, 1, torch.cuda.current_stream().cuda_stream)
        assert rc == 0, capi.last_error()
    us = sweep.time_graph(fn, 2 * R)
    row += f"  b{b} g{g} {us:6.2f} us {synth.dq_bytes(K, N, b, g) / us / 1e3 / sweep.PEAK * 100:3.0f}%"
print(row, flush=True)
