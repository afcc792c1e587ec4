"""Small invocations of every kernel family for compute-sanitizer (tools/sanitize.sh): dequant block / element kernels,
the persistent W4 GEMV (integer and fp16 block math, tile-aligned and block-granular with the cross-CTA fix-up, several
matrices per launch), the cluster split-K kernel (clustered and not), stream-K, the generic kernel, and the
world = 1 forms of the peer / signal / flag-in-data entry points.  Results are checked against a @ dequant."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xbitops_b200 import capi  # noqa: E402
import xbitops_b200 as X  # noqa: E402

dev = torch.device("cuda:0")
lib = capi.load()
gen = torch.Generator(device=dev).manual_seed(7)


def rand(K, N, bits, g):
    qw = torch.randint(-2**31, 2**31 - 1, ((K * bits + 31) // 32, N), dtype=torch.int32, device=dev, generator=gen)
    qz = torch.randint(-2**31, 2**31 - 1, ((K + g - 1) // g, (N * bits + 31) // 32), dtype=torch.int32, device=dev, generator=gen)
    s = (torch.rand(((K + g - 1) // g, N), device=dev, generator=gen) * 0.018 + 0.002).to(torch.float16)
    return qw, s, qz


def check(y, a, w, what):
    truth = a.double() @ w.double()
    err = float((y.double() - truth).abs().max() / truth.abs().max())
    assert err < 1e-2, (what, err)
    print(f"ok {what}: {err:.2e}", flush=True)


# dequant: block kernel (3, 4, 7 bits), element fallback (N % 8 != 0)
for (K, N, b, g) in ((256, 512, 4, 128), (192, 256, 3, 64), (128, 128, 7, 32), (96, 36, 5, 48)):
    qw, s, qz = rand(K, N, b, g)
    w = X.dequant(qw, s, qz, g, b, K, 1)
    assert w.shape == (K, N) and bool(torch.isfinite(w).all())
    print(f"ok dequant b={b} {K}x{N}", flush=True)

cases = [  # (K, N, M, g, family, env)
    (1024, 512, 1, 128, capi.GEMV_PERSIST, {"XBIT_W4P_FINE": "0"}),          # integer block math, tile aligned, CTAs without work
    (1024, 4768, 2, 128, capi.GEMV_PERSIST, {"XBIT_W4P_FINE": "1"}),         # block granular: tiles shared between CTAs
    (2048, 512, 3, 64, capi.GEMV_PERSIST, {"XBIT_W4P_FINE": "1"}),           # fp16 block math
    (512, 4736, 1, 32, capi.GEMV_PERSIST, {"XBIT_W4P_FINE": "0"}),
    (1024, 512, 1, 128, capi.GEMV_MMA, {"XBIT_GEMV_STREAMK": "0"}),          # cluster split-K
    (512, 4800, 9, 128, capi.GEMV_MMA, {"XBIT_GEMV_STREAMK": "0"}),          # no cluster, two MMA tiles
    (2048, 1024, 2, 128, capi.GEMV_MMA, {"XBIT_GEMV_STREAMK": "1"}),         # persistent stream-K of round 1
    (1024, 512, 1, 128, capi.GEMV_SIMT, {}),
    (300, 100, 3, 48, capi.GEMV_GENERIC, {}),
]
for (K, N, M, g, fam, env) in cases:
    for k, v in env.items():
        capi.set_option(k, int(v))
    qw, s, qz = rand(K, N, 4, g)
    a = torch.randn((M, K), device=dev, generator=gen).to(torch.float16)
    y = X.gemv(a, qw, s, qz, g, 4, K, 1, family=fam)
    check(y, a, X.dequant(qw, s, qz, g, 4, K, 1), f"gemv family={fam} {K}x{N} M={M} g={g} {env}")
    for k in env:
        capi.set_option(k)

# several matrices in one launch
K = 1024
projs = [rand(K, n, 4, 128) for n in (512, 256, 512)]
a = torch.randn((1, K), device=dev, generator=gen).to(torch.float16)
for (q, s, z), y in zip(projs, X.gemv_multi(a, projs, 128, 4, K, 1)):
    check(y, a, X.dequant(q, s, z, 128, 4, K, 1), f"gemv_multi N={q.shape[1]}")

# world = 1 forms of the multi-GPU entry points
K = N = 1024
M = 2
qw, s, qz = rand(K, N, 4, 128)
s = s * 0.05
a = torch.randn((M, K), device=dev, generator=gen).to(torch.float16)
w = X.dequant(qw, s, qz, 128, 4, K, 1)
st = torch.cuda.current_stream().cuda_stream
y = torch.empty((M, N), dtype=torch.float16, device=dev)
ws = torch.zeros(lib.xbit_gemv_workspace_bytes(M, K, N, 4, 128), dtype=torch.uint8, device=dev)
capi.check(lib.xbit_gemv_f16_peers(a.data_ptr(), qw.data_ptr(), s.data_ptr(), qz.data_ptr(), (ctypes.c_void_p * 1)(y.data_ptr()), 1, M, K, N, 4, 128, 1, N, 0,
                                   ws.data_ptr(), ws.numel(), st))
check(y, a, w, "peers world=1")
flags = torch.zeros(8, dtype=torch.int32, device=dev)
state = torch.zeros(4, dtype=torch.int32, device=dev)
for _ in range(2):
    capi.check(lib.xbit_gemv_f16_peers_signal(a.data_ptr(), qw.data_ptr(), s.data_ptr(), qz.data_ptr(), (ctypes.c_void_p * 1)(y.data_ptr()),
                                              (ctypes.c_void_p * 1)(flags.data_ptr()), state.data_ptr(), 1, 0, M, K, N, 4, 128, 1, N, 0,
                                              capi.GEMV_AUTO | capi.GEMV_FLAG_WAIT_PEERS, st))
capi.check(lib.xbit_peers_wait(flags.data_ptr(), 1, 0, state.data_ptr() + 12, st))
check(y, a, w, "signal world=1")
ll = torch.zeros((2, M, N), dtype=torch.int32, device=dev)
state = torch.zeros(4, dtype=torch.int32, device=dev)
out = torch.empty((M, N), dtype=torch.float16, device=dev)
for call, (src, flag) in enumerate(((a.data_ptr(), 0), (ll[0].data_ptr(), capi.GEMV_FLAG_A_IS_LL))):
    capi.check(lib.xbit_gemv_f16_peers_ll(src, qw.data_ptr(), s.data_ptr(), qz.data_ptr(), (ctypes.c_void_p * 1)(ll[call].data_ptr()), state.data_ptr(), call,
                                          1, 0, M, K, N, 4, 128, 1, N, 0, capi.GEMV_AUTO | flag, st))
capi.check(lib.xbit_ll_unpack_f16(ll[1].data_ptr(), out.data_ptr(), M * N, state.data_ptr(), 2, state.data_ptr() + 12, st))
torch.cuda.synchronize()
y1 = X.gemv(a, qw, s, qz, 128, 4, K, 1)
check(out, y1, w, "flag-in-data chain world=1")
# 8- and 2-bit weights on the persistent kernel (integer block math), both CTA-boundary modes
for bits in (8, 2):
    for fine in (0, 1):
        capi.set_option("XBIT_W4P_FINE", fine)
        q8, s8, z8 = rand(1024, 512, bits, 128)
        a8 = torch.randn((2, 1024), device=dev, generator=gen).to(torch.float16)
        check(X.gemv(a8, q8, s8, z8, 128, bits, 1024, 1), a8, X.dequant(q8, s8, z8, 128, bits, 1024, 1), f"gemv bits={bits} fine={fine}")
    capi.set_option("XBIT_W4P_FINE")
# bf16-native forms
X.set_native_bf16(True)
try:
    q4, s4, z4 = rand(1024, 512, 4, 128)
    sb = s4.to(torch.bfloat16)
    ab = torch.randn((2, 1024), device=dev, generator=gen).to(torch.bfloat16)
    wb = X.dequant(q4, sb, z4, 128, 4, 1024, 1)
    check(X.gemv(ab, q4, sb, z4, 128, 4, 1024, 1), ab, wb, "bf16-native gemv / dequant")
finally:
    X.set_native_bf16(False)
print("all sanitizer cases ran", flush=True)
