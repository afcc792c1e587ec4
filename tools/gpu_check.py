"""Developer GPU check (not a test, not the bench): runs the CUDA path against the oracle on many
configurations and prints a compact report + rough timings.  Usage on a B200 box:
    python tools/gpu_check.py [--quick] > gpurun_out/check.log
Uses oracle/ as the checker only."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xbitops_b200 as X  # noqa: E402
from xbitops_b200 import capi, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

dev = torch.device("cuda:0")
co = O.COracle()


def to_dev(*arrs):
    return [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in arrs]


def dq_case(K, N, bits, g, bias, mode="gptq", seed=0):
    qw, s, qz, a = synth.make_inputs(K, N, bits, g, seed=seed, scale_mode=mode)
    want = co.dequant(qw, s, qz, g, bits, K, bias)
    tq, ts, tz = to_dev(qw, s.view(np.int16), qz)
    got = X.dequant(tq, ts.view(torch.float16), tz, g, bits, K, bias)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    bad = int((got.view(np.uint16) != want.view(np.uint16)).sum())
    return bad, got.size


def gemv_err(y, y64):
    y = y.astype(np.float64)
    denom = np.abs(y64).max() + 1e-30
    return float(np.abs(y - y64).max() / denom)


def gemv_case(M, K, N, bits, g, bias, family, seed=0):
    qw, s, qz, a = synth.make_inputs(K, N, bits, g, M=M, seed=seed)
    w = co.dequant(qw, s, qz, g, bits, K, bias)
    y64 = a.astype(np.float64) @ w.astype(np.float64)
    tq, ts, tz, ta = to_dev(qw, s.view(np.int16), qz, a.view(np.int16))
    got = X.gemv(ta.view(torch.float16), tq, ts.view(torch.float16), tz, g, bits, K, bias, family=family)
    torch.cuda.synchronize()
    return gemv_err(got.cpu().numpy(), y64)


def time_graph(fn, calls, reps=20, warm=5):
    """fn(i) enqueues call i on the current stream; all `calls` captured in one graph."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(min(calls, 3)):
            fn(i)
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
        ts.append(e0.elapsed_time(e1) * 1e3 / calls)   # us per call
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-time", action="store_true")
    args = ap.parse_args()
    print("device:", torch.cuda.get_device_name(0), "lib version", capi.load().xbit_version(), flush=True)
    fails = 0

    print("== DQ block kernel vs oracle (bit-exact)")
    for bits in range(2, 9):
        for g in (32, 64, 128):
            for (K, N) in ((256, 64), (416, 136), (1024, 512)):
                for bias in (0, 1):
                    mode = "bits" if (bias and g == 64) else "gptq"
                    bad, tot = dq_case(K, N, bits, g, bias, mode, seed=bits)
                    if bad:
                        fails += 1
                        print(f"  DQ MISMATCH bits={bits} g={g} K={K} N={N} bias={bias} {mode}: {bad}/{tot}")
    print("== DQ element kernel (N%8!=0, g%32!=0)")
    for bits in range(2, 9):
        for (K, N, g) in ((100, 12, 48), (333, 34, 16), (64, 8, 48)):
            bad, tot = dq_case(K, N, bits, g, 1, "gptq", seed=3)
            if bad:
                fails += 1
                print(f"  DQ-elem MISMATCH bits={bits} g={g} K={K} N={N}: {bad}/{tot}")
    print("   dq done, fails so far:", fails, flush=True)

    print("== GEMV vs fp64 truth (normalised max error; bar 1e-2)")
    worst = {}
    for fam, name, Ms in ((capi.GEMV_SIMT, "simt", (1, 3)), (capi.GEMV_MMA, "mma", (1, 2, 5, 8, 9, 16)),
                          (capi.GEMV_GENERIC, "generic", (1, 3, 5))):
        for M in Ms:
            for (K, N, g) in ((4096, 4096, 128), (1024, 256, 32), (11008, 512, 128), (2048, 8192 + 64, 64), (4096, 96, 128), (384, 32, 64)):
                for bias in (0, 1):
                    try:
                        e = gemv_case(M, K, N, 4, g, bias, fam, seed=M)
                    except Exception as ex:  # noqa: BLE001
                        e = float("inf")
                        print(f"  GEMV {name} M={M} K={K} N={N} g={g} EXC {ex}")
                    worst[name] = max(worst.get(name, 0), e)
                    if not (e < 1e-2):
                        fails += 1
                        print(f"  GEMV {name} M={M} K={K} N={N} g={g} bias={bias} err={e:.3e}")
        print(f"   {name}: worst normalised err {worst.get(name):.3e}", flush=True)
    for bits in (2, 3, 5, 6, 7, 8):
        for (M, K, N, g) in ((1, 1024, 256, 128), (2, 777, 100, 48), (17, 512, 64, 32)):
            e = gemv_case(M, K, N, bits, g, 1, capi.GEMV_AUTO, seed=bits)
            if not (e < 1e-2):
                fails += 1
                print(f"  GEMV auto bits={bits} M={M} K={K} N={N} g={g} err={e:.3e}")
    e = gemv_case(40, 4096, 4096, 4, 128, 0, capi.GEMV_AUTO)
    print("   M=40 auto err", e)
    fails += 0 if e < 1e-2 else 1

    print("== reference GPU extension parity")
    ref = O.load_ref_gpu()
    if ref is None:
        print("   reference extension not built: skipped")
    else:
        for bits in range(2, 9):
            K, N, g = 1024, 512, 128
            qw, s, qz, a = synth.make_inputs(K, N, bits, g, seed=11)
            tq, ts, tz = to_dev(qw, s.view(np.int16), qz)
            ts = ts.view(torch.float16)
            for bias in (0, 1):
                r = ref.dequant(tq, ts, tz, g, bits, K, bias)
                m = X.dequant(tq, ts, tz, g, bits, K, bias)
                torch.cuda.synchronize()
                rows = 32 if bits == 6 else K     # reference b=6 is wrong beyond row 31
                bad = int((r[:rows].view(torch.int16) != m[:rows].view(torch.int16)).sum())
                bad_all = int((r.view(torch.int16) != m.view(torch.int16)).sum())
                print(f"   ref DQ bits={bits} bias={bias}: mismatches rows<{rows}: {bad}; all rows: {bad_all}")
                if bad:
                    fails += 1
        for (K, N) in ((4096, 4096), (4096, 11008), (11008, 4096)):
            qw, s, qz, a = synth.make_inputs(K, N, 4, 128, seed=5)
            tq, ts, tz, ta = to_dev(qw, s.view(np.int16), qz, a.view(np.int16))
            ts, ta = ts.view(torch.float16), ta.view(torch.float16)
            w = co.dequant(qw, s, qz, 128, 4, K, 0)
            y64 = a.astype(np.float64) @ w.astype(np.float64)
            r = ref.gemv(ta, tq, ts, tz, 128, 4, K, 0)
            torch.cuda.synchronize()
            m1 = X.gemv(ta, tq, ts, tz, 128, 4, K, 0, family=capi.GEMV_SIMT)
            m2 = X.gemv(ta, tq, ts, tz, 128, 4, K, 0, family=capi.GEMV_MMA)
            torch.cuda.synchronize()
            print(f"   GEMV {K}x{N}: ref err {gemv_err(r.cpu().numpy(), y64):.2e}  simt {gemv_err(m1.cpu().numpy(), y64):.2e}"
                  f"  mma {gemv_err(m2.cpu().numpy(), y64):.2e}")

    if not args.no_time:
        print("== timings (us/call: median, min) ; rotating weights > L2 unless noted")
        lib = capi.load()
        WS = torch.zeros(max(256, lib.xbit_gemv_workspace_bytes(16, 0, 0, 4, 128)), dtype=torch.uint8, device=dev)
        peak = 6549.8
        shapes = [(4096, 4096), (4096, 11008), (11008, 4096)] + ([] if args.quick else [(8192, 8192), (8192, 28672), (28672, 8192)])
        for (K, N) in shapes:
            g, bits = 128, 4
            nbytes = synth.gemv_bytes(K, N, bits, g)
            R = max(2, min(256, (1 << 30) // nbytes + 1))
            G = K // g
            qw = torch.randint(-2**31, 2**31 - 1, (R, K // 8, N), dtype=torch.int32, device=dev)
            sc = (torch.rand((R, G, N), device=dev) * 0.018 + 0.002).to(torch.float16)
            qz = torch.randint(-2**31, 2**31 - 1, (R, G, N // 8), dtype=torch.int32, device=dev)
            a = torch.randn((16, K), device=dev, dtype=torch.float16)
            out = torch.empty((R, 16, N), device=dev, dtype=torch.float16)
            line = f"   {K}x{N} ({nbytes/1e6:.1f} MB, R={R}):"
            for fam, name in ((capi.GEMV_SIMT, "simt"), (capi.GEMV_MMA, "mma")):
                for flags, fl in ((0, ""), (capi.GEMV_FLAG_STATIC_WEIGHTS, "+pdl")):
                    def fn(i, fam=fam, flags=flags):
                        j = i % R
                        st = torch.cuda.current_stream().cuda_stream
                        rc = lib.xbit_gemv_f16_ex(a.data_ptr(), qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(),
                                                  out[j].data_ptr(), 1, K, N, bits, g, 0, N, WS.data_ptr(), WS.numel(), fam | flags, st)
                        assert rc == 0, capi.last_error()
                    med, mn = time_graph(fn, R)
                    line += f"  {name}{fl} {med:.2f}/{mn:.2f}us ({nbytes/med/1e3/peak*100:.0f}%)"
            print(line, flush=True)
            if ref is not None:
                def fnr(i):
                    j = i % R
                    ref.gemv(a[:1], qw[j], sc[j], qz[j], g, bits, K, 0)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for i in range(R):
                    fnr(i)
                torch.cuda.synchronize()
                e0.record(torch.cuda.default_stream())
                for i in range(R):
                    fnr(i)
                e1.record(torch.cuda.default_stream())
                torch.cuda.synchronize()
                print(f"      reference gemv (eager, legacy stream, incl. at::zeros): {e0.elapsed_time(e1)*1e3/R:.2f} us/call")
            del qw, sc, qz, out
        # skinny M sweep on 8192x8192
        if not args.quick:
            K = N = 8192
            g, bits = 128, 4
            nbytes1 = synth.gemv_bytes(K, N, bits, g)
            R = (1 << 30) // nbytes1 + 1
            G = K // g
            qw = torch.randint(-2**31, 2**31 - 1, (R, K // 8, N), dtype=torch.int32, device=dev)
            sc = (torch.rand((R, G, N), device=dev) * 0.018 + 0.002).to(torch.float16)
            qz = torch.randint(-2**31, 2**31 - 1, (R, G, N // 8), dtype=torch.int32, device=dev)
            a = torch.randn((16, K), device=dev, dtype=torch.float16)
            out = torch.empty((R, 16, N), device=dev, dtype=torch.float16)
            for M in (1, 2, 4, 8, 16):
                line = f"   skinny 8192x8192 M={M}:"
                for fam, name in ((capi.GEMV_SIMT, "simt"), (capi.GEMV_MMA, "mma")):
                    if fam == capi.GEMV_SIMT and M > 2:
                        continue
                    def fn(i, fam=fam):
                        j = i % R
                        st = torch.cuda.current_stream().cuda_stream
                        rc = lib.xbit_gemv_f16_ex(a.data_ptr(), qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(),
                                                  out[j].data_ptr(), M, K, N, bits, g, 0, N, WS.data_ptr(), WS.numel(),
                                                  fam | capi.GEMV_FLAG_STATIC_WEIGHTS, st)
                        assert rc == 0, capi.last_error()
                    med, mn = time_graph(fn, R)
                    nb = synth.gemv_bytes(K, N, bits, g, M)
                    line += f"  {name} {med:.2f}/{mn:.2f}us ({nb/med/1e3/peak*100:.0f}%)"
                print(line, flush=True)
            del qw, sc, qz, out
        # DQ sweep
        K, N = 4096, 11008
        for bits in range(2, 9):
            for g in ((128,) if args.quick else (32, 64, 128)):
                G = K // g
                qrows = (K * bits + 31) // 32
                zw = (N * bits + 31) // 32
                nbytes = synth.dq_bytes(K, N, bits, g)
                R = 12
                qw = torch.randint(-2**31, 2**31 - 1, (R, qrows, N), dtype=torch.int32, device=dev)
                sc = (torch.rand((R, G, N), device=dev) * 0.018 + 0.002).to(torch.float16)
                qz = torch.randint(-2**31, 2**31 - 1, (R, G, zw), dtype=torch.int32, device=dev)
                out = torch.empty((R, K, N), device=dev, dtype=torch.float16)
                def fn(i):
                    j = i % R
                    st = torch.cuda.current_stream().cuda_stream
                    rc = lib.xbit_dequant_f16(qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(), out[j].data_ptr(),
                                              K, N, bits, g, 0, st)
                    assert rc == 0, capi.last_error()
                med, mn = time_graph(fn, R, reps=10)
                print(f"   DQ 4096x11008 bits={bits} g={g}: {med:.2f}/{mn:.2f} us  {nbytes/med/1e3:.0f} GB/s ({nbytes/med/1e3/peak*100:.0f}%)", flush=True)
                del qw, sc, qz, out
    print("TOTAL FAILS:", fails)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
