"""Developer tool: per-CTA phase timeline of the W4 GEMV inside a CUDA graph of back-to-back calls
(XBIT_GEMV_TRACE, globaltimer stamps).  Answers "where do the microseconds of a small shape go".
    python tools/trace.py [K N]...     env knobs of tools/sweep.py apply (XBIT_GEMV_SPLITS, _WC, _RING)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xbitops_b200 import capi  # noqa: E402
from sweep import make, PEAK, WS  # noqa: E402

dev = torch.device("cuda:0")
lib = capi.load()
NAMES = ["cta start", "producer: first TMA issue", "consumers past griddep wait", "activations staged",
         "first stage landed", "last block consumed", "cluster reduce done", "results stored", "", "", "loop entered (before the first wait)"]
if os.environ.get("TRACE_EXTRA"):      # library built with -DW4P_TRACE_EXTRA: slots 8 / 9 / 11 are stamps inside the tile epilogue
    NAMES[8:10] = ["  partial tile written", "  slices found (ballot)"]
    NAMES.append("  handshake done")


def run(K, N, fam, pdl=True, calls=48):
    R, qw, sc, qz, a, out, nbytes = make(K, N, R=max(2, min(64, (300 << 20) // (K * N // 2) + 1)), M=1)
    trace = torch.zeros((64, 1024, 16), dtype=torch.int64, device=dev)
    os.environ["XBIT_GEMV_TRACE"] = hex(trace.data_ptr())
    flags = capi.GEMV_FLAG_STATIC_WEIGHTS if pdl else 0
    order = []

    def fn(i):
        j = i % R
        rc = lib.xbit_gemv_f16_ex(a.data_ptr(), qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(), out[j].data_ptr(),
                                  1, K, N, 4, 128, 0, N, WS.data_ptr(), WS.numel(), fam | flags, torch.cuda.current_stream().cuda_stream)
        assert rc == 0, capi.last_error()

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(0)                      # slot k of the trace buffer <- launch number k (mod 64): count launches
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(calls):
            fn(i)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    trace.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / calls
    del os.environ["XBIT_GEMV_TRACE"]
    t = trace.cpu().numpy().astype(np.int64)
    if t[:, :, 13].max() > 0:
        # persistent kernel: slots 0..7 are SM clock values; 13 / 14 = %globaltimer at CTA start / at the dump, 15 = the
        # clock at the dump.  One ns-per-clock rate for the run (median over the CTAs), anchored at each CTA's start.
        live = (t[:, :, 13] > 0) & (t[:, :, 15] > t[:, :, 0])
        rate = float(np.median((t[:, :, 14] - t[:, :, 13])[live] / (t[:, :, 15] - t[:, :, 0])[live]))
        print(f"   (SM clock {1e3 / rate:.0f} MHz from the stamps)")
        c0 = t[:, :, 0].copy()
        for k in (0, 1, 2, 3, 4, 5, 6, 7, 10) + ((8, 9, 11) if os.environ.get('TRACE_EXTRA') else ()):
            has = live & (t[:, :, k] > 0)
            t[:, :, k] = np.where(has, t[:, :, 13] + ((t[:, :, k] - c0) * rate).astype(np.int64), 0)
    # launch l (0 = the eager call) used slot l % 64; graph launches are 1..calls -> slots (1..calls) % 64
    slots = [(1 + i) % 64 for i in range(calls)][-40:]        # the last 40 graph launches (not overwritten)
    print(f"== {K}x{N} family {fam} pdl={int(pdl)}: {us:.2f} us/call by events ({nbytes/us/1e3/PEAK*100:.0f}% of peak), roofline {nbytes/PEAK/1e3:.2f} us")
    starts, rows = [], []
    for sl in slots[4:-2]:
        tt = t[sl]
        live = tt[:, 0] > 0
        if not live.any():
            continue
        t0 = tt[live, 0].min()
        starts.append(t0)
        rel = np.where(tt[live] > 0, tt[live] - t0, -1)
        rows.append(rel)
    starts = np.array(starts)
    d = np.diff(starts)
    print(f"   launch-to-launch (first CTA start): median {np.median(d)/1e3:.2f} us  min {d.min()/1e3:.2f}  max {d.max()/1e3:.2f}; ctas {rows[0].shape[0]}")
    loop = np.concatenate([t[sl][:, 8][t[sl][:, 0] > 0] for sl in slots[4:-2]]).astype(float)
    cwait = np.concatenate([t[sl][:, 9][t[sl][:, 0] > 0] for sl in slots[4:-2]]).astype(float)
    pwait = np.concatenate([t[sl][:, 10][t[sl][:, 0] > 0] for sl in slots[4:-2]]).astype(float)
    nt = np.concatenate([t[sl][:, 11][t[sl][:, 0] > 0] for sl in slots[4:-2]]).astype(float)
    per = loop / np.maximum(nt, 1)
    print(f"   clk/stage percentiles 5/25/50/75/95: " + " ".join(f"{np.percentile(per, q):.0f}" for q in (5, 25, 50, 75, 95)))
    print(f"   consumer warp 0: loop {np.median(loop):.0f} clk for {np.median(nt):.0f} stages ({np.median(loop / np.maximum(nt, 1)):.0f} clk/stage), "
          f"waiting for data {100 * cwait.sum() / loop.sum():.0f}% of it; (slot 10: {np.median(pwait):.0f})")
    # CTAs per SM within one launch (slot 12 = smid + 1, persistent kernel only)
    if t[slots[10]][:, 12].max() > 0:
        import collections
        hist = collections.Counter()
        for sl in slots[4:-2]:
            sm = t[sl][:, 12]
            sm = sm[sm > 0] - 1
            per = collections.Counter(sm.tolist())
            hist.update(collections.Counter(per.values()))
        print(f"   CTAs of one launch per SM (histogram over launches): {dict(sorted(hist.items()))}")
        # do slow CTAs share an SM with another CTA of the same launch?
        slow, fast = [], []
        for sl in slots[4:-2]:
            tt = t[sl]
            live = tt[:, 12] > 0
            sm = tt[live, 12]
            cnt = collections.Counter(sm.tolist())
            dur = (tt[live, 5] - tt[live, 3]).astype(float)
            for s_, d_ in zip(sm.tolist(), dur.tolist()):
                (slow if cnt[s_] > 1 else fast).append(d_)
        if slow:
            print(f"   staged -> consumed: alone on the SM median {np.median(fast)/1e3:.2f} us (n={len(fast)}), sharing it {np.median(slow)/1e3:.2f} us (n={len(slow)})")
        else:
            print(f"   staged -> consumed: alone on the SM median {np.median(fast)/1e3:.2f} us p95 {np.percentile(fast, 95)/1e3:.2f} (n={len(fast)}); no SM ever holds two CTAs of a launch")
    for k, name in enumerate(NAMES):
        if not name:
            continue
        med, mx, mn = [], [], []
        for rel in rows:
            v = rel[:, k]
            v = v[v >= 0]
            if v.size:
                med.append(np.median(v)); mx.append(v.max()); mn.append(v.min())
        if med:
            print(f"   {name:30s} min {np.median(mn)/1e3:6.2f}  median {np.median(med)/1e3:6.2f}  max {np.median(mx)/1e3:6.2f} us after the launch's first CTA")
    del qw, sc, qz, out


def main():
    fam = int(os.environ.get("TRACE_FAMILY", capi.GEMV_MMA))
    shapes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)] or \
        [(4096, 4096), (4096, 11008), (11008, 4096), (8192, 8192)]
    for K, N in shapes:
        run(K, N, fam, pdl=True)
        if os.environ.get("TRACE_NO_PDL"):
            run(K, N, fam, pdl=False)


if __name__ == "__main__":
    main()
