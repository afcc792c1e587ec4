#!/usr/bin/env python
"""bench.py -- the driver contract for the XbitOps hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

metric   a16w4_g128_gemv_effective_bandwidth, GB/s = algorithmic bytes / time;
         us/call per shape rides along in `per_shape` (BASELINE.json: "A16W4 GEMV us/call & achieved HBM GB/s").
step     one pass over the workload's rotating set: for every shape, R distinct weight sets with
         R * bytes >= 1 GiB (>> the 126 MB L2, so every call streams from HBM), one GEMV call per set,
         all captured in one CUDA graph (steady-state decode: programmatic dependent launch lets call
         n+1 prefetch weights while call n drains).
N = 1    workload llama2-7b  = BASELINE.json configs[1]: 4096x4096, 4096x11008, 11008x4096, batch 1.
N > 1    workload llama2-70b = configs[3]: 8192x8192, 8192x28672, 28672x8192 N-split over the ranks,
         strong scaling.  --combine says how the output slices are exchanged: `ll` (default) = the
         step ordered as ONE dependent chain (8192x8192 -> 8192x28672 -> 28672x8192 -> ...), every
         call's epilogue stores {results, call number} slots into every rank's buffer over NVLink and
         the next call spins on the slots it needs while staging its activations; `peers` = fused
         peer stores + a symmetric-memory barrier per call; `signal` = peer stores + completion
         flags awaited inside the next call; `nccl` = all_gather_into_tensor per call; `none` =
         kernel only.  per_shape carries the other modes for comparison; rank 0 also times the
         unsharded workload alone and reports it as `single_gpu_same_workload` for context.
e2e      (N = 1) the same calls through xbit_gemv_f16_host with pinned host buffers: activations
         pulled over PCIe and results stored into host memory inside every call, replayed from a
         CUDA graph, host wall clock including one sync per step.
--impl reference   the reference's CPU implementation of the path (oracle/_ref: the unmodified
         src/cpp_simulate.cc dequant, N-sliced over all host threads, then the oracle's dot) on a
         bounded sample (one call per shape).  This is the only place besides cpu_baseline where
         oracle/ is executed, and only as the thing the baseline arm measures.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "llama2-7b": [(4096, 4096), (4096, 11008), (11008, 4096)],
    "llama2-70b": [(8192, 8192), (8192, 28672), (28672, 8192)],
}
BITS, GROUP = 4, 128
METRIC = "a16w4_g128_gemv_effective_bandwidth"
UNIT = "GB/s"
MIN_ROTATE_BYTES = 1 << 30


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- clocks

class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (pynvml, 20 ms period)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------- reference / CPU arms

def cpu_reference_run(shapes, threads: int):
    """One call per shape of the reference CPU path: unmodified cpp_simulate.cc dequant (oracle/_ref)
    N-sliced over `threads` host threads, then the oracle's fp64 dot over the dequantised slice.
    Falls back to the oracle port for the dequant when oracle/_ref was not built.
    Returns (seconds, bytes, kind)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    from xbitops_b200 import synth
    co = O.COracle()
    rc = O.RefCpu() if O.RefCpu.available() else None
    kind = "reference" if rc is not None else "port"
    total_s, total_b = 0.0, 0
    for (K, N) in shapes:
        qw, s, qz, a = synth.make_inputs(K, N, BITS, GROUP, seed=0)
        t = max(1, min(threads, N // 64))
        while N % (t * 8):
            t -= 1
        width = N // t
        slices = []
        for i in range(t):
            sl = slice(i * width, (i + 1) * width)
            slices.append((np.ascontiguousarray(qw[:, sl]), np.ascontiguousarray(s[:, sl]),
                           np.ascontiguousarray(qz[:, i * width // 8:(i + 1) * width // 8])))

        def work(args):
            q, sc, z = args
            w = rc.dequant(q, sc, z, GROUP, BITS, K) if rc is not None else co.dequant(q, sc, z, GROUP, BITS, K, 0)
            return co.gemv_from_dq(a, w)[1]

        t0 = time.perf_counter()
        if t == 1:
            work(slices[0])
        else:
            with ThreadPoolExecutor(t) as ex:
                list(ex.map(work, slices))
        total_s += time.perf_counter() - t0
        total_b += synth.gemv_bytes(K, N, BITS, GROUP)
    return total_s, total_b, kind


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    shapes = WORKLOADS[args.workload or ("llama2-7b" if args.gpus == 1 else "llama2-70b")]
    cores = os.cpu_count() or 1
    steps, warm = max(1, min(args.steps, 5)), max(0, min(args.warmup, 1))
    for _ in range(warm):
        cpu_reference_run(shapes, cores)
    secs, nbytes, kind = 0.0, 0, "port"
    for _ in range(steps):
        s, b, kind = cpu_reference_run(shapes, cores)
        secs += s
        nbytes += b
    val = nbytes / secs / 1e9
    sample = f"one call per shape of {shapes} per step; {steps} steps"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": secs / steps * 1e3, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": args.workload or ("llama2-7b" if args.gpus == 1 else "llama2-70b"),
                       "shapes": shapes, "bits": BITS, "groupsize": GROUP, "batch": 1,
                       "note": "reference CPU simulator (src/cpp_simulate.cc) dequant + fp64 dot, N-sliced over host threads"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- GPU arm

class ShapeSet:
    """R rotating weight sets of one (K, N_local) shape, generated on the device."""

    def __init__(self, torch, dev, K, N_total, world, rank, gen):
        from xbitops_b200 import synth
        self.K, self.N_total, self.world, self.rank = K, N_total, world, rank
        self.N = N_total // world
        assert N_total % (world * 64) == 0
        self.bytes_call = synth.gemv_bytes(K, N_total, BITS, GROUP)          # unsharded algorithmic bytes
        self.bytes_rank = synth.gemv_bytes(K, self.N, BITS, GROUP)
        self.R = max(2, (MIN_ROTATE_BYTES + self.bytes_rank - 1) // self.bytes_rank)
        G = K // GROUP
        self.qw = torch.randint(-2**31, 2**31 - 1, (self.R, K * BITS // 32, self.N), dtype=torch.int32, device=dev, generator=gen)
        self.sc = (torch.rand((self.R, G, self.N), device=dev, generator=gen) * 0.018 + 0.002).to(torch.float16)
        self.qz = torch.randint(-2**31, 2**31 - 1, (self.R, G, self.N * BITS // 32), dtype=torch.int32, device=dev, generator=gen)
        self.a = torch.randn((1, K), device=dev, generator=gen).to(torch.float16)
        self.out = torch.zeros((self.R, 1, N_total), device=dev, dtype=torch.float16)
        self.col0 = rank * self.N
        self.symm = None          # (buffer, handle, peer base pointers) for the fused peer-store epilogue

    def make_symmetric(self, torch, dist):
        import torch.distributed._symmetric_memory as symm_mem
        buf = symm_mem.empty((self.R, 1, self.N_total), dtype=torch.float16, device=self.out.device)
        hdl = symm_mem.rendezvous(buf, dist.group.WORLD.group_name)
        self.symm = (buf, hdl, [int(p) for p in hdl.buffer_ptrs])


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from xbitops_b200 import capi, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: xbitops_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = capi.load()
    ws_bytes = max(256, lib.xbit_gemv_workspace_bytes(1, 0, 0, BITS, GROUP)) if not args.no_streamk else 0
    ws = torch.zeros(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
    ws_ptr, ws_len = (ws.data_ptr(), ws_bytes) if ws_bytes else (None, 0)
    workload = args.workload or ("llama2-7b" if world == 1 else "llama2-70b")
    shapes = WORKLOADS[workload]
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    sets = [ShapeSet(torch, dev, K, N, world, rank, gen) for (K, N) in shapes]
    family = {"auto": capi.GEMV_AUTO, "simt": capi.GEMV_SIMT, "mma": capi.GEMV_MMA}[args.family]
    flags = 0 if args.no_pdl else capi.GEMV_FLAG_STATIC_WEIGHTS
    peak, peak_src = measured_peak_gbs()

    combine = args.combine if world > 1 else "none"
    sig = None
    llctx = None
    if world > 1:
        for ss in sets:                       # replicated activations (the N-split shards columns only)
            dist.broadcast(ss.a, 0)
    if combine == "ll":
        try:
            import torch.distributed._symmetric_memory as symm_mem
            max_n = max(ss.N_total for ss in sets)
            llbuf = symm_mem.empty((8, max_n), dtype=torch.int32, device=dev)     # rotating LL buffers, 8 bytes per result pair
            llbuf.zero_()
            lhdl = symm_mem.rendezvous(llbuf, dist.group.WORLD.group_name)
            lstate = torch.zeros(4, dtype=torch.int32, device=dev)
            lplain = torch.empty((max_n,), dtype=torch.float16, device=dev)
            torch.cuda.synchronize()
            dist.barrier()
            llctx = {"buf": llbuf, "ptrs": [int(p) for p in lhdl.buffer_ptrs], "state": lstate, "plain": lplain,
                     "stride": max_n * 4, "calls": 0, "last_n": 0, "mod": 2}
        except Exception as ex:  # noqa: BLE001
            if rank == 0:
                print(f"[bench] symmetric memory unavailable ({ex}); falling back to nccl all-gather", file=sys.stderr)
            combine = "nccl"
    if world > 1:
        try:
            for ss in sets:                   # peer-mapped result buffers: the peers / signal forms, and the plain-output figures beside ll
                ss.make_symmetric(torch, dist)
            if True:
                import torch.distributed._symmetric_memory as symm_mem
                sflags = symm_mem.empty((64,), dtype=torch.int32, device=dev)
                sflags.zero_()
                fhdl = symm_mem.rendezvous(sflags, dist.group.WORLD.group_name)
                state = torch.zeros(4, dtype=torch.int32, device=dev)
                torch.cuda.synchronize()
                dist.barrier()
                sig = (sflags, [int(p) for p in fhdl.buffer_ptrs], state)
        except Exception as ex:  # noqa: BLE001
            if rank == 0:
                print(f"[bench] symmetric memory unavailable ({ex}); falling back to nccl all-gather", file=sys.stderr)
            if combine in ("peers", "signal"):
                combine = "nccl"
            sig = None
            for ss in sets:
                ss.symm = None

    def launch(ss: ShapeSet, j: int, mode: str = None):
        mode = mode or combine
        st = torch.cuda.current_stream().cuda_stream
        if mode == "peers":
            buf, hdl, bases = ss.symm
            off = j * ss.N_total * 2
            arr = (ctypes.c_void_p * world)(*[b + off for b in bases])
            rc = lib.xbit_gemv_f16_peers_ex(ss.a.data_ptr(), ss.qw[j].data_ptr(), ss.sc[j].data_ptr(), ss.qz[j].data_ptr(),
                                            arr, world, 1, ss.K, ss.N, BITS, GROUP, 0, ss.N_total, ss.col0, ws_ptr, ws_len,
                                            family | flags, st)
            if rc != 0:
                raise RuntimeError(capi.last_error())
            hdl.barrier()          # every rank's slice has landed in every buffer
            return
        if mode == "ll":
            # dependent chain: this call's activations are the previous call's gathered result, read from the LL
            # buffer slot by slot as the ranks deliver them; its own result goes into the other LL buffer
            c = llctx["calls"]
            chained = c > 0 and llctx["last_n"] == ss.K
            mod = llctx["mod"]
            src = (llctx["ptrs"][rank] + ((c - 1) % mod) * llctx["stride"]) if chained else ss.a.data_ptr()
            outs = (ctypes.c_void_p * world)(*[b + (c % mod) * llctx["stride"] for b in llctx["ptrs"]])
            rc = lib.xbit_gemv_f16_peers_ll(src, ss.qw[j].data_ptr(), ss.sc[j].data_ptr(), ss.qz[j].data_ptr(), outs,
                                            llctx["state"].data_ptr(), c, world, rank, 1, ss.K, ss.N, BITS, GROUP, 0, ss.N_total,
                                            ss.col0, family | flags | (capi.GEMV_FLAG_A_IS_LL if chained else 0), st)
            if rc != 0:
                raise RuntimeError(capi.last_error())
            if not chained and c > 0:
                raise RuntimeError("LL chain broken: the bench order must feed every call from the previous one")
            llctx["calls"], llctx["last_n"] = c + 1, ss.N_total
            return
        if mode == "signal":
            buf, hdl, bases = ss.symm
            off = j * ss.N_total * 2
            arr = (ctypes.c_void_p * world)(*[b + off for b in bases])
            sflags, fptrs, state = sig
            farr = (ctypes.c_void_p * world)(*fptrs)
            # the gather of call i is awaited inside call i + 1 (as in a decode chain, where call i + 1 reads
            # call i's gathered output); finish_chain() awaits the last one
            rc = lib.xbit_gemv_f16_peers_signal(ss.a.data_ptr(), ss.qw[j].data_ptr(), ss.sc[j].data_ptr(), ss.qz[j].data_ptr(),
                                                arr, farr, state.data_ptr(), world, rank, 1, ss.K, ss.N, BITS, GROUP, 0,
                                                ss.N_total, ss.col0, family | flags | (0 if os.environ.get("BENCH_SIGNAL_NOWAIT") else capi.GEMV_FLAG_WAIT_PEERS), st)
            if rc != 0:
                raise RuntimeError(capi.last_error())
            return
        out = ss.out[j]
        rc = lib.xbit_gemv_f16_peers_ex(ss.a.data_ptr(), ss.qw[j].data_ptr(), ss.sc[j].data_ptr(), ss.qz[j].data_ptr(),
                                        (ctypes.c_void_p * 1)(out.data_ptr()), 1, 1, ss.K, ss.N, BITS, GROUP, 0,
                                        ss.N_total, ss.col0, ws_ptr, ws_len, family | flags, st)
        if rc != 0:
            raise RuntimeError(capi.last_error())
        if mode == "nccl":
            dist.all_gather_into_tensor(out.view(-1), out[:, ss.col0:ss.col0 + ss.N].reshape(-1))

    def finish_chain(mode: str = None):
        if (mode or combine) == "ll":
            c = llctx["calls"]
            rc = lib.xbit_ll_unpack_f16(llctx["ptrs"][rank] + ((c - 1) % llctx["mod"]) * llctx["stride"], llctx["plain"].data_ptr(),
                                        llctx["last_n"], llctx["state"].data_ptr(), c, llctx["state"].data_ptr() + 12,
                                        torch.cuda.current_stream().cuda_stream)
            if rc != 0:
                raise RuntimeError(capi.last_error())
            llctx["calls"], llctx["last_n"] = 0, 0      # the next chain starts from plain activations
        if (mode or combine) == "signal":
            sflags, fptrs, state = sig
            rc = lib.xbit_peers_wait(fptrs[rank], world, rank, state.data_ptr() + 12, torch.cuda.current_stream().cuda_stream)
            if rc != 0:
                raise RuntimeError(capi.last_error())

    def capture(set_list, mode=None, same_set=False):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm outside capture (module load, NCCL channels)
            for ss in set_list:
                launch(ss, 0, mode)
            finish_chain(mode)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n = 0
        order = [(ss, 0 if same_set else j) for ss in set_list for j in range(ss.R)]   # same_set: one weight set, L2-warm
        if (mode or combine) == "ll" and len(set_list) > 1:
            # every weight set once, ordered as a dependent chain: cycles (8192x8192 -> 8192x28672 -> 28672x8192),
            # whose output width is the next call's K, then the remaining square calls (which chain with themselves)
            cyc = min(ss.R for ss in set_list)
            order = [(ss, j) for j in range(cyc) for ss in set_list]
            order += [(ss, j) for ss in set_list for j in range(cyc, ss.R)]
        if (mode or combine) == "ll":
            # a replayed chain starts again on buffer 0 while a slower rank may still be unpacking the last call's
            # buffer: rotate over m buffers with (len - 1) % m != 0 so that the two never coincide
            llctx["mod"] = next(m for m in range(2, 9) if (len(order) - 1) % m != 0)
        with torch.cuda.graph(g):
            for ss, j in order:
                launch(ss, j, mode)
                n += 1
            finish_chain(mode)
        return g, n

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(graph, steps, warmup):
        for _ in range(warmup):
            graph.replay()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            graph.replay()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- N > 1: correctness before speed.  For every exchange form this line reports, the sharded result of every shape
    # (weight set 0) against the unsharded call on the gathered weights, computed on rank 0 and broadcast.
    selfcheck = None
    if world > 1:
        import xbitops_b200 as X
        selfcheck = {}
        modes = [combine] + [m for m in ("none", "nccl", "peers", "signal") if m != combine and not (m in ("peers", "signal") and sets[0].symm is None)
                             and not (m == "signal" and sig is None)]
        for ss in sets:
            parts = [[torch.empty_like(t_) for _ in range(world)] for t_ in (ss.qw[0], ss.sc[0], ss.qz[0])]
            for lst, t_ in zip(parts, (ss.qw[0], ss.sc[0], ss.qz[0])):
                dist.all_gather(lst, t_.contiguous())
            full = torch.empty((1, ss.N_total), dtype=torch.float16, device=dev)
            if rank == 0:
                fq, fs, fz = (torch.cat(lst, dim=1).contiguous() for lst in parts)
                full.copy_(X.gemv(ss.a, fq, fs, fz, GROUP, BITS, ss.K, 0))
            dist.broadcast(full, 0)
            del parts
            ref_max = float(full.double().abs().max())
            for mode in modes:
                if mode == "ll" and llctx is None:
                    continue
                if mode == "ll":
                    llctx["calls"], llctx["last_n"] = 0, 0
                launch(ss, 0, mode)
                finish_chain(mode)
                torch.cuda.synchronize()
                dist.barrier()
                if mode == "ll":
                    got = llctx["plain"][:ss.N_total].view(1, -1)
                elif mode in ("peers", "signal"):
                    got = ss.symm[0][0]
                else:
                    got = ss.out[0]
                sl = slice(ss.col0, ss.col0 + ss.N) if mode == "none" else slice(0, ss.N_total)
                err = float((got[:, sl].double() - full[:, sl].double()).abs().max()) / ref_max
                key = f"{ss.K}x{ss.N_total}:{mode}"
                selfcheck[key] = err
        worst = torch.tensor([max(selfcheck.values())], device=dev, dtype=torch.float64)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        if float(worst.item()) > 2e-3:
            raise RuntimeError(f"sharded self-check failed on rank {rank}: {selfcheck}")
        selfcheck = {"status": "ok", "max_normalised_error": float(worst.item()), "modes": modes,
                     "how": "sharded call of every shape (weight set 0) in every exchange form vs the unsharded call on the gathered weights, <= 2e-3"}
        if llctx is not None:
            llctx["calls"], llctx["last_n"] = 0, 0

    graph, calls_per_step = capture(sets)
    step_bytes = sum(ss.bytes_call * ss.R for ss in sets)
    with ClockSampler(local_rank) as clk:
        total_ms = timed(graph, args.steps, args.warmup)
    ms_per_step = total_ms / args.steps
    value = step_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- per-shape breakdown (separate graphs; same protocol) -- explains the aggregate
    per_shape = {}
    for ss in sets:
        # (flag-in-data mode needs a dependent chain: only a square shape chains with itself; the others are
        # timed kernel-only here and appear with their exchange in the aggregate)
        # (flag-in-data mode needs a dependent chain, and only a square shape chains with itself: the others are timed here
        # with the plain-[M, N] fused exchange -- peer stores + barrier -- and appear with the ll exchange in the aggregate)
        solo_mode = ("peers" if ss.symm is not None else "nccl") if (combine == "ll" and ss.K != ss.N_total) else None
        g1, n1 = capture([ss], solo_mode)
        ms1 = timed(g1, max(3, args.steps // 4), 3) / max(3, args.steps // 4)
        us = ms1 * 1e3 / n1
        gbs = ss.bytes_call / us / 1e3
        per_shape[f"{ss.K}x{ss.N_total}"] = {
            "us_per_call": round(us, 3), "GBps": round(gbs, 1), "frac_of_measured_peak": round(gbs / (peak * world), 4),
            "frac_of_8TBps_nominal": round(gbs / (8000.0 * world), 4), "algorithmic_bytes": ss.bytes_call,
            "rotating_sets": ss.R, "family": lib.xbit_gemv_pick_family(1, ss.K, ss.N, BITS, GROUP) if family == 0 else family}
        if solo_mode:
            per_shape[f"{ss.K}x{ss.N_total}"]["us_per_call_is"] = f"{solo_mode} exchange, plain [M, N] output (this shape cannot chain with itself in the ll form)"
        if world == 1:
            # labelled aside: the same calls on ONE weight set (it stays in the 126 MB L2) -- not a roofline figure
            gw, nw = capture([ss], None, same_set=True)
            msw = timed(gw, max(3, args.steps // 4), 3) / max(3, args.steps // 4)
            per_shape[f"{ss.K}x{ss.N_total}"]["us_per_call_l2_warm_single_buffer"] = round(msw * 1e3 / nw, 3)
            del gw
        del g1
        if world > 1:
            for mode in ("none", "nccl", "peers", "signal"):
                if mode == (solo_mode or combine) or (mode in ("peers", "signal") and ss.symm is None) or (mode == "signal" and sig is None):
                    continue
                if mode == "signal" and sig is None:
                    continue
                g2, n2 = capture([ss], mode)
                ms2 = timed(g2, max(3, args.steps // 4), 3) / max(3, args.steps // 4)
                per_shape[f"{ss.K}x{ss.N_total}"][f"us_per_call_{'kernel_only' if mode == 'none' else mode}"] = round(ms2 * 1e3 / n2, 3)
                del g2

    if sig is not None and int(sig[2][3].item()) != 0:
        raise RuntimeError("xbit_peers_wait timed out: a rank never published its completion flag")
    if llctx is not None and int(llctx["state"][3].item()) != 0:
        raise RuntimeError("xbit_ll_unpack_f16 timed out: a rank never delivered its slice")
    line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 5), "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": workload, "shapes": shapes, "bits": BITS, "groupsize": GROUP, "batch": 1,
                       "calls_per_step": calls_per_step,
                       "l2_policy": "inputs larger than L2: every call reads a distinct weight set, >= 1 GiB rotated per shape",
                       "launch": "one CUDA graph per step, programmatic dependent launch" + (" off" if args.no_pdl else ""),
                       "schedule": "cluster split-K" if args.no_streamk else "auto: persistent per-SM schedule (per-warp TMA rings, integer block math) where measured ahead, cluster split-K otherwise",
                       "combine": {"ll": "fused, flag-in-data: the kernel's epilogue stores every pair of results into every rank's buffer over NVLink as one 8-byte {results, call number} store; the next call of the dependent chain spins on the slots it needs while staging its activations (no barrier, no fence, no wait launch; one unpack kernel at the end of a step)",
                                   "signal": "fused: the kernel's epilogue stores its slice into every rank's buffer over NVLink and raises a per-rank completion flag; the next call's kernel awaits the flags before it reads its activations (one wait kernel at the end of a step)",
                                   "nccl": "nccl all_gather_into_tensor per call",
                                   "peers": "fused epilogue: NVLink peer stores into every rank's buffer + one symmetric-memory barrier per call",
                                   "none": "none"}[combine],
                       "parallelism": f"n-split x{world}" if world > 1 else "single"},
            "per_shape": per_shape, "selfcheck": selfcheck, "clocks": clk.summary(),
            "gpu_launches": (calls_per_step + (1 if combine in ("signal", "ll") else 0)) * args.steps}

    if rank == 0:
        # roofline of the dominant (only) kernel family over the timed region
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(workload)
            except Exception:  # noqa: BLE001
                traffic = None
        avg_us = ms_per_step * 1e3 / calls_per_step
        line["roofline"] = {"bound": "hbm", "achieved": round(value / world, 2), "peak": peak, "unit": "GB/s",
                            "frac": round(value / world / peak, 4), "traffic": traffic, "peak_source": peak_src,
                            "kernel": "xbit::gemv_w4p_kernel (persistent schedule; xbit::gemv_w4_kernel where AUTO keeps the cluster split-K kernel)", "avg_launch_us": round(avg_us, 3),
                            "algorithmic_bytes_per_launch": round(step_bytes / calls_per_step / world),
                            "frac_of_8TBps_nominal": round(value / world / 8000.0, 4)}

    # ---- the other BASELINE.json configs and the op surface, same protocol (N = 1 only; explain the headline, not part of it)
    if world == 1 and not args.quick:
        import xbitops_b200 as X
        from xbitops_b200 import ops
        reps = max(3, args.steps // 10)

        def graph_us(fn, calls, reps=reps, stream=None):
            side = stream or torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn(0)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for i in range(calls):
                    fn(i)
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e3 / (reps * calls)

        # (1) the operator surface: XbitOps.gemv as a reference user calls it, static weights asserted once
        ops.set_static_weights(True)
        surf = {}
        tot_us = 0.0
        for ss in sets:
            def fn(i, ss=ss):
                j = i % ss.R
                ops.gemv(ss.a, ss.qw[j], ss.sc[j], ss.qz[j], GROUP, BITS, ss.K, 0, out=ss.out[j])
            us = graph_us(fn, ss.R)
            surf[f"{ss.K}x{ss.N_total}"] = round(us, 3)
            tot_us += us * ss.R
        ops.set_static_weights(False)
        surf_gbs = step_bytes / tot_us / 1e3
        line["op_surface"] = {"us_per_call": surf, "value": round(surf_gbs, 2), "unit": UNIT, "vs_value": round(surf_gbs / value, 4),
                              "how": "xbitops_b200.gemv (the reference's op signature) with set_static_weights(True), same rotating sets, one CUDA graph per shape"}

        # (2) multi-projection launches: Q/K/V and gate + up through xbit_gemv_f16_multi
        fused = {}
        for name, ss, P in (("qkv_fused", sets[0], 3), ("gate_up_fused", sets[1], 2)):
            Rm = ss.R - ss.R % P
            probs = []
            for j in range(0, Rm, P):
                arr = (capi.GemvProblem * P)()
                for i in range(P):
                    arr[i] = capi.GemvProblem(ss.qw[j + i].data_ptr(), ss.sc[j + i].data_ptr(), ss.qz[j + i].data_ptr(), ss.out[j + i].data_ptr(), ss.N, ss.N_total)
                probs.append(arr)

            def fn(i, ss=ss, probs=probs, P=P):
                rc = lib.xbit_gemv_f16_multi(ss.a.data_ptr(), ctypes.cast(probs[i % len(probs)], ctypes.c_void_p), P, 1, ss.K, BITS, GROUP, 0,
                                             ws_ptr, ws_len, family | flags, torch.cuda.current_stream().cuda_stream)
                if rc != 0:
                    raise RuntimeError(capi.last_error())
            us = graph_us(fn, len(probs))
            fused[name] = {"matrices": P, "shape": f"{ss.K}x{ss.N_total}", "us_per_launch": round(us, 3), "us_per_matrix": round(us / P, 3),
                           "GBps": round(ss.bytes_call * P / us / 1e3, 1), "frac_of_measured_peak": round(ss.bytes_call * P / us / 1e3 / peak, 4)}
        line["fused"] = fused

        # (3) configs[4]: skinny GEMM M = 1..16 on 8192 x 8192
        sk = ShapeSet(torch, dev, 8192, 8192, 1, 0, gen)
        a16 = torch.randn((16, 8192), device=dev, generator=gen).to(torch.float16)
        o16 = torch.empty((sk.R, 16, 8192), device=dev, dtype=torch.float16)
        skinny = {}
        for M in (1, 2, 4, 8, 16):
            def fn(i, M=M):
                j = i % sk.R
                rc = lib.xbit_gemv_f16_ex(a16.data_ptr(), sk.qw[j].data_ptr(), sk.sc[j].data_ptr(), sk.qz[j].data_ptr(), o16[j].data_ptr(),
                                          M, 8192, 8192, BITS, GROUP, 0, 8192, ws_ptr, ws_len, family | flags, torch.cuda.current_stream().cuda_stream)
                if rc != 0:
                    raise RuntimeError(capi.last_error())
            us = graph_us(fn, sk.R)
            nb = synth.gemv_bytes(8192, 8192, BITS, GROUP, M)
            skinny[str(M)] = {"us_per_call": round(us, 3), "GBps": round(nb / us / 1e3, 1), "frac_of_measured_peak": round(nb / us / 1e3 / peak, 4),
                              "family": lib.xbit_gemv_pick_family(M, 8192, 8192, BITS, GROUP)}
        line["skinny"] = {"shape": "8192x8192", "by_M": skinny}
        del sk, o16

        # (3b) A16W8 (8-bit weights, groupsize 128) on the same shapes: the persistent kernel's integer block math takes
        #      the packed words as MMA operands unchanged (the reference aborts on bits != 4, gemv_w4a16_pt.cu:152-155)
        w8 = {}
        for (K8, N8) in shapes:
            nb8 = synth.gemv_bytes(K8, N8, 8, GROUP, 1)
            R8 = max(2, (1 << 30) // nb8 + 1)
            qw8 = torch.randint(-2**31, 2**31 - 1, (R8, K8 // 4, N8), dtype=torch.int32, device=dev, generator=gen)
            qz8 = torch.randint(-2**31, 2**31 - 1, (R8, K8 // GROUP, N8 // 4), dtype=torch.int32, device=dev, generator=gen)
            sc8 = (torch.rand((R8, K8 // GROUP, N8), device=dev, generator=gen) * 0.018 + 0.002).to(torch.float16)
            a8 = torch.randn((1, K8), device=dev, generator=gen).to(torch.float16)
            o8 = torch.empty((R8, 1, N8), device=dev, dtype=torch.float16)

            def fn(i, K8=K8, N8=N8, R8=R8, qw8=qw8, qz8=qz8, sc8=sc8, a8=a8, o8=o8):
                j = i % R8
                rc = lib.xbit_gemv_f16_ex(a8.data_ptr(), qw8[j].data_ptr(), sc8[j].data_ptr(), qz8[j].data_ptr(), o8[j].data_ptr(),
                                          1, K8, N8, 8, GROUP, 0, N8, ws_ptr, ws_len, capi.GEMV_AUTO | flags, torch.cuda.current_stream().cuda_stream)
                if rc != 0:
                    raise RuntimeError(capi.last_error())
            us = graph_us(fn, R8)
            w8[f"{K8}x{N8}"] = {"us_per_call": round(us, 3), "GBps": round(nb8 / us / 1e3, 1), "frac_of_measured_peak": round(nb8 / us / 1e3 / peak, 4),
                                "family": lib.xbit_gemv_pick_family(1, K8, N8, 8, GROUP)}
            del qw8, qz8, sc8, o8
        line["w8"] = {"bits": 8, "groupsize": GROUP, "per_shape": w8}

        # (3c) bf16-native GEMV (bf16 activations, scales and output, no fp16 round trip) on the headline sets
        bfn = {}
        for ss in sets:
            ab = ss.a.to(torch.bfloat16)
            scb = ss.sc.to(torch.bfloat16)
            ob = torch.empty((ss.R, 1, ss.N_total), device=dev, dtype=torch.bfloat16)

            def fn(i, ss=ss, ab=ab, scb=scb, ob=ob):
                j = i % ss.R
                rc = lib.xbit_gemv_bf16(ab.data_ptr(), ss.qw[j].data_ptr(), scb[j].data_ptr(), ss.qz[j].data_ptr(), ob[j].data_ptr(),
                                        1, ss.K, ss.N, BITS, GROUP, 0, ss.N_total, ws_ptr, ws_len, flags, torch.cuda.current_stream().cuda_stream)
                if rc != 0:
                    raise RuntimeError(capi.last_error())
            us = graph_us(fn, ss.R)
            bfn[f"{ss.K}x{ss.N_total}"] = {"us_per_call": round(us, 3), "GBps": round(ss.bytes_call / us / 1e3, 1),
                                           "frac_of_measured_peak": round(ss.bytes_call / us / 1e3 / peak, 4)}
            del ab, scb, ob
        line["bf16_native"] = {"per_shape": bfn, "how": "xbit_gemv_bf16 on the headline weight sets (scales cast to bf16)"}

        # (4) configs[2]: dequant to fp16, bits 2..8 x group size 32 / 64 / 128 on 4096 x 11008
        Kd, Nd, Rd = 4096, 11008, 3
        outd = torch.empty((Rd, Kd, Nd), device=dev, dtype=torch.float16)
        dq = {}
        for b in range(2, 9):
            for g in (32, 64, 128):
                qw = torch.randint(-2**31, 2**31 - 1, (Rd, (Kd * b + 31) // 32, Nd), dtype=torch.int32, device=dev, generator=gen)
                qz = torch.randint(-2**31, 2**31 - 1, (Rd, Kd // g, (Nd * b + 31) // 32), dtype=torch.int32, device=dev, generator=gen)
                sc = (torch.rand((Rd, Kd // g, Nd), device=dev, generator=gen) * 0.018 + 0.002).to(torch.float16)

                def fn(i, b=b, g=g, qw=qw, qz=qz, sc=sc):
                    j = i % Rd
                    rc = lib.xbit_dequant_f16(qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(), outd[j].data_ptr(), Kd, Nd, b, g, 1,
                                              torch.cuda.current_stream().cuda_stream)
                    if rc != 0:
                        raise RuntimeError(capi.last_error())
                us = graph_us(fn, 2 * Rd)
                nb = synth.dq_bytes(Kd, Nd, b, g)
                dq[f"b{b}_g{g}"] = {"us": round(us, 2), "GBps": round(nb / us / 1e3, 1), "frac_of_measured_peak": round(nb / us / 1e3 / peak, 4)}
                del qw, qz, sc
        # ... and the bf16-native form (scales and output bf16, one rounding) for three of them
        dqb = {}
        outb = outd.view(torch.bfloat16)
        for b in (3, 4, 8):
            g = 128
            qw = torch.randint(-2**31, 2**31 - 1, (Rd, (Kd * b + 31) // 32, Nd), dtype=torch.int32, device=dev, generator=gen)
            qz = torch.randint(-2**31, 2**31 - 1, (Rd, Kd // g, (Nd * b + 31) // 32), dtype=torch.int32, device=dev, generator=gen)
            sc = (torch.rand((Rd, Kd // g, Nd), device=dev, generator=gen) * 0.018 + 0.002).to(torch.bfloat16)

            def fn(i, b=b, g=g, qw=qw, qz=qz, sc=sc):
                j = i % Rd
                rc = lib.xbit_dequant_bf16(qw[j].data_ptr(), sc[j].data_ptr(), qz[j].data_ptr(), outb[j].data_ptr(), Kd, Nd, b, g, 1,
                                           torch.cuda.current_stream().cuda_stream)
                if rc != 0:
                    raise RuntimeError(capi.last_error())
            us = graph_us(fn, 2 * Rd)
            nb = synth.dq_bytes(Kd, Nd, b, g)
            dqb[f"b{b}_g{g}"] = {"us": round(us, 2), "GBps": round(nb / us / 1e3, 1), "frac_of_measured_peak": round(nb / us / 1e3 / peak, 4)}
            del qw, qz, sc
        line["dq_bf16_native"] = {"shape": "4096x11008", "by_bits_groupsize": dqb}
        line["dq"] = {"shape": "4096x11008", "by_bits_groupsize": dq,
                      "note": "algorithmic bytes = packed weights + scales + zeros + fp16 output (80 % of it is the output write)"}
        del outd

        # (5) the reference's own GPU kernels (unmodified, built for compute_100 into oracle/_ref/refgpu) on the same tensors:
        #     a reported baseline like cpu_baseline.  They launch on the legacy default stream and cannot be captured,
        #     so they are timed as back-to-back eager calls between CUDA events, at::zeros of the op included.
        import glob
        import importlib.util
        so = sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "refgpu", "XbitOps*.so")))
        if so:
            try:
                spec = importlib.util.spec_from_file_location("XbitOps", so[0])
                refmod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(refmod)
                refgpu = {}
                for ss in sets:
                    if ss.N_total % 64:
                        continue
                    n = min(ss.R, 64)
                    for j in range(3):
                        refmod.gemv(ss.a, ss.qw[j], ss.sc[j], ss.qz[j], GROUP, BITS, ss.K, 0)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for j in range(n):
                        refmod.gemv(ss.a, ss.qw[j], ss.sc[j], ss.qz[j], GROUP, BITS, ss.K, 0)
                    e1.record()
                    torch.cuda.synchronize()
                    us = e0.elapsed_time(e1) * 1e3 / n
                    refgpu[f"{ss.K}x{ss.N_total}"] = {"us_per_call": round(us, 2), "GBps": round(ss.bytes_call / us / 1e3, 1)}
                line["reference_gpu"] = {"per_shape": refgpu, "how": "unmodified reference extension (gemv_w4a16_pt.cu compiled for compute_100), "
                                         "eager back-to-back calls over the same rotating sets, CUDA events, its at::zeros included"}
            except Exception as ex:  # noqa: BLE001
                line["reference_gpu"] = {"unavailable": str(ex)[:200]}
        else:
            line["reference_gpu"] = {"unavailable": "oracle/_ref/refgpu not built (oracle/build_ref_gpu.sh)"}

    # ---- e2e: the public host-buffer entry point, H2D activations + D2H result inside the timed region
    if world == 1:
        hs = [(torch.randn((1, ss.K)).to(torch.float16).pin_memory(), torch.empty((1, ss.N_total), dtype=torch.float16).pin_memory(),
               torch.empty((1, ss.K), dtype=torch.float16, device=dev), torch.empty((1, ss.N_total), dtype=torch.float16, device=dev))
              for ss in sets]
        def e2e_calls():
            st = torch.cuda.current_stream().cuda_stream
            for ss, (ha, ho, da, do) in zip(sets, hs):
                for j in range(ss.R):
                    rc = lib.xbit_gemv_f16_host(ha.data_ptr(), ho.data_ptr(), da.data_ptr(), do.data_ptr(), ss.qw[j].data_ptr(),
                                                ss.sc[j].data_ptr(), ss.qz[j].data_ptr(), 1, ss.K, ss.N, BITS, GROUP, 0, ws_ptr, ws_len, st)
                    if rc != 0:
                        raise RuntimeError(capi.last_error())

        def timed_host(fn, steps):
            for _ in range(3):
                fn()
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
                torch.cuda.synchronize()      # the step's results are in host memory
            return (time.perf_counter() - t0) / steps

        e2e_steps = max(3, min(args.steps, 20))
        dt_eager = timed_host(e2e_calls, e2e_steps)
        # the same calls recorded once into a CUDA graph (the entry point is capturable: H2D copy node ->
        # kernel -> D2H copy node per call) and replayed per step
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            e2e_calls()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ge = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ge):
            e2e_calls()
        dt = timed_host(ge.replay, e2e_steps)
        line["e2e"] = {"value": round(step_bytes / dt / 1e9, 2), "unit": UNIT,
                       "h2d_bytes_per_step": int(sum(ss.K * 2 * ss.R for ss in sets)),
                       "d2h_bytes_per_step": int(sum(ss.N_total * 2 * ss.R for ss in sets)),
                       "us_per_call": round(dt * 1e6 / calls_per_step, 3),
                       "eager": {"value": round(step_bytes / dt_eager / 1e9, 2), "us_per_call": round(dt_eager * 1e6 / calls_per_step, 3)},
                       "how": "xbit_gemv_f16_host per call (pinned H2D activations -> gemv -> D2H result into pinned host memory), "
                              "the step's calls captured once into a CUDA graph and replayed, host wall clock incl. one sync per step; "
                              "weights resident; 'eager' = the same calls issued one by one from Python"}
    else:
        # N > 1: host buffers in, host buffers out, through the sharded public path with a plain [M, N] result: pinned
        # activations -> device (copy node), N-split GEMV with the fused peer-store exchange (+ barrier) or NCCL all-gather,
        # gathered result -> pinned host memory (copy node); every rank does all of it, host wall clock, max over ranks
        e2e_mode = "peers" if sets[0].symm is not None else "nccl"
        hs = [(torch.randn((1, ss.K)).to(torch.float16).pin_memory(), torch.empty((1, ss.N_total), dtype=torch.float16).pin_memory()) for ss in sets]
        saved_a = [ss.a for ss in sets]
        stage = [torch.empty_like(ss.a) for ss in sets]

        def e2e_calls():
            for ss, (ha, ho), da in zip(sets, hs, stage):
                ss.a = da
                for j in range(ss.R):
                    da.copy_(ha, non_blocking=True)
                    launch(ss, j, e2e_mode)
                    res = ss.symm[0][j] if e2e_mode == "peers" else ss.out[j]
                    ho.copy_(res, non_blocking=True)

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            e2e_calls()
        torch.cuda.current_stream().wait_stream(side)
        sync_all()
        ge = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ge):
            e2e_calls()
        e2e_steps = max(3, min(args.steps, 20))
        for _ in range(3):
            ge.replay()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ge.replay()
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        for ss, sa in zip(sets, saved_a):
            ss.a = sa
        line["e2e"] = {"value": round(step_bytes / dt / 1e9, 2), "unit": UNIT,
                       "h2d_bytes_per_step": int(sum(ss.K * 2 * ss.R for ss in sets)) * world,
                       "d2h_bytes_per_step": int(sum(ss.N_total * 2 * ss.R for ss in sets)) * world,
                       "us_per_call": round(dt * 1e6 / calls_per_step, 3),
                       "how": f"per call and rank: pinned activations -> device copy, N-split GEMV with the {e2e_mode} exchange (plain [M, N] output), "
                              "gathered result -> pinned host copy; the step's calls captured into one CUDA graph per rank, host wall clock incl. "
                              "one sync per step, max over ranks"}
        del ge

    # ---- cpu baseline beside it (rank 0, N=1 only): bounded sample, single thread, stated
    if world == 1 and rank == 0 and not args.no_cpu:
        secs, nbytes, kind = cpu_reference_run([shapes[0]], 1)
        line["cpu_baseline"] = {"value": round(nbytes / secs / 1e9, 5), "unit": UNIT, "cores": 1, "kind": kind,
                                "sample": f"1 call of {shapes[0][0]}x{shapes[0][1]} (dequant via cpp_simulate.cc + fp64 dot), {secs:.2f} s"}
    elif rank == 0 and not args.no_cpu:
        secs, nbytes, kind = cpu_reference_run([shapes[0]], 1)
        line["cpu_baseline"] = {"value": round(nbytes / secs / 1e9, 5), "unit": UNIT, "cores": 1, "kind": kind,
                                "sample": f"1 call of {shapes[0][0]}x{shapes[0][1]} (dequant via cpp_simulate.cc + fp64 dot), {secs:.2f} s, rank 0"}
    elif rank == 0:
        line["cpu_baseline"] = None

    if world > 1 and not args.no_single:
        # the same (unsharded) workload on ONE GPU, for the strong-scaling context: rank 0 alone, others wait
        del graph
        sets_single = None
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            for ss in sets:
                ss.qw = ss.sc = ss.qz = ss.out = None
            torch.cuda.empty_cache()
            sets_single = [ShapeSet(torch, dev, K, N, 1, 0, gen) for (K, N) in shapes]
            saved = (world, combine)

            def launch1(ss, j):
                rc = lib.xbit_gemv_f16_peers_ex(ss.a.data_ptr(), ss.qw[j].data_ptr(), ss.sc[j].data_ptr(), ss.qz[j].data_ptr(),
                                                (ctypes.c_void_p * 1)(ss.out[j].data_ptr()), 1, 1, ss.K, ss.N, BITS, GROUP, 0,
                                                ss.N_total, 0, ws_ptr, ws_len, family | flags, torch.cuda.current_stream().cuda_stream)
                if rc != 0:
                    raise RuntimeError(capi.last_error())

            side = torch.cuda.Stream()
            with torch.cuda.stream(side):
                for ss in sets_single:
                    launch1(ss, 0)
            torch.cuda.synchronize()
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                for ss in sets_single:
                    for j in range(ss.R):
                        launch1(ss, j)
            for _ in range(3):
                g1.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(3, args.steps // 4)
            e0.record()
            for _ in range(reps):
                g1.replay()
            e1.record()
            torch.cuda.synchronize()
            ms1 = e0.elapsed_time(e1) / reps
            b1 = sum(ss.bytes_call * ss.R for ss in sets_single)
            v1 = b1 / (ms1 * 1e-3) / 1e9
            line["single_gpu_same_workload"] = {"value": round(v1, 2), "unit": UNIT,
                                                "ms_per_step": round(ms1, 5), "calls_per_step": sum(ss.R for ss in sets_single)}
            line["efficiency_same_workload"] = round(value / (world * v1), 4)
            del g1
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)      # NCCL communicators captured in CUDA graphs can stall interpreter teardown
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, *WORKLOADS])
    ap.add_argument("--family", default="auto", choices=["auto", "simt", "mma"])
    ap.add_argument("--no-pdl", action="store_true", help="do not assert static weights (no prefetch before griddepcontrol.wait)")
    ap.add_argument("--combine", default="ll", choices=["ll", "peers", "signal", "nccl", "none"],
                    help="N>1: how output slices are combined (fused NVLink peer stores + symmetric-memory barrier | flag-in-data "
                         "8-byte stores consumed by the next call of a dependent chain | peer stores + in-kernel completion flags "
                         "awaited by the next call | NCCL all-gather | kernel only)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-streamk", action="store_true", help="no workspace: cluster split-K kernel instead of the persistent stream-K schedule")
    ap.add_argument("--no-single", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline, roofline, e2e and cpu_baseline only (skip the op-surface, fused, skinny, dq and reference-GPU legs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
