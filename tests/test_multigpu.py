"""N-split GEMV over 2 GPUs (NCCL all-gather baseline and the fused NVLink peer-store epilogue)
against the unsharded single-GPU result.  Needs >= 2 CUDA devices: skipped on a 1-GPU box."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

pytestmark = pytest.mark.gpu

from xbitops_b200 import synth  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _chain(rank, world, dev, X):
    """Two dependent N-split calls through the C ABI: call 2 reads call 1's gathered output straight from
    the symmetric buffer and awaits the gather inside its own kernel (XBIT_GEMV_FLAG_WAIT_PEERS)."""
    import ctypes
    import torch.distributed._symmetric_memory as symm_mem
    from xbitops_b200 import capi
    from xbitops_b200.sharded import shard_columns
    lib = capi.load()
    K = N = 4096
    qw, s, qz, a = synth.make_inputs(K, N, 4, 128, M=1, seed=77)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
    tq, ts, tz = d(qw), d((s.view(np.int16))).view(torch.float16) * 0.02, d(qz)   # small scales: keep y2 in fp16 range
    ta = d(a.view(np.int16)).view(torch.float16)
    y1_ref = X.gemv(ta, tq, ts, tz, 128, 4, K, 1)
    y2_ref = X.gemv(y1_ref, tq, ts, tz, 128, 4, K, 1)
    q, sc, z = shard_columns(tq, ts, tz, 4, world, rank)
    name = dist.group.WORLD.group_name
    buf = symm_mem.empty((2, 1, N), dtype=torch.float16, device=dev)
    hdl = symm_mem.rendezvous(buf, name)
    flags = symm_mem.empty((64,), dtype=torch.int32, device=dev)
    flags.zero_()
    fhdl = symm_mem.rendezvous(flags, name)
    state = torch.zeros(4, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    hdl.barrier()
    st = torch.cuda.current_stream().cuda_stream
    fptrs = [int(p) for p in fhdl.buffer_ptrs]
    farr = (ctypes.c_void_p * world)(*fptrs)
    n_local = N // world
    for call, (src, k, extra) in enumerate(((ta, 0, 0), (buf[0], 1, capi.GEMV_FLAG_WAIT_PEERS))):
        outs = (ctypes.c_void_p * world)(*[int(p) + k * N * 2 for p in hdl.buffer_ptrs])
        capi.check(lib.xbit_gemv_f16_peers_signal(src.data_ptr(), q.data_ptr(), sc.data_ptr(), z.data_ptr(), outs, farr,
                                                  state.data_ptr(), world, rank, 1, K, n_local, 4, 128, 1, N, rank * n_local,
                                                  capi.GEMV_AUTO | extra, st))
    capi.check(lib.xbit_peers_wait(fptrs[rank], world, rank, state.data_ptr() + 12, st))
    torch.cuda.synchronize()
    why = []
    if int(state[3].item()) != 0:
        why.append("peer wait timed out")
    if int(flags[:world].min().item()) != 2:
        why.append(f"flags {flags[:world].tolist()}")
    # (the K split chosen by the planner depends on N_local, so the fp32 summation order may differ from the unsharded call)
    e1 = float((buf[0, 0].double() - y1_ref[0].double()).abs().max()) / float(y1_ref.double().abs().max())
    e2 = float((buf[1, 0].double() - y2_ref[0].double()).abs().max()) / float(y2_ref.double().abs().max())
    if not (e1 < 2e-3 and e2 < 4e-3):
        why.append(f"chain errors {e1:.3e} {e2:.3e}")
    hdl.barrier()
    if why:
        print(f"[rank {rank}] chain test: " + "; ".join(why), flush=True)
    return True if not why else "; ".join(why)


def _chain_ll(rank, world, dev, X):
    """Three dependent N-split calls in the flag-in-data form: every call stores {results, call number} slots
    into every rank's LL buffer, the next call's activation staging spins on the slots it needs; two LL
    buffers alternate; the last result is unpacked to plain fp16."""
    import ctypes
    import torch.distributed._symmetric_memory as symm_mem
    from xbitops_b200 import capi
    from xbitops_b200.sharded import shard_columns
    lib = capi.load()
    K = N = 4096
    qw, s, qz, a = synth.make_inputs(K, N, 4, 128, M=2, seed=78)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
    tq, ts, tz = d(qw), d((s.view(np.int16))).view(torch.float16) * 0.02, d(qz)
    ta = d(a.view(np.int16)).view(torch.float16)
    M = ta.shape[0]
    ref = ta
    for _ in range(3):
        ref = X.gemv(ref, tq, ts, tz, 128, 4, K, 1)
    q, sc, z = shard_columns(tq, ts, tz, 4, world, rank)
    name = dist.group.WORLD.group_name
    NBUF, REPS = 3, 4                                                   # (chain_len - 1) % NBUF != 0: see include/xbitops_b200.h
    ll = symm_mem.empty((NBUF, M, N), dtype=torch.int32, device=dev)    # 8-byte slot per pair of results
    ll.zero_()
    hdl = symm_mem.rendezvous(ll, name)
    state = torch.zeros(4, dtype=torch.int32, device=dev)
    out = torch.empty((M, N), dtype=torch.float16, device=dev)
    torch.cuda.synchronize()
    hdl.barrier()
    st = torch.cuda.current_stream().cuda_stream
    n_local = N // world
    bufsz = M * N * 4
    for _ in range(REPS):                   # the same chain back to back, as a replayed CUDA graph would issue it
        src, flag = ta.data_ptr(), 0
        for call in range(3):
            k = call % NBUF
            outs = (ctypes.c_void_p * world)(*[int(p) + k * bufsz for p in hdl.buffer_ptrs])
            capi.check(lib.xbit_gemv_f16_peers_ll(src, q.data_ptr(), sc.data_ptr(), z.data_ptr(), outs, state.data_ptr(), call, world, rank,
                                                  M, K, n_local, 4, 128, 1, N, rank * n_local, capi.GEMV_AUTO | flag, st))
            src, flag = int(hdl.buffer_ptrs[rank]) + k * bufsz, capi.GEMV_FLAG_A_IS_LL
        capi.check(lib.xbit_ll_unpack_f16(src, out.data_ptr(), M * N, state.data_ptr(), 3, state.data_ptr() + 12, st))
    torch.cuda.synchronize()
    why = []
    if int(state[3].item()) != 0:
        why.append("unpack timed out")
    if int(state[2].item()) != 3 * REPS:
        why.append(f"chain base {int(state[2].item())}")
    e = float((out.double() - ref.double()).abs().max()) / float(ref.double().abs().max())
    if not e < 6e-3:
        why.append(f"chain error {e:.3e}")
    hdl.barrier()
    # the ready-made wrapper on the same weights: 4096 -> 4096 -> 4096, twice
    from xbitops_b200.sharded import ShardedQChain
    chain = ShardedQChain([(q, sc, z, K, N)] * 3, 128, 4, 1, max_rows=M)
    for _ in range(8):                      # back-to-back chains: the chain base must be seen fresh by every call
        yc = chain(ta)
    torch.cuda.synchronize()
    chain.check_timeout()
    one = ShardedQChain([(q, sc, z, K, N)], 128, 4, 1, max_rows=M)      # a single-layer chain alternates its two buffers
    y1 = X.gemv(ta, tq, ts, tz, 128, 4, K, 1)
    for _ in range(4):
        yo = one(ta)
    torch.cuda.synchronize()
    one.check_timeout()
    eo = float((yo.double() - y1.double()).abs().max()) / float(y1.double().abs().max())
    if not eo < 6e-3:
        why.append(f"single-layer chain error {eo:.3e}")
    ec = float((yc.double() - ref.double()).abs().max()) / float(ref.double().abs().max())
    if not ec < 6e-3 or chain._bufs[2].tolist()[2:] != [24, 0]:
        why.append(f"ShardedQChain error {ec:.3e} state {chain._bufs[2].tolist()}")
    if why:
        print(f"[rank {rank}] chain test: " + "; ".join(why), flush=True)
    return True if not why else "; ".join(why)


def _worker(rank, world, port, combine, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import xbitops_b200 as X
        from xbitops_b200.sharded import ShardedQLinear, shard_columns
        if combine == "ll":
            ret[rank] = _chain_ll(rank, world, dev, X)
            return
        ok = True
        for (K, N, M) in ((8192, 8192, 1), (4096, 1024, 3)):
            qw, s, qz, a = synth.make_inputs(K, N, 4, 128, M=M, seed=K + M)
            d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
            tq, ts, tz = d(qw), d(s.view(np.int16)).view(torch.float16), d(qz)
            ta = d(a.view(np.int16)).view(torch.float16)
            full = X.gemv(ta, tq, ts, tz, 128, 4, K, 1)
            q, sc, z = shard_columns(tq, ts, tz, 4, world, rank)
            lin = ShardedQLinear(q, sc, z, 128, 4, K, N, 1, combine=combine)
            for _ in range(3):                  # repeated calls: call counters / double buffering of the signal form
                y = lin(ta)
            torch.cuda.synchronize()
            truth = ta.double() @ X.dequant(tq, ts, tz, 128, 4, K, 1).double()
            err = (y.double() - truth).abs().max() / truth.abs().max()
            ok = ok and tuple(y.shape) == (M, N) and float(err) < 1e-2
            # every rank holds the same gathered result
            ref = y.clone()
            dist.broadcast(ref, 0)
            ok = ok and bool(torch.equal(ref, y))
            # and it agrees with the unsharded call to fp16 rounding of the same fp32 sums
            ok = ok and float((y.double() - full.double()).abs().max() / truth.abs().max()) < 2e-3
        if combine == "signal" and ok:
            ok = _chain(rank, world, dev, X)
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", (2, 4, 8))
@pytest.mark.parametrize("combine", ("nccl", "peers", "signal", "ll"))
def test_sharded_gemv_multi_gpu(combine, world):
    """N-split over 2 / 4 / 8 GPUs of one box in every exchange form: against a @ dequant, identical on every rank, and
    against the unsharded call (skipped where the box has fewer GPUs; bench.py --gpus N repeats the check before timing)."""
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test without a CUDA device")
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), combine, ret), nprocs=world, join=True)
        assert all(ret.get(r) is True for r in range(world)), dict(ret)
