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


def _worker(rank, world, port, combine, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import xbitops_b200 as X
        from xbitops_b200.sharded import ShardedQLinear, shard_columns
        ok = True
        for (K, N, M) in ((8192, 8192, 1), (4096, 1024, 3)):
            qw, s, qz, a = synth.make_inputs(K, N, 4, 128, M=M, seed=K + M)
            d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
            tq, ts, tz = d(qw), d(s.view(np.int16)).view(torch.float16), d(qz)
            ta = d(a.view(np.int16)).view(torch.float16)
            full = X.gemv(ta, tq, ts, tz, 128, 4, K, 1)
            q, sc, z = shard_columns(tq, ts, tz, 4, world, rank)
            lin = ShardedQLinear(q, sc, z, 128, 4, K, N, 1, combine=combine)
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
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("combine", ("nccl", "peers"))
def test_sharded_gemv_two_gpus(combine):
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test without a CUDA device")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), combine, ret), nprocs=2, join=True)
        assert ret.get(0) is True and ret.get(1) is True
