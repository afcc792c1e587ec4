"""Host-side logic of the N-split path on CPU: column sharding against the oracle, and the
gather plumbing of ShardedQLinear with world_size 2 over gloo (the per-rank compute is injected:
the oracle plays the device here, because this test checks the sharding/gather logic, not a kernel)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from xbitops_b200 import synth
from xbitops_b200.sharded import ShardedQLinear, shard_columns


@pytest.mark.parametrize("bits", (2, 3, 4, 8))
def test_shard_columns_matches_unsharded_dequant(bits, c_oracle):
    K, N, g, world = 256, 128, 64, 4
    qw, s, qz, _ = synth.make_inputs(K, N, bits, g, seed=bits)
    full = c_oracle.dequant(qw, s, qz, g, bits, K, 1)
    for rank in range(world):
        q, sc, z = shard_columns(qw, s, qz, bits, world, rank)
        part = c_oracle.dequant(q, sc, z, g, bits, K, 1)
        n = N // world
        assert (part.view(np.uint16) == full[:, rank * n:(rank + 1) * n].view(np.uint16)).all()
    with pytest.raises(ValueError):
        shard_columns(qw, s, qz, bits, 3, 0)
    if bits == 3:
        with pytest.raises(ValueError):     # 128/16 = 8 columns x 3 bits is not a whole qzeros word
            shard_columns(qw, s, qz, bits, 16, 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, M, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        K, N, g, bits = 256, 64, 128, 4
        qw, s, qz, a = synth.make_inputs(K, N, bits, g, M=M, seed=3)
        co = O.COracle()
        q, sc, z = shard_columns(qw, s, qz, bits, world, rank)

        def oracle_compute(x, out_full, col0):          # stands in for the CUDA kernel on CPU
            w = co.dequant(q, sc, z, g, bits, K, 0)
            _, y16 = co.gemv_from_dq(x.numpy().view(np.float16), w)
            out_full[:, col0:col0 + y16.shape[1]] = torch.from_numpy(y16.view(np.int16)).view(torch.float16)

        lin = ShardedQLinear(torch.from_numpy(q), torch.from_numpy(sc.view(np.int16)).view(torch.float16),
                             torch.from_numpy(z), g, bits, K, N, 0, local_gemv=oracle_compute)
        x = torch.from_numpy(a.view(np.int16)).view(torch.float16)
        y = lin(x)
        _, want = co.gemv(a, qw, s, qz, g, bits, K, 0)
        ok = bool((y.numpy().view(np.uint16) == want.view(np.uint16)).all())
        ret[rank] = ok and tuple(y.shape) == (M, N)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("M", (1, 3))
def test_sharded_qlinear_gloo_world2(M):
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        port = _free_port()
        mp.spawn(_worker, args=(world, port, M, ret), nprocs=world, join=True)
        assert ret.get(0) is True and ret.get(1) is True


def test_sharded_chain_argument_checks():
    """ShardedQChain validates the chain on the host, before anything touches the GPU: consecutive layers must
    fit, shards must have the right width, only the W4 fast path and M <= 16 are accepted; the LL buffers rotate
    over m slots with (len - 1) % m != 0."""
    from xbitops_b200.sharded import ShardedQChain

    def layer(K, N, bits=4, g=128):
        qw, s, qz, _ = synth.make_inputs(K, N, bits, g, M=1, seed=K + N)
        return (torch.from_numpy(qw), torch.from_numpy(s.view(np.int16)).view(torch.float16), torch.from_numpy(qz), K, N)

    a, b, c = layer(256, 512), layer(512, 256), layer(256, 256)
    assert ShardedQChain([a, b, c], 128).nbuf == 3            # 3 layers: (3 - 1) % 2 == 0 -> three buffers
    assert ShardedQChain([a, b], 128).nbuf == 2
    assert ShardedQChain([a, b, c, c, c, c, c], 128).nbuf == 4          # (7 - 1) % 2 == (7 - 1) % 3 == 0
    with pytest.raises(ValueError):
        ShardedQChain([], 128)
    with pytest.raises(ValueError):
        ShardedQChain([a, c], 128)                              # 512 outputs do not feed 256 inputs
    with pytest.raises(ValueError):
        ShardedQChain([(a[0][:, :64], a[1], a[2], 256, 512)], 128)   # shard width does not match out_features / world
    with pytest.raises(ValueError):
        ShardedQChain([layer(256, 512, bits=3)], 128, bits=3)  # fast path only
    with pytest.raises(ValueError):
        ShardedQChain([layer(192, 512, g=64)], 64)              # K % 128 != 0
    with pytest.raises(ValueError):
        ShardedQChain([a], 128, max_rows=17)
