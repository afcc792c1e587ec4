"""N-split multi-GPU wrapper for the A16Wx GEMV; no reference counterpart -- the
reference is single-GPU (no nccl / torch.distributed anywhere in /root/reference).

Output columns are independent, so rank p of P owns columns [p*N/P, (p+1)*N/P): qweight[:, slice],
scales[:, slice] and qzeros[:, slice*bits/32] (the slice must fall on qzeros word boundaries, i.e.
(N/P * bits) % 32 == 0).  Activations are replicated.  The only exchange step is the gather of the
fp16 output slices (2-56 KiB in total at batch 1), done either
  * "nccl":  in-place all_gather_into_tensor on the compute stream (baseline), or
  * "peers": fused into the GEMV epilogue -- every rank's kernel stores its slice straight into every
             rank's output buffer through NVLink peer mappings (torch symmetric memory), followed by
             one symmetric-memory barrier.
One process per GPU; torch.distributed supplies the plumbing only.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import capi


def shard_columns(qweight, scales, qzeros, bits: int, world: int, rank: int):
    """Column shard `rank` of `world` of (qweight [R, N], scales [G, N], qzeros [G, N*bits/32]).
    Works on numpy arrays and torch tensors; returns contiguous copies."""
    n_total = qweight.shape[1]
    if n_total % world:
        raise ValueError(f"out_features {n_total} not divisible by world size {world}")
    n = n_total // world
    if (n * bits) % 32:
        raise ValueError(f"shard width {n} x {bits} bits does not fall on qzeros word boundaries")
    zw = n * bits // 32
    cols = slice(rank * n, (rank + 1) * n)
    zcols = slice(rank * zw, (rank + 1) * zw)
    parts = (qweight[:, cols], scales[:, cols], qzeros[:, zcols])
    if isinstance(qweight, torch.Tensor):
        return tuple(p.contiguous() for p in parts)
    import numpy as np
    return tuple(np.ascontiguousarray(p) for p in parts)


def _device_gemv_into(x, qweight, scales, qzeros, groupsize, bits, in_features, add_zero_bias, out_full, col_offset,
                      peer_ptrs=None, family=capi.GEMV_AUTO):
    """Local shard GEMV through the C ABI, written into out_full[:, col_offset:col_offset+n]
    (and, with peer_ptrs, into every rank's buffer)."""
    from .ops import gemv_workspace
    lib = capi.load()
    m, n = x.shape[0], qweight.shape[1]
    ptrs = peer_ptrs if peer_ptrs is not None else [out_full.data_ptr()]
    arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
    ws = gemv_workspace(x.device)
    capi.check(lib.xbit_gemv_f16_peers_ex(x.data_ptr(), qweight.data_ptr(), scales.data_ptr(), qzeros.data_ptr(), arr,
                                          len(ptrs), m, in_features, n, bits, groupsize, int(add_zero_bias),
                                          out_full.shape[1], col_offset, ws.data_ptr(), ws.numel(), int(family),
                                          torch.cuda.current_stream().cuda_stream))


def _device_gemv_signal(x, qweight, scales, qzeros, groupsize, bits, in_features, add_zero_bias, out_view, col_offset,
                        peer_ptrs, flag_ptrs, state, rank, family=capi.GEMV_AUTO) -> bool:
    """Shard GEMV whose epilogue stores the slice into every rank's buffer AND raises the per-rank
    completion flag (xbit_gemv_f16_peers_signal), followed by the one-warp wait.  False = this shape is not
    covered by the fused signal (the caller falls back to the barrier form)."""
    lib = capi.load()
    m, n = x.shape[0], qweight.shape[1]
    if m > 16 or bits != 4 or groupsize not in (32, 64, 128) or in_features % 128 or n % 32:
        return False
    world = len(peer_ptrs)
    outs = (ctypes.c_void_p * world)(*peer_ptrs)
    flags = (ctypes.c_void_p * world)(*flag_ptrs)
    st = torch.cuda.current_stream().cuda_stream
    capi.check(lib.xbit_gemv_f16_peers_signal(x.data_ptr(), qweight.data_ptr(), scales.data_ptr(), qzeros.data_ptr(), outs,
                                              flags, state.data_ptr(), world, rank, m, in_features, n, bits, groupsize,
                                              int(add_zero_bias), out_view.shape[1], col_offset, int(family), st))
    capi.check(lib.xbit_peers_wait(flag_ptrs[rank], world, rank, state.data_ptr() + 12, st))
    return True


class ShardedQLinear:
    """y = x @ DQ(W) with W split by output columns over the ranks of `group`.

    local_gemv(x, out_full, col_offset) may be injected (the CPU/gloo tests do) -- the default is the
    CUDA path through the C ABI."""

    def __init__(self, qweight_shard, scales_shard, qzeros_shard, groupsize: int, bits: int, in_features: int,
                 out_features: int, add_zero_bias: int = 0, group=None, combine: str = "nccl", local_gemv=None):
        self.qweight, self.scales, self.qzeros = qweight_shard, scales_shard, qzeros_shard
        self.groupsize, self.bits, self.in_features = groupsize, bits, in_features
        self.out_features, self.add_zero_bias = out_features, add_zero_bias
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if out_features % self.world:
            raise ValueError("out_features must be divisible by the world size")
        self.n_local = out_features // self.world
        if qweight_shard.shape[1] != self.n_local:
            raise ValueError(f"qweight shard has {qweight_shard.shape[1]} columns, expected {self.n_local}")
        if combine not in ("nccl", "peers", "signal", "none"):
            raise ValueError(combine)
        self.combine = combine
        self._local = local_gemv
        self._symm = None      # (buffer, handle, peer_ptrs) for combine == "peers"

    # -- local compute -------------------------------------------------------------------------
    def _local_gemv(self, x, out_full, peer_ptrs=None):
        col0 = self.rank * self.n_local
        if self._local is not None:
            self._local(x, out_full, col0)
        else:
            _device_gemv_into(x, self.qweight, self.scales, self.qzeros, self.groupsize, self.bits, self.in_features,
                              self.add_zero_bias, out_full, col0, peer_ptrs)

    # -- symmetric memory for the fused epilogue -------------------------------------------------
    def _symm_buffer(self, m: int, device):
        """Two [m, out_features] result buffers (used alternately: a rank publishes call e+1 only after it has
        consumed result e on the same stream, so whoever has seen all flags of e+1 may overwrite buffer e%2)
        and a flag array, all in symmetric memory; the call counters live in ordinary device memory."""
        if self._symm is not None and self._symm[0].shape[1] >= m:
            return self._symm
        import torch.distributed._symmetric_memory as symm_mem
        name = self.group.group_name if self.group is not None else dist.group.WORLD.group_name
        buf = symm_mem.empty((2, m, self.out_features), dtype=torch.float16, device=device)
        hdl = symm_mem.rendezvous(buf, name)
        flags = symm_mem.empty((64,), dtype=torch.int32, device=device)
        flags.zero_()
        fhdl = symm_mem.rendezvous(flags, name)
        state = torch.zeros(4, dtype=torch.int32, device=device)   # tiles done, send epoch, wait epoch, timeout
        torch.cuda.synchronize(device)
        hdl.barrier()                                              # every rank's flags are zero before anyone signals
        self._symm = (buf, hdl, [int(p) for p in hdl.buffer_ptrs], flags, [int(p) for p in fhdl.buffer_ptrs], state)
        self._calls = 0
        return self._symm

    # -- forward -----------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """x [M, K] fp16 (replicated on every rank) -> [M, out_features] fp16 on every rank."""
        m = x.shape[0]
        if self.combine in ("peers", "signal") and self.world > 1:
            buf, hdl, ptrs, flags, fptrs, state = self._symm_buffer(m, x.device)
            k = self._calls & 1
            self._calls += 1
            view = buf[k, :m]
            off = k * buf.shape[1] * buf.shape[2] * 2
            kptrs = [p + off for p in ptrs]
            if self.combine == "signal" and self._local is None and _device_gemv_signal(x, self.qweight, self.scales, self.qzeros, self.groupsize, self.bits,
                                                           self.in_features, self.add_zero_bias, view, self.rank * self.n_local,
                                                           kptrs, fptrs, state, self.rank):
                return view                    # synchronisation fused into the kernel + a one-warp wait
            hdl.barrier()                      # everyone has consumed the previous result
            self._local_gemv(x, view, kptrs)
            hdl.barrier()                      # every slice has landed everywhere
            return view
        if out is None:
            out = torch.empty((m, self.out_features), dtype=torch.float16, device=x.device)
        self._local_gemv(x, out)
        if self.world == 1 or self.combine == "none":
            return out
        col0 = self.rank * self.n_local
        if m == 1:
            # in place: send = recv + rank*count
            dist.all_gather_into_tensor(out.view(-1), out[:, col0:col0 + self.n_local].reshape(-1), group=self.group)
            return out
        # gathered layout is [P][M][N/P]: gather into scratch, then one strided copy back
        scratch = torch.empty((self.world, m, self.n_local), dtype=torch.float16, device=x.device)
        dist.all_gather_into_tensor(scratch.view(-1), out[:, col0:col0 + self.n_local].contiguous().view(-1), group=self.group)
        out.view(m, self.world, self.n_local).copy_(scratch.permute(1, 0, 2))
        return out

    __call__ = forward


class ShardedQChain:
    """A chain of dependent N-split GEMVs (a decode step: y = L_n(...L_2(L_1(x)))) in the flag-in-data form
    (xbit_gemv_f16_peers_ll): every layer's epilogue stores {two results, call number} slots into every rank's LL
    buffer over NVLink, the next layer's kernel spins on the slots it needs while staging its activations -- no
    barrier, no fence, no wait launch between the layers -- and xbit_ll_unpack_f16 hands back plain fp16.

    layers: sequence of (qweight_shard, scales_shard, qzeros_shard, in_features, out_features) with
    out_features[i] == in_features[i + 1]; W4 fast path only (bits 4, groupsize 32/64/128, in_features % 128 == 0,
    shard width % 32 == 0), M <= 16."""

    def __init__(self, layers, groupsize: int, bits: int = 4, add_zero_bias: int = 0, group=None, max_rows: int = 1):
        self.layers = list(layers)
        if not self.layers:
            raise ValueError("empty chain")
        self.groupsize, self.bits, self.add_zero_bias, self.group = groupsize, bits, add_zero_bias, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        for i, (q, _, _, k, n) in enumerate(self.layers):
            if n % self.world or q.shape[1] != n // self.world:
                raise ValueError(f"layer {i}: qweight shard has {q.shape[1]} columns, expected {n // self.world}")
            if i + 1 < len(self.layers) and self.layers[i + 1][3] != n:
                raise ValueError(f"layer {i}: out_features {n} != in_features {self.layers[i + 1][3]} of the next layer")
            if bits != 4 or groupsize not in (32, 64, 128) or k % 128 or (n // self.world) % 32:
                raise ValueError("the flag-in-data chain needs the W4 fast path (bits 4, groupsize 32/64/128, K%128=0, shard%32=0)")
        if max_rows > 16:
            raise ValueError("the flag-in-data chain takes at most 16 activation rows")
        self.max_rows = max_rows
        # a replayed / repeated chain starts again on buffer 0 while a slower rank may still be unpacking the last
        # layer's buffer: rotate over m buffers with (len - 1) % m != 0 so that the two never coincide
        # (a single-layer chain has no such m: it alternates two buffers from one forward() to the next instead, which
        # covers eager calls; a CAPTURED single-layer chain must not be replayed back to back without a rank barrier)
        self.nbuf = next((m for m in range(2, 9) if (len(self.layers) - 1) % m != 0), 2)
        self._first_buf = 0
        self._bufs = None

    def _buffers(self, device):
        if self._bufs is not None:
            return self._bufs
        max_n = max(l[4] for l in self.layers)
        shape = (self.nbuf, self.max_rows, max_n)            # one int32 per result: 8 bytes per pair
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm_mem
            ll = symm_mem.empty(shape, dtype=torch.int32, device=device)
            ll.zero_()
            hdl = symm_mem.rendezvous(ll, self.group.group_name if self.group is not None else dist.group.WORLD.group_name)
            torch.cuda.synchronize(device)
            hdl.barrier()                                    # every rank's slots are zero before anyone stores
            ptrs = [int(p) for p in hdl.buffer_ptrs]
        else:
            ll = torch.zeros(shape, dtype=torch.int32, device=device)
            ptrs = [ll.data_ptr()]
        state = torch.zeros(4, dtype=torch.int32, device=device)     # [2] = chain base, [3] = timeout flag
        self._bufs = (ll, ptrs, state, self.max_rows * max_n * 4)
        return self._bufs

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [M, in_features of the first layer] fp16, replicated on every rank -> [M, out_features of the last]."""
        lib = capi.load()
        m = x.shape[0]
        if m > self.max_rows or x.dtype != torch.float16 or not x.is_contiguous():
            raise ValueError("x must be a contiguous fp16 [M <= max_rows, K] tensor")
        ll, ptrs, state, stride = self._buffers(x.device)
        st = torch.cuda.current_stream().cuda_stream
        src = x.data_ptr()
        for i, (q, sc, z, k, n) in enumerate(self.layers):
            b = (self._first_buf + i) % self.nbuf
            outs = (ctypes.c_void_p * self.world)(*[p + b * stride for p in ptrs])
            n_local = n // self.world
            capi.check(lib.xbit_gemv_f16_peers_ll(src, q.data_ptr(), sc.data_ptr(), z.data_ptr(), outs, state.data_ptr(), i,
                                                  self.world, self.rank, m, k, n_local, self.bits, self.groupsize,
                                                  int(self.add_zero_bias), n, self.rank * n_local,
                                                  capi.GEMV_AUTO | (capi.GEMV_FLAG_A_IS_LL if i else 0), st))
            src = ptrs[self.rank] + b * stride
        n_last = self.layers[-1][4]
        out = torch.empty((m, n_last), dtype=torch.float16, device=x.device)
        capi.check(lib.xbit_ll_unpack_f16(src, out.data_ptr(), m * n_last, state.data_ptr(), len(self.layers),
                                          state.data_ptr() + 12, st))
        if len(self.layers) == 1:
            self._first_buf ^= 1                             # never the buffer a slower rank may still be unpacking
        return out

    def check_timeout(self) -> None:
        """Synchronises and raises if a slot of any call so far did not arrive within the spin guard (a peer died or
        never issued its call); clears the flag.  The kernels never hang, but their results are then garbage."""
        if self._bufs is None:
            return
        state = self._bufs[2]
        if int(state[3].item()) != 0:
            state[3] = 0
            raise RuntimeError("xbitops_b200: a flag-in-data slot never arrived (peer rank missing or out of step)")

    __call__ = forward
