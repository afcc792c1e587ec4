"""Drop-in quantised linear layer over the two ops, and the packer it needs.

The reference ships neither: its callers are GPTQ-style ``nn.Module``s in other repositories that hold
``qweight / scales / qzeros (/ g_idx)`` buffers and pick ``XbitOps.gemv`` for decode batches and
``XbitOps.dequant`` + ``matmul`` for prefill (README.md:2; the act-order kernel the author was heading towards is the
``#if 0`` block at /root/reference/src/cu/quant_cuda_kernel.cu:257-336).  This module is that caller:

  pack_stream / unpack_stream    LSB-first bit streams on the device (the layout of synth.py / oracle.unpack_*:
                                 qweight along K per column, qzeros along N per group row), any width 2..8;
  QLinear                        buffers in the reference's layout, forward = gemv for <= `gemv_max_rows` rows,
                                 dequant + fp16 matmul above; optional bias; act-order (`g_idx`) support.

Act-order: GPTQ with ``desc_act`` quantises the input channels in a permuted order and records, per original channel
k, its scale group ``g_idx[k]``.  The kernels want groups that are contiguous in k, so at load time the packed rows
are re-ordered ONCE by the stable permutation that sorts ``g_idx`` (unpack -> index -> repack, on the device) and the
forward pass gathers the activations with the same permutation: ``y = x[:, perm] @ DQ(W[perm, :])`` -- the algebra of
the reference kernel's per-row group lookup, without a gather in the hot loop.
"""
from __future__ import annotations

import torch

from . import ops


def _ceil_div(a: int, b: int) -> int:
    return (a + b - 1) // b


def pack_stream(values: torch.Tensor, bits: int, dim: int) -> torch.Tensor:
    """b-bit unsigned ``values`` -> int32 words of LSB-first streams running along ``dim`` (ceil(count*bits/32) words)."""
    if not 2 <= bits <= 8:
        raise ValueError("bits must be in [2, 8]")
    v = values.movedim(dim, 0).to(torch.int64) & ((1 << bits) - 1)
    count = v.shape[0]
    nwords = _ceil_div(count * bits, 32)
    rest = v.shape[1:]
    # every value as `bits` single bits, LSB first, then 32 consecutive stream bits per word
    shifts = torch.arange(bits, device=v.device, dtype=torch.int64).view(1, bits, *([1] * len(rest)))
    stream = ((v.unsqueeze(1) >> shifts) & 1).reshape(count * bits, *rest)
    pad = nwords * 32 - count * bits
    if pad:
        stream = torch.cat([stream, stream.new_zeros((pad, *rest))], 0)
    weights = (torch.ones((), device=v.device, dtype=torch.int64) << torch.arange(32, device=v.device, dtype=torch.int64))
    words = (stream.reshape(nwords, 32, *rest) * weights.view(1, 32, *([1] * len(rest)))).sum(1)
    words = torch.where(words >= 2**31, words - 2**32, words).to(torch.int32)
    return words.movedim(0, dim).contiguous()


def unpack_stream(words: torch.Tensor, count: int, bits: int, dim: int) -> torch.Tensor:
    """Inverse of pack_stream: the first ``count`` b-bit values of the streams along ``dim`` (uint8)."""
    w = words.movedim(dim, 0).to(torch.int64) & 0xFFFFFFFF
    rest = w.shape[1:]
    shifts = torch.arange(32, device=w.device, dtype=torch.int64).view(1, 32, *([1] * len(rest)))
    stream = ((w.unsqueeze(1) >> shifts) & 1).reshape(w.shape[0] * 32, *rest)[: count * bits]
    weights = (torch.ones((), device=w.device, dtype=torch.int64) << torch.arange(bits, device=w.device, dtype=torch.int64))
    vals = (stream.reshape(count, bits, *rest) * weights.view(1, bits, *([1] * len(rest)))).sum(1)
    return vals.to(torch.uint8).movedim(0, dim).contiguous()


def pack_qweight(w_int: torch.Tensor, bits: int) -> torch.Tensor:
    """uint8 [K, N] -> qweight int32 [ceil(K*bits/32), N]."""
    return pack_stream(w_int, bits, 0)


def pack_qzeros(z_int: torch.Tensor, bits: int) -> torch.Tensor:
    """uint8 [G, N] (stored zeros = effective zero - add_zero_bias) -> qzeros int32 [G, ceil(N*bits/32)]."""
    return pack_stream(z_int, bits, 1)


def quantize_rtn(weight_kn: torch.Tensor, bits: int, groupsize: int, add_zero_bias: int = 1):
    """Asymmetric round-to-nearest group quantisation of a [K, N] matrix (K a multiple of the group size) into the packed
    format: enough of a quantiser for examples and tests, not a GPTQ solver.  -> (qweight, scales fp16, qzeros)."""
    K, N = weight_kn.shape
    if K % groupsize:
        raise ValueError("in_features must be a multiple of the group size")
    qmax = (1 << bits) - 1
    wg = weight_kn.float().reshape(K // groupsize, groupsize, N)
    lo, hi = wg.amin(1), wg.amax(1)
    scale = ((hi - lo) / qmax).clamp_min(1e-6).to(torch.float16)
    s32 = scale.float()
    zero = torch.clamp(torch.round(-lo / s32), add_zero_bias, qmax + add_zero_bias)
    q = torch.clamp(torch.round(wg / s32[:, None, :] + zero[:, None, :]), 0, qmax).to(torch.uint8)
    return pack_qweight(q.reshape(K, N), bits), scale, pack_qzeros((zero - add_zero_bias).to(torch.uint8), bits)


class QLinear(torch.nn.Module):
    """``y = x @ DQ(qweight, scales, qzeros) (+ bias)`` with the reference's buffer layout (W is stored [in, out])."""

    def __init__(self, in_features: int, out_features: int, bits: int = 4, groupsize: int = 128, bias: bool = False,
                 add_zero_bias: int = 1, gemv_max_rows: int = 16, dtype: torch.dtype = torch.float16, device=None):
        super().__init__()
        if dtype not in (torch.float16, torch.bfloat16):
            raise ValueError("dtype must be float16 or bfloat16")
        self.in_features, self.out_features = in_features, out_features
        self.bits, self.groupsize, self.add_zero_bias, self.gemv_max_rows = bits, groupsize, add_zero_bias, gemv_max_rows
        groups = _ceil_div(in_features, groupsize)
        self.register_buffer("qweight", torch.zeros((_ceil_div(in_features * bits, 32), out_features), dtype=torch.int32, device=device))
        self.register_buffer("scales", torch.ones((groups, out_features), dtype=dtype, device=device))
        self.register_buffer("qzeros", torch.zeros((groups, _ceil_div(out_features * bits, 32)), dtype=torch.int32, device=device))
        self.register_buffer("bias", torch.zeros(out_features, dtype=dtype, device=device) if bias else None)
        self.register_buffer("perm", None)          # act-order: activations are gathered with it (see the module docstring)

    @classmethod
    def from_packed(cls, qweight, scales, qzeros, bits: int, groupsize: int, in_features: int, bias=None, add_zero_bias: int = 1,
                    g_idx=None, **kw) -> "QLinear":
        """Adopt GPTQ-layout tensors; ``g_idx`` (int [in_features], group of every input channel) enables act-order."""
        m = cls(in_features, qweight.shape[1], bits, groupsize, bias is not None, add_zero_bias, dtype=scales.dtype, device=qweight.device, **kw)
        m.qweight, m.scales, m.qzeros = qweight.contiguous(), scales.contiguous(), qzeros.contiguous()
        if bias is not None:
            m.bias = bias.to(scales.dtype)
        if g_idx is not None:
            g_idx = g_idx.to(qweight.device).long()
            natural = torch.arange(in_features, device=g_idx.device) // groupsize
            if not torch.equal(g_idx, natural):
                counts = torch.bincount(g_idx, minlength=_ceil_div(in_features, groupsize))
                if int(counts[:-1].min()) != groupsize or int(counts.max()) != groupsize:
                    raise ValueError("g_idx must assign exactly `groupsize` input channels to every (full) group")
                perm = torch.argsort(g_idx, stable=True)
                rows = unpack_stream(m.qweight, in_features, bits, 0)        # [K, N] uint8, original channel order
                m.qweight = pack_qweight(rows[perm], bits)                      # groups contiguous in k
                m.perm = perm
        return m

    @classmethod
    def from_linear(cls, linear: torch.nn.Linear, bits: int = 4, groupsize: int = 128, add_zero_bias: int = 1, **kw) -> "QLinear":
        """Round-to-nearest quantisation of an ``nn.Linear`` (weight [out, in] -> W^T [in, out])."""
        w = linear.weight.detach().t().contiguous()
        qw, sc, qz = quantize_rtn(w, bits, groupsize, add_zero_bias)
        dt = linear.weight.dtype if linear.weight.dtype in (torch.float16, torch.bfloat16) else torch.float16
        return cls.from_packed(qw, sc.to(dt), qz, bits, groupsize, linear.in_features,
                               None if linear.bias is None else linear.bias.detach(), add_zero_bias, **kw)

    def dequantized_weight(self) -> torch.Tensor:
        """[in_features, out_features] in the layer's dtype, in the ORIGINAL input-channel order."""
        w = ops.dequant(self.qweight, self.scales, self.qzeros, self.groupsize, self.bits, self.in_features, self.add_zero_bias)
        if self.perm is not None:
            inv = torch.empty_like(self.perm)
            inv[self.perm] = torch.arange(self.in_features, device=self.perm.device)
            w = w[inv]
        return w

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        lead = x.shape[:-1]
        x2 = x.reshape(-1, self.in_features)
        if self.perm is not None:
            x2 = x2[:, self.perm]
        rows = x2.shape[0]
        if rows <= self.gemv_max_rows:
            # (bf16 layers keep bf16 activations when the bf16-native kernels are switched on, ops.set_native_bf16; the
            # op falls back to the reference's fp16 arithmetic where no native kernel exists)
            native = (ops.get_native_bf16() and self.scales.dtype == torch.bfloat16 and self.bits in (2, 4, 8) and self.groupsize == 128
                      and self.in_features % 128 == 0 and self.out_features % 32 == 0 and self.in_features <= 16384)
            xin = x2.to(torch.bfloat16 if native else torch.float16).contiguous()
            y = ops.gemv(xin, self.qweight, self.scales, self.qzeros, self.groupsize, self.bits,
                         self.in_features, self.add_zero_bias)
        else:
            w = ops.dequant(self.qweight, self.scales, self.qzeros, self.groupsize, self.bits, self.in_features, self.add_zero_bias)
            y = x2.to(w.dtype) @ w
        y = y.to(self.scales.dtype)
        if self.bias is not None:
            y = y + self.bias
        return y.reshape(*lead, self.out_features)

    def extra_repr(self) -> str:
        return (f"in_features={self.in_features}, out_features={self.out_features}, bits={self.bits}, groupsize={self.groupsize}, "
                f"bias={self.bias is not None}, act_order={self.perm is not None}")
