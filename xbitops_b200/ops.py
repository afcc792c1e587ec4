"""Host-side mirror of the reference's operator interface, over the C ABI.

Same names, positional order, shapes, dtypes and error behaviour (RuntimeError) as the reference's
extension module (/root/reference/src/dq_torch_ops.cc:23-44 `dequant`, :46-78 `gemv`, registered at
:80-85).  PyTorch is used for device memory and streams only; all arithmetic is in
libxbitops_b200.so.  The compiled twin of this file is csrc/dq_torch_ops.cc (module `XbitOps`).
"""
from __future__ import annotations

import os

import torch

from . import capi


def _check_input(x: torch.Tensor, name: str) -> None:
    # dq_torch_ops.cc:5-9
    if not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not x.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def _check_quant_args(qweight, scales, qzeros, groupsize: int, bits: int, in_features: int) -> None:
    _check_input(qweight, "qweight")
    _check_input(scales, "scales")
    _check_input(qzeros, "qzeros")
    # dq_torch_ops.cc:28-31
    if qweight.dim() != 2:
        raise RuntimeError("qweight must be 2-dimensional")
    if groupsize < 16:
        raise RuntimeError("groupsize must be >= 16")
    if not (1 <= bits <= 8):
        raise RuntimeError("bits must be >= 1 and <= 8")
    if (in_features * bits + 31) // 32 != qweight.size(0):
        raise RuntimeError("in_features must be >= 1")
    # explicit versions of what the reference only enforces through data_ptr<T>() throwing
    if qweight.dtype != torch.int32 or qzeros.dtype != torch.int32:
        raise RuntimeError("qweight and qzeros must be int32")
    if scales.dtype not in (torch.float16, torch.bfloat16):
        raise RuntimeError("scales must be float16 or bfloat16")
    n = qweight.size(1)
    groups = (in_features + groupsize - 1) // groupsize
    if scales.dim() != 2 or scales.size(0) < groups or scales.size(1) != n:
        raise RuntimeError("scales must be [ceil(in_features/groupsize), out_features]")
    if qzeros.dim() != 2 or qzeros.size(0) < groups or qzeros.size(1) != (n * bits + 31) // 32:
        raise RuntimeError("qzeros must be [ceil(in_features/groupsize), ceil(out_features*bits/32)]")
    if scales.device != qweight.device or qzeros.device != qweight.device:
        raise RuntimeError("qweight, scales and qzeros must be on the same device")


def _stream_handle() -> int:
    return torch.cuda.current_stream().cuda_stream


_WORKSPACES: dict = {}
_STATIC_WEIGHTS = os.environ.get("XBIT_STATIC_WEIGHTS", "0") not in ("", "0")


def set_static_weights(on: bool) -> None:
    """The caller's promise that qweight / scales / qzeros passed to gemv are resident model weights, never written by
    the kernel that precedes the call in its stream (default False; env XBIT_STATIC_WEIGHTS=1 at import).  gemv then
    prefetches them under programmatic dependent launch while that kernel drains (XBIT_GEMV_FLAG_STATIC_WEIGHTS);
    activations are still read only after it has finished.  Same switch as XbitOps.set_static_weights of the twin."""
    global _STATIC_WEIGHTS
    _STATIC_WEIGHTS = bool(on)


def get_static_weights() -> bool:
    return _STATIC_WEIGHTS


_NATIVE_BF16 = os.environ.get("XBIT_NATIVE_BF16", "0") not in ("", "0")


def set_native_bf16(on: bool) -> None:
    """bf16 without the fp16 round trip (default False = the reference's behaviour, which converts
    bf16 scales to fp16, computes in fp16 and casts the result back, dq_torch_ops.cc:33-42, :65-76 -- and loses bf16's
    range on the way).  When on, `dequant` with bf16 scales returns RN_bf16((w - z) * s) (xbit_dequant_bf16), and `gemv`
    with bf16 activations AND bf16 scales runs the bf16-native kernel where it exists (bits 4, groupsize 128,
    K % 128 = 0, N % 32 = 0: xbit_gemv_bf16); every other case keeps the reference's fp16 arithmetic."""
    global _NATIVE_BF16
    _NATIVE_BF16 = bool(on)


def get_native_bf16() -> bool:
    return _NATIVE_BF16


def gemv_workspace(device: torch.device) -> torch.Tensor:
    """Zero-initialised scratch for the persistent stream-K GEMV schedule (include/xbitops_b200.h:
    xbit_gemv_workspace_bytes), one per (device, stream): calls on one stream are ordered, calls on
    different streams must not share it.  Every call leaves it zeroed again."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream_handle())
    ws = _WORKSPACES.get(key)
    if ws is None:
        nbytes = capi.load().xbit_gemv_workspace_bytes(16, 0, 0, 4, 128)
        ws = torch.zeros(max(nbytes, 256), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def dequant(qweight, scales, qzeros, groupsize, bits, in_features, add_zero_bias):
    """-> Tensor[in_features, out_features] in scales.dtype (i.e. W^T of nn.Linear.weight)."""
    _check_quant_args(qweight, scales, qzeros, groupsize, bits, in_features)
    lib = capi.load()
    with torch.cuda.device(qweight.device):
        if _NATIVE_BF16 and scales.dtype == torch.bfloat16:
            out = torch.empty((in_features, qweight.size(1)), dtype=torch.bfloat16, device=qweight.device)
            capi.check(lib.xbit_dequant_bf16(qweight.data_ptr(), scales.data_ptr(), qzeros.data_ptr(), out.data_ptr(),
                                             in_features, qweight.size(1), bits, groupsize, int(add_zero_bias),
                                             _stream_handle()))
            return out
        f16_scale = scales.to(torch.float16) if scales.dtype == torch.bfloat16 else scales
        out = torch.empty((in_features, qweight.size(1)), dtype=torch.float16, device=qweight.device)
        capi.check(lib.xbit_dequant_f16(qweight.data_ptr(), f16_scale.data_ptr(), qzeros.data_ptr(), out.data_ptr(),
                                        in_features, qweight.size(1), bits, groupsize, int(add_zero_bias),
                                        _stream_handle()))
        if scales.dtype == torch.bfloat16:
            out = out.to(torch.bfloat16)
    return out


def gemv(input_a, qweight, scales, qzeros, groupsize, bits, in_features, add_zero_bias, *, family: int = capi.GEMV_AUTO,
         out: torch.Tensor | None = None):
    """-> Tensor[M, N] (or [B, S, N] for 3-D activations) in scales.dtype.
    `family` / `out` are keyword-only extensions (kernel-family override for the crossover sweep,
    caller-owned output for CUDA-graph capture); the positional surface is the reference's."""
    _check_input(input_a, "input_a")
    _check_quant_args(qweight, scales, qzeros, groupsize, bits, in_features)
    if qweight.device.index != input_a.device.index:
        raise RuntimeError("input and weight must be on the same device")
    native = (_NATIVE_BF16 and input_a.dtype == torch.bfloat16 and scales.dtype == torch.bfloat16 and bits in (2, 4, 8) and
              groupsize == 128 and in_features % 128 == 0 and qweight.size(1) % 32 == 0 and family == capi.GEMV_AUTO and
              in_features <= 16384)
    if input_a.dtype != torch.float16 and not native:
        raise RuntimeError("input_a must be float16")
    if input_a.dim() < 2 or input_a.size(-1) != in_features:
        raise RuntimeError("input_a must be [..., in_features]")
    lib = capi.load()
    n = qweight.size(1)
    outshape = [input_a.size(0), n]
    m = input_a.size(0)
    if input_a.dim() > 2:                      # dq_torch_ops.cc:60-64
        outshape.insert(1, input_a.size(1))
        m *= input_a.size(1)
    if native:
        with torch.cuda.device(qweight.device):
            if out is None:
                out = torch.empty(outshape, dtype=torch.bfloat16, device=qweight.device)
            elif out.dtype != torch.bfloat16 or not out.is_contiguous() or list(out.shape) != outshape:
                raise RuntimeError("out must be a contiguous bfloat16 tensor of the result shape")
            if m > 0:
                ws = gemv_workspace(qweight.device)
                capi.check(lib.xbit_gemv_bf16(input_a.data_ptr(), qweight.data_ptr(), scales.data_ptr(), qzeros.data_ptr(),
                                              out.data_ptr(), m, in_features, n, bits, groupsize, int(add_zero_bias), n,
                                              ws.data_ptr(), ws.numel(),
                                              capi.GEMV_FLAG_STATIC_WEIGHTS if _STATIC_WEIGHTS else 0, _stream_handle()))
        return out
    with torch.cuda.device(qweight.device):
        f16_scale = scales.to(torch.float16) if scales.dtype == torch.bfloat16 else scales
        if out is None:
            out16 = torch.empty(outshape, dtype=torch.float16, device=qweight.device)
        else:
            if out.dtype != torch.float16 or not out.is_contiguous() or list(out.shape) != outshape:
                raise RuntimeError("out must be a contiguous float16 tensor of the result shape")
            out16 = out
        if m > 0:
            ws = gemv_workspace(qweight.device)
            capi.check(lib.xbit_gemv_f16_ex(input_a.data_ptr(), qweight.data_ptr(), f16_scale.data_ptr(),
                                            qzeros.data_ptr(), out16.data_ptr(), m, in_features, n, bits, groupsize,
                                            int(add_zero_bias), n, ws.data_ptr(), ws.numel(),
                                            int(family) | (capi.GEMV_FLAG_STATIC_WEIGHTS if _STATIC_WEIGHTS else 0),
                                            _stream_handle()))
        if scales.dtype == torch.bfloat16:
            return out16.to(torch.bfloat16)
    return out16


def gemv_multi(input_a, projections, groupsize, bits, in_features, add_zero_bias, *, family: int = capi.GEMV_AUTO):
    """Several weight matrices applied to ONE activation tensor (Q/K/V, gate + up): `projections` is a sequence of
    (qweight, scales, qzeros) triples; returns the list [gemv(input_a, q, s, z, ...) for (q, s, z) in projections],
    bit-identical to those calls, through ONE launch where the persistent W4 schedule applies to all of them
    (xbit_gemv_f16_multi).  The reference has one op call per projection (dq_torch_ops.cc:46-78); this is the
    multi-projection entry of SURVEY.md 8(f)."""
    import ctypes
    projections = list(projections)
    if not 1 <= len(projections) <= 4:
        raise RuntimeError("gemv_multi takes 1 to 4 projections")
    _check_input(input_a, "input_a")
    if input_a.dtype != torch.float16 or input_a.dim() != 2 or input_a.size(1) != in_features:
        raise RuntimeError("input_a must be a float16 [M, in_features] tensor")
    dt = projections[0][1].dtype
    outs, keep = [], []
    arr = (capi.GemvProblem * len(projections))()
    m = input_a.size(0)
    for i, (q, s, z) in enumerate(projections):
        _check_quant_args(q, s, z, groupsize, bits, in_features)
        if q.device.index != input_a.device.index or s.dtype != dt:
            raise RuntimeError("all projections must live on the activation's device and share one scales dtype")
        s16 = s.to(torch.float16) if s.dtype == torch.bfloat16 else s
        o = torch.empty((m, q.size(1)), dtype=torch.float16, device=q.device)
        keep.append(s16)
        outs.append(o)
        arr[i] = capi.GemvProblem(q.data_ptr(), s16.data_ptr(), z.data_ptr(), o.data_ptr(), q.size(1), q.size(1))
    if m > 0:
        with torch.cuda.device(input_a.device):
            ws = gemv_workspace(input_a.device)
            capi.check(capi.load().xbit_gemv_f16_multi(input_a.data_ptr(), ctypes.cast(arr, ctypes.c_void_p), len(projections), m,
                                                       in_features, bits, groupsize, int(add_zero_bias), ws.data_ptr(), ws.numel(),
                                                       int(family) | (capi.GEMV_FLAG_STATIC_WEIGHTS if _STATIC_WEIGHTS else 0),
                                                       _stream_handle()))
    return [o.to(torch.bfloat16) for o in outs] if dt == torch.bfloat16 else outs
