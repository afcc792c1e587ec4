"""Synthetic GPTQ-layout inputs and the host-side packer.

The packed formats are the reference's (/root/reference
src/cu/unpack_weight_2_to_7.cu:53-66,196-217,256-281 and src/dq_torch_ops.cc:31):

  qweight int32 [ceil(K*b/32), N]   value k of column n = bits [k*b, k*b+b) of the LSB-first
                                    stream formed by qweight[:, n]
  scales  fp16  [G, N]              G = ceil(K/g)
  qzeros  int32 [G, ceil(N*b/32)]   zero of (group, n) = bits [n*b, n*b+b) of row `group`;
                                    effective zero = z + add_zero_bias

The reference ships no packer (its callers live in other repositories); this numpy one is the
first row of SURVEY.md 8(f).  Inputs follow SURVEY.md 8(d): PCG64 seeds 0/1/2/3 for
qweight/qzeros/scales/activations.
"""
from __future__ import annotations

import numpy as np


def ceil_div(a: int, b: int) -> int:
    return (a + b - 1) // b


def pack_stream(values: np.ndarray, bits: int, axis: int) -> np.ndarray:
    """Pack b-bit unsigned ``values`` into LSB-first uint32 streams running along ``axis``.
    Returns int32 words with ``ceil(count*bits/32)`` entries along that axis."""
    v = np.moveaxis(np.asarray(values), axis, 0).astype(np.uint64) & np.uint64((1 << bits) - 1)
    count = v.shape[0]
    nwords = ceil_div(count * bits, 32)
    out = np.zeros((nwords + 1,) + v.shape[1:], np.uint64)
    pos = np.arange(count, dtype=np.int64) * bits
    wi, sh = pos >> 5, (pos & 31).astype(np.uint64)
    shifted = v << sh.reshape((-1,) + (1,) * (v.ndim - 1))
    np.bitwise_or.at(out, wi, shifted & np.uint64(0xFFFFFFFF))
    np.bitwise_or.at(out, wi + 1, shifted >> np.uint64(32))
    words = out[:nwords].astype(np.uint32).view(np.int32)
    return np.ascontiguousarray(np.moveaxis(words, 0, axis))


def pack_qweight(w: np.ndarray, bits: int) -> np.ndarray:
    """w uint8 [K, N] -> qweight int32 [ceil(K*b/32), N]."""
    return pack_stream(w, bits, axis=0)


def pack_qzeros(z: np.ndarray, bits: int) -> np.ndarray:
    """z uint8 [G, N] (raw stored zeros, i.e. effective zero minus add_zero_bias)
    -> qzeros int32 [G, ceil(N*b/32)]."""
    return pack_stream(z, bits, axis=1)


def quantize(weight_kn: np.ndarray, bits: int, groupsize: int, add_zero_bias: int = 0):
    """Asymmetric round-to-nearest group quantisation of an fp [K, N] matrix into the packed
    format (enough of a quantiser to build end-to-end examples; not a GPTQ solver).
    Returns (qweight, scales_f16, qzeros)."""
    K, N = weight_kn.shape
    G = ceil_div(K, groupsize)
    qmax = (1 << bits) - 1
    pad = G * groupsize - K
    wpad = np.concatenate([weight_kn, np.repeat(weight_kn[-1:], pad, 0)], 0) if pad else weight_kn
    wg = wpad.astype(np.float32).reshape(G, groupsize, N)
    lo, hi = wg.min(1), wg.max(1)
    scale = np.maximum((hi - lo) / qmax, 1e-6).astype(np.float16)
    s32 = scale.astype(np.float32)
    zero = np.clip(np.rint(-lo / s32), add_zero_bias, qmax + add_zero_bias)
    q = np.clip(np.rint(wg / s32[:, None, :] + zero[:, None, :]), 0, qmax).astype(np.uint8)
    qweight = pack_qweight(q.reshape(G * groupsize, N)[:K], bits)
    qzeros = pack_qzeros((zero - add_zero_bias).astype(np.uint8), bits)
    return qweight, scale, qzeros


def make_inputs(K: int, N: int, bits: int = 4, groupsize: int = 128, M: int = 1, seed: int = 0,
                scale_mode: str = "gptq"):
    """Synthetic tensors of SURVEY.md 8(d). Every bit pattern is a valid packing, so qweight and
    qzeros are uniform random words. ``scale_mode``: "gptq" = fp16(uniform(0.002, 0.02));
    "bits" = random fp16 bit patterns with |s| < 64, incl. subnormals, negatives and +-0; "ones" = 1.0."""
    G = ceil_div(K, groupsize)
    rng_w = np.random.Generator(np.random.PCG64(seed))
    rng_z = np.random.Generator(np.random.PCG64(seed + 1))
    rng_s = np.random.Generator(np.random.PCG64(seed + 2))
    rng_a = np.random.Generator(np.random.PCG64(seed + 3))
    qweight = rng_w.integers(0, 1 << 32, size=(ceil_div(K * bits, 32), N), dtype=np.uint64).astype(np.uint32).view(np.int32)
    qzeros = rng_z.integers(0, 1 << 32, size=(G, ceil_div(N * bits, 32)), dtype=np.uint64).astype(np.uint32).view(np.int32)
    if scale_mode == "gptq":
        scales = rng_s.uniform(0.002, 0.02, size=(G, N)).astype(np.float16)
    elif scale_mode == "bits":
        raw = rng_s.integers(0, 1 << 16, size=(G, N), dtype=np.uint32).astype(np.uint16)
        # keep |s| < 64 (exponent field <= 20) so (z+bias)*s and w*s - sz stay finite in fp16
        raw = np.where(((raw >> 10) & 0x1F) > 20, raw & 0x83FF, raw).astype(np.uint16)
        scales = raw.view(np.float16)
    elif scale_mode == "ones":
        scales = np.ones((G, N), np.float16)
    else:
        raise ValueError(scale_mode)
    a = rng_a.standard_normal(size=(M, K)).astype(np.float16)
    return qweight, scales, qzeros, a


def gemv_bytes(K: int, N: int, bits: int, groupsize: int, M: int = 1) -> int:
    """Algorithmic bytes of one GEMV call; weights counted once regardless of M."""
    G = ceil_div(K, groupsize)
    return ceil_div(K * bits, 32) * N * 4 + G * N * 2 + G * ceil_div(N * bits, 32) * 4 + M * K * 2 + M * N * 2


def dq_bytes(K: int, N: int, bits: int, groupsize: int) -> int:
    """Algorithmic bytes of one dequant call."""
    G = ceil_div(K, groupsize)
    return ceil_div(K * bits, 32) * N * 4 + G * N * 2 + G * ceil_div(N * bits, 32) * 4 + K * N * 2
