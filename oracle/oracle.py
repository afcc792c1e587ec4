"""TEST INFRASTRUCTURE ONLY -- the parity checker, never the product.

Python face of the oracle:
  * ``COracle``      ctypes binding of oracle/xbit_oracle.c (our plain-C restatement);
  * ``np_*``         an independent numpy restatement of the same semantics (cross-checks the C);
  * ``RefCpu``       ctypes binding of oracle/_ref/libxbit_refcpu.so = the UNMODIFIED reference CPU
                     simulator (/root/reference/src/cpp_simulate.cc:568-691) behind ref_cpu_driver.cc;
  * ``load_ref_gpu`` the UNMODIFIED reference PyTorch extension built by oracle/build_ref_gpu.sh.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Nothing under xbitops_b200/ does (tests/test_no_oracle_in_product.py enforces it).

Reference semantics restated (paths relative to /root/reference):
  formats  LSB-first bit streams along K per column (qweight) / along N per group row (qzeros)
           -- src/cu/unpack_weight_2_to_7.cu:53-66,196-217,256-281; src/dq_torch_ops.cc:31
  dequant  sz = RN16(RN16(z+bias)*s); out = RN16(RN16(w)*s - sz)  -- unpack_weight_2_to_7.cu:58-61,72-75
  gemv     y = RN16(sum_k a_k * out[k, n]) with wide accumulation  -- src/cu/gemv_w4a16_pt.cu:35-145
"""
from __future__ import annotations

import ctypes
import importlib.util
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_ORACLE = HERE / "_build" / "libxbit_oracle.so"
LIB_REFCPU = HERE / "_ref" / "libxbit_refcpu.so"
DIR_REFGPU = HERE / "_ref" / "refgpu"

_i32p = ctypes.POINTER(ctypes.c_int32)
_u16p = ctypes.POINTER(ctypes.c_uint16)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> None:
    """Compile the C restatement and (when /root/reference is present) the reference CPU driver."""
    if force or not LIB_ORACLE.exists() or LIB_ORACLE.stat().st_mtime < (HERE / "xbit_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "liboracle", "CC=gcc"], check=True, capture_output=True)
    if Path("/root/reference/src/cpp_simulate.cc").exists() and (force or not LIB_REFCPU.exists()):
        subprocess.run(["make", "-C", str(HERE), "refcpu", "CXX=g++"], check=True, capture_output=True)


def _p(a: np.ndarray, ty):
    assert a.flags["C_CONTIGUOUS"], "oracle expects contiguous arrays"
    return a.ctypes.data_as(ty)


def ceil_div(a: int, b: int) -> int:
    return (a + b - 1) // b


class COracle:
    """ctypes binding of oracle/xbit_oracle.c. All arrays are host numpy arrays."""

    def __init__(self):
        build()
        self.lib = ctypes.CDLL(str(LIB_ORACLE))
        L = self.lib
        L.xo_unpack_qweight.argtypes = [_i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _u8p]
        L.xo_unpack_qzeros.argtypes = [_i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _u8p]
        L.xo_dequant_f16.argtypes = [_i32p, _u16p, _i32p, _u16p] + [ctypes.c_int] * 5
        L.xo_dequant_f16_rows.argtypes = [_i32p, _u16p, _i32p, _u16p] + [ctypes.c_int] * 7
        L.xo_gemv_from_dq.argtypes = [_u16p, _u16p, _f64p, _u16p] + [ctypes.c_int] * 3
        L.xo_gemv_from_dq_cols.argtypes = [_u16p, _u16p, _f64p, _u16p] + [ctypes.c_int] * 5
        L.xo_gemv_w4_ref_arith.argtypes = [_u16p, _i32p, _u16p, _i32p, _u16p] + [ctypes.c_int] * 5
        L.xo_pack_qweight.argtypes = [_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i32p]
        L.xo_pack_qzeros.argtypes = [_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i32p]
        L.xo_f64_to_f16.argtypes = [ctypes.c_double]
        L.xo_f64_to_f16.restype = ctypes.c_uint16
        L.xo_f16_to_f64.argtypes = [ctypes.c_uint16]
        L.xo_f16_to_f64.restype = ctypes.c_double
        for f in (L.xo_unpack_qweight, L.xo_unpack_qzeros, L.xo_dequant_f16, L.xo_dequant_f16_rows,
                  L.xo_gemv_from_dq, L.xo_gemv_from_dq_cols, L.xo_gemv_w4_ref_arith,
                  L.xo_pack_qweight, L.xo_pack_qzeros):
            f.restype = None

    # ---- integer level
    def unpack_qweight(self, qweight: np.ndarray, K: int, bits: int) -> np.ndarray:
        N = qweight.shape[1]
        out = np.empty((K, N), np.uint8)
        self.lib.xo_unpack_qweight(_p(qweight, _i32p), K, N, bits, _p(out, _u8p))
        return out

    def unpack_qzeros(self, qzeros: np.ndarray, N: int, bits: int) -> np.ndarray:
        G = qzeros.shape[0]
        out = np.empty((G, N), np.uint8)
        self.lib.xo_unpack_qzeros(_p(qzeros, _i32p), G, N, bits, _p(out, _u8p))
        return out

    def pack_qweight(self, w: np.ndarray, bits: int) -> np.ndarray:
        K, N = w.shape
        q = np.zeros((ceil_div(K * bits, 32), N), np.int32)
        self.lib.xo_pack_qweight(_p(np.ascontiguousarray(w, np.uint8), _u8p), K, N, bits, _p(q, _i32p))
        return q

    def pack_qzeros(self, z: np.ndarray, bits: int) -> np.ndarray:
        G, N = z.shape
        q = np.zeros((G, ceil_div(N * bits, 32)), np.int32)
        self.lib.xo_pack_qzeros(_p(np.ascontiguousarray(z, np.uint8), _u8p), G, N, bits, _p(q, _i32p))
        return q

    # ---- dequant (fp16 bit patterns as uint16 / float16 views)
    def dequant(self, qweight, scales, qzeros, groupsize, bits, K, add_zero_bias, k_range=None) -> np.ndarray:
        N = qweight.shape[1]
        s = np.ascontiguousarray(scales).view(np.uint16)
        out = np.zeros((K, N), np.uint16)
        k0, k1 = k_range if k_range is not None else (0, K)
        self.lib.xo_dequant_f16_rows(_p(qweight, _i32p), _p(s, _u16p), _p(qzeros, _i32p), _p(out, _u16p),
                                     K, N, bits, groupsize, int(add_zero_bias), k0, k1)
        return out.view(np.float16)

    # ---- gemv truth: returns (y64 [M,N] float64, y16 [M,N] float16)
    def gemv_from_dq(self, a: np.ndarray, w_f16: np.ndarray, n_range=None):
        a = np.ascontiguousarray(a)
        M, K = a.shape
        N = w_f16.shape[1]
        y64 = np.zeros((M, N), np.float64)
        y16 = np.zeros((M, N), np.uint16)
        n0, n1 = n_range if n_range is not None else (0, N)
        self.lib.xo_gemv_from_dq_cols(_p(a.view(np.uint16), _u16p), _p(w_f16.view(np.uint16), _u16p),
                                      _p(y64, _f64p), _p(y16, _u16p), M, K, N, n0, n1)
        return y64, y16.view(np.float16)

    def gemv(self, a, qweight, scales, qzeros, groupsize, bits, K, add_zero_bias):
        w = self.dequant(qweight, scales, qzeros, groupsize, bits, K, add_zero_bias)
        return self.gemv_from_dq(a.reshape(-1, K), w)

    def gemv_w4_ref_arith(self, a, qweight, scales, qzeros, groupsize, K, add_zero_bias) -> np.ndarray:
        a = np.ascontiguousarray(a).reshape(-1, K)
        M, N = a.shape[0], qweight.shape[1]
        y16 = np.zeros((M, N), np.uint16)
        s = np.ascontiguousarray(scales).view(np.uint16)
        self.lib.xo_gemv_w4_ref_arith(_p(a.view(np.uint16), _u16p), _p(qweight, _i32p), _p(s, _u16p),
                                      _p(qzeros, _i32p), _p(y16, _u16p), M, K, N, groupsize, int(add_zero_bias))
        return y16.view(np.float16)


# --------------------------------------------------------------------------- numpy restatement

def np_unpack_stream(words: np.ndarray, count: int, bits: int, axis: int) -> np.ndarray:
    """Extract ``count`` b-bit values from LSB-first streams of uint32 words running along ``axis``."""
    w = np.moveaxis(np.ascontiguousarray(words).view(np.uint32), axis, 0).astype(np.uint64)
    nwords = w.shape[0]
    pos = np.arange(count, dtype=np.uint64) * np.uint64(bits)
    wi = (pos >> np.uint64(5)).astype(np.int64)
    sh = (pos & np.uint64(31)).reshape((-1,) + (1,) * (w.ndim - 1))
    pad = np.concatenate([w, np.zeros((2,) + w.shape[1:], np.uint64)], axis=0)
    lo = pad[np.minimum(wi, nwords + 1)]
    hi = pad[np.minimum(wi + 1, nwords + 1)]
    both = lo | (hi << np.uint64(32))
    vals = ((both >> sh) & np.uint64((1 << bits) - 1)).astype(np.uint8)
    return np.moveaxis(vals, 0, axis)


def np_unpack_qweight(qweight: np.ndarray, K: int, bits: int) -> np.ndarray:
    return np_unpack_stream(qweight, K, bits, axis=0)           # [K, N]


def np_unpack_qzeros(qzeros: np.ndarray, N: int, bits: int) -> np.ndarray:
    return np_unpack_stream(qzeros, N, bits, axis=1)            # [G, N]


def np_sim_float_to_half(x: np.ndarray) -> np.ndarray:
    """The reference CPU simulator's software fp32->fp16 (cpp_simulate.cc:44-58), restated:
    adds half an fp16 ulp to the magnitude and truncates (ties round AWAY from zero, unlike the
    GPU's round-to-nearest-even), no infinities (saturates to 0x7FFF). Returns uint16 bit patterns."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    b = (u + np.uint64(0x1000)) & np.uint64(0xFFFFFFFF)
    e = ((b & np.uint64(0x7F800000)) >> np.uint64(23)).astype(np.int64)
    m = b & np.uint64(0x007FFFFF)
    sign = (b & np.uint64(0x80000000)) >> np.uint64(16)
    normal = (((np.clip(e - 112, 0, None).astype(np.uint64) << np.uint64(10)) & np.uint64(0x7C00)) | (m >> np.uint64(13)))
    shift = np.clip(125 - e, 0, 63).astype(np.uint64)
    denorm = (((np.uint64(0x007FF000) + m) >> shift) + np.uint64(1)) >> np.uint64(1)
    out = sign | np.where(e > 112, normal, np.uint64(0)) | np.where((e < 113) & (e > 101), denorm, np.uint64(0)) \
        | np.where(e > 143, np.uint64(0x7FFF), np.uint64(0))
    return out.astype(np.uint16)


def np_dequant(qweight, scales, qzeros, groupsize, bits, K, add_zero_bias, rounding: str = "rne") -> np.ndarray:
    """rounding="rne": IEEE RN-even restatement of the GPU arithmetic -- numpy float16 products are
    exact-in-f32-then-rounded, and the fma is evaluated exactly in float64 and rounded once
    (numpy's f64->f16 cast is RN-even).
    rounding="sim": the same formula with the reference CPU simulator's own float_to_half at the
    same two rounding points (cpp_simulate.cc:226-246) -- must equal RefCpu.dequant bit for bit."""
    N = qweight.shape[1]
    w = np_unpack_qweight(qweight, K, bits)
    z = np_unpack_qzeros(qzeros, N, bits).astype(np.int64) + int(add_zero_bias)
    grp = np.arange(K) // groupsize
    s16 = np.asarray(scales).view(np.float16) if scales.dtype != np.float16 else scales
    if rounding == "rne":
        sz = (z.astype(np.float16).astype(np.float32) * s16.astype(np.float32)).astype(np.float16)  # hmul2
        s = s16.astype(np.float64)[grp]                       # [K, N]
        with np.errstate(over="ignore", invalid="ignore"):
            return (w.astype(np.float64) * s + (-sz.astype(np.float64)[grp])).astype(np.float16)
    assert rounding == "sim"
    s32 = s16.astype(np.float32)
    sz = np_sim_float_to_half(z.astype(np.float32) * s32).view(np.float16).astype(np.float32)
    res = w.astype(np.float32) * s32[grp] + np.float32(-1.0) * sz[grp]
    return np_sim_float_to_half(res).view(np.float16)


def np_f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """IEEE round-to-nearest-even fp32 -> bf16 (uint16 bit patterns); inf / nan pass through."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    rounded = (u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))) >> np.uint64(16)
    special = (u & np.uint64(0x7F800000)) == np.uint64(0x7F800000)
    return np.where(special, u >> np.uint64(16), rounded).astype(np.uint16)


def np_bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b).view(np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def np_dequant_bf16_native(qweight, scales_bf16_bits, qzeros, groupsize, bits, K, add_zero_bias) -> np.ndarray:
    """The bf16-native arithmetic of xbit_dequant_bf16 (NOT the reference's, which rounds through fp16,
    dq_torch_ops.cc:33-42): out = RN_bf16((w - z) * s).  (w - z) * s has at most 9 + 8 significant bits: exact in
    fp32 and in float64, so one rounding.  Returns uint16 bf16 bit patterns [K, N]."""
    N = qweight.shape[1]
    w = np_unpack_qweight(qweight, K, bits).astype(np.int64)
    z = np_unpack_qzeros(qzeros, N, bits).astype(np.int64) + int(add_zero_bias)
    grp = np.arange(K) // groupsize
    s = np_bf16_bits_to_f32(scales_bf16_bits).astype(np.float64)
    with np.errstate(over="ignore", invalid="ignore"):
        prod = ((w - z[grp]).astype(np.float64) * s[grp]).astype(np.float32)        # exact unless it overflows fp32
    return np_f32_to_bf16_bits(prod)


def np_gemv_truth(a: np.ndarray, w_f16: np.ndarray):
    y64 = a.astype(np.float64) @ w_f16.astype(np.float64)
    return y64, y64.astype(np.float16)


# --------------------------------------------------------------------------- the reference itself

class RefCpu:
    """The unmodified reference CPU simulator (cpp_simulate.cc) behind oracle/ref_cpu_driver.cc.
    Not re-entrant (the reference keeps MATRIX_K/N in globals)."""

    def __init__(self):
        build()
        if not LIB_REFCPU.exists():
            raise FileNotFoundError(f"{LIB_REFCPU} not built (needs /root/reference at build time)")
        self.lib = ctypes.CDLL(str(LIB_REFCPU))
        self.lib.refcpu_dequant.argtypes = [_u16p, _i32p, _u16p, _i32p] + [ctypes.c_int] * 4
        self.lib.refcpu_dequant.restype = ctypes.c_int
        self.lib.refcpu_gemv_as_shipped.argtypes = [_u16p, _u16p, _u16p, _i32p, _u16p, _i32p] + [ctypes.c_int] * 3
        self.lib.refcpu_gemv_as_shipped.restype = ctypes.c_int
        self.lib.refcpu_block_k.restype = ctypes.c_int

    @staticmethod
    def available() -> bool:
        try:
            build()
        except Exception:
            pass
        return LIB_REFCPU.exists()

    def dequant(self, qweight, scales, qzeros, groupsize, bits, K) -> np.ndarray:
        """cpu::DequantizeAndUnpackWeight3567_v2<ushort,bits> -- no add_zero_bias (cpp_simulate.cc:640).
        The simulator also reads scale/zero row (k0+32)/g for the last 32-row block (one past the
        end, cpp_simulate.cc:601-603): one padding row is appended here so the read stays in bounds."""
        N = qweight.shape[1]
        assert N % 2 == 0
        s = np.ascontiguousarray(scales).view(np.uint16)
        s_pad = np.concatenate([s, np.zeros((1, N), np.uint16)], axis=0)
        z_pad = np.concatenate([qzeros, np.zeros((1, qzeros.shape[1]), np.int32)], axis=0)
        # the simulator walks whole 256-"thread" blocks: rows beyond K are written when K % 32 != 0
        # is false only, but thread ids beyond the matrix still index out[]: give it slack.
        rows_pad = ceil_div(ceil_div(K, 32) * (N // 2), 256) * 256 // (N // 2) * 32 + 64
        out = np.zeros((max(rows_pad, K) + 32, N), np.uint16)
        qrows = qweight.shape[0]
        qw_pad = np.concatenate([qweight, np.zeros((bits * 40, N), np.int32)], axis=0)
        rc = self.lib.refcpu_dequant(_p(out, _u16p), _p(qw_pad, _i32p), _p(s_pad, _u16p), _p(z_pad, _i32p),
                                     K, N, bits, groupsize)
        assert rc == 0 and qrows == ceil_div(K * bits, 32)
        return out[:K].copy().view(np.float16)

    def gemv_as_shipped(self, w_unpack, a, qweight, scales, qzeros, groupsize, K):
        N = qweight.shape[1]
        out = np.zeros((1, N), np.uint16)
        rc = self.lib.refcpu_gemv_as_shipped(
            _p(out, _u16p), _p(np.ascontiguousarray(w_unpack).view(np.uint16), _u16p),
            _p(np.ascontiguousarray(a).view(np.uint16), _u16p), _p(qweight, _i32p),
            _p(np.ascontiguousarray(scales).view(np.uint16), _u16p), _p(qzeros, _i32p), K, N, groupsize)
        if rc != 0:
            raise ValueError("reference cpu_gemv only runs at K=4096, N=11008 (cpp_simulate.cc:90)")
        return out.view(np.float16)


def load_ref_gpu():
    """Import the unmodified reference extension (module name XbitOps) under the alias
    ``xbitops_ref`` from oracle/_ref/refgpu/.  Returns None when it was not built."""
    if "xbitops_ref" in sys.modules:
        return sys.modules["xbitops_ref"]
    sos = sorted(DIR_REFGPU.glob("XbitOps*.so")) if DIR_REFGPU.exists() else []
    if not sos:
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location("XbitOps", str(sos[0]))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["xbitops_ref"] = mod
    return mod
