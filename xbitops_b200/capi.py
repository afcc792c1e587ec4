"""ctypes binding of the C ABI (include/xbitops_b200.h).  No torch types cross this boundary:
plain device pointers, sizes and a stream handle -- exactly what any host language would bind.
Loading fails loudly when libxbitops_b200.so is missing; there is no CPU fallback."""
from __future__ import annotations

import ctypes
from pathlib import Path

from . import _build

XBIT_OK = 0
GEMV_AUTO, GEMV_SIMT, GEMV_MMA, GEMV_GENERIC, GEMV_PERSIST = 0, 1, 2, 3, 5
GEMV_FLAG_STATIC_WEIGHTS = 0x100
GEMV_FLAG_WAIT_PEERS = 0x200
GEMV_FLAG_A_IS_LL = 0x400

_vp, _i, _i64, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t

# symbol -> (restype, argtypes); tests/test_capi_symbols.py checks this table against the header
SIGNATURES = {
    "xbit_version": (_i, []),
    "xbit_last_error": (ctypes.c_char_p, []),
    "xbit_set_option": (_i, [ctypes.c_char_p, _i]),
    "xbit_dequant_f16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "xbit_dequant_bf16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "xbit_gemv_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i64, _vp, _sz, _i, _vp]),
    "xbit_gemv_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "xbit_gemv_f16": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i64, _vp, _sz, _vp]),
    "xbit_gemv_f16_ex": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i64, _vp, _sz, _i, _vp]),
    "xbit_gemv_f16_multi": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _sz, _i, _vp]),
    "xbit_gemv_pick_family": (_i, [_i, _i, _i, _i, _i]),
    "xbit_gemv_f16_peers": (_i, [_vp, _vp, _vp, _vp, ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i, _i64, _i64,
                                 _vp, _sz, _vp]),
    "xbit_gemv_f16_peers_ex": (_i, [_vp, _vp, _vp, _vp, ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i, _i64, _i64,
                                    _vp, _sz, _i, _vp]),
    "xbit_gemv_f16_peers_signal": (_i, [_vp, _vp, _vp, _vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _vp, _i, _i, _i, _i, _i,
                                        _i, _i, _i, _i64, _i64, _i, _vp]),
    "xbit_peers_wait": (_i, [_vp, _i, _i, _vp, _vp]),
    "xbit_gemv_f16_peers_ll": (_i, [_vp, _vp, _vp, _vp, ctypes.POINTER(_vp), _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i64, _i64,
                                    _i, _vp]),
    "xbit_ll_unpack_f16": (_i, [_vp, _vp, _i64, _vp, _i, _vp, _vp]),
    "xbit_gemv_f16_host": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
}



class GemvProblem(ctypes.Structure):
    """struct xbit_gemv_problem (include/xbitops_b200.h)"""
    _fields_ = [("qweight", _vp), ("scales_f16", _vp), ("qzeros", _vp), ("out_f16", _vp), ("N", _i), ("out_row_stride", _i64)]


_lib = None


def lib_path() -> Path:
    # XBIT_DEVTOOLS_LIB=1 (tools/*.py only): the -DXBIT_DEVTOOLS build with phase stamps and skip-math knobs
    import os
    if os.environ.get("XBIT_B200_LIB"):            # tools/*.py only: an experimental build of the same ABI
        return Path(os.environ["XBIT_B200_LIB"])
    return _build.LIB_DEV if os.environ.get("XBIT_DEVTOOLS_LIB") else _build.LIB


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load libxbitops_b200.so (building it with nvcc when absent or stale and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing and path in (_build.LIB, _build.LIB_DEV):
        try:
            _build.build_lib(dev=(path == _build.LIB_DEV))
        except Exception:
            if not path.exists():
                raise
    if not path.exists():
        raise ImportError(f"{path} is missing: build it with `python -m xbitops_b200._build` "
                          "(xbitops_b200 has no CPU fallback)")
    lib = ctypes.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


UNSET = -2**31          # xbit_set_option(name, UNSET): back to the built-in policy


def set_option(name: str, value: int = UNSET) -> None:
    check(load().xbit_set_option(name.encode(), int(value)))


def last_error() -> str:
    return load().xbit_last_error().decode()


def check(rc: int) -> None:
    if rc != XBIT_OK:
        raise RuntimeError(f"xbitops_b200: {last_error()} (code {rc})")
