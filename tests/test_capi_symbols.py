"""The C-ABI library loads and exports every symbol include/xbitops_b200.h declares (no compute).
Also: the product never touches the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "xbitops_b200.h")).read()
    return sorted(set(re.findall(r"XBIT_API\s+[\w\s\*]+?\b(xbit_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    names = _declared()
    for must in ("xbit_dequant_f16", "xbit_dequant_bf16", "xbit_gemv_bf16", "xbit_gemv_f16", "xbit_gemv_f16_ex", "xbit_gemv_f16_peers", "xbit_gemv_f16_host",
                 "xbit_gemv_workspace_bytes", "xbit_last_error", "xbit_version"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from xbitops_b200 import capi
    lib = capi.load()
    for name in _declared():
        assert hasattr(lib, name), f"libxbitops_b200.so does not export {name}"
        assert name in capi.SIGNATURES, f"capi.SIGNATURES lacks {name}"
    assert set(capi.SIGNATURES) == set(_declared())
    assert lib.xbit_version() >= 100


def test_validation_errors_without_gpu():
    """Argument validation happens before any CUDA call: error codes + messages, never abort."""
    from xbitops_b200 import capi
    lib = capi.load()
    one = ctypes.c_void_p(16)
    assert lib.xbit_dequant_f16(one, one, one, one, 128, 64, 9, 128, 0, None) == -1
    assert "bits" in capi.last_error()
    assert lib.xbit_dequant_f16(one, one, one, one, 128, 64, 4, 8, 0, None) == -1
    assert "groupsize" in capi.last_error()
    assert lib.xbit_dequant_f16(None, one, one, one, 128, 64, 4, 128, 0, None) == -1
    assert lib.xbit_gemv_f16(one, one, one, one, one, 0, 128, 64, 4, 128, 0, 64, None, 0, None) == -1
    assert lib.xbit_gemv_f16(one, one, one, one, one, 1, 128, 64, 4, 128, 0, 32, None, 0, None) == -1
    assert "out_row_stride" in capi.last_error()
    assert lib.xbit_gemv_f16(one, one, one, one, one, 1, 128, 64, 4, 128, 2, 64, None, 0, None) == -1
    assert lib.xbit_gemv_workspace_bytes(1, 4096, 4096, 4, 128) > 0       # optional stream-K scratch (W4 only)
    assert lib.xbit_gemv_workspace_bytes(1, 4096, 4096, 3, 128) == 0
    assert lib.xbit_gemv_workspace_bytes(1, 4096, 4096, 8, 128) > 0       # 8- and 2-bit weights run the persistent kernel too
    # bf16-native forms: same validation; the GEMV covers bits 2 / 4 / 8 at groupsize 128 only
    assert lib.xbit_dequant_bf16(one, one, one, one, 128, 64, 9, 128, 0, None) == -1
    assert lib.xbit_dequant_bf16(one, one, one, None, 128, 64, 4, 128, 0, None) == -1
    assert lib.xbit_gemv_bf16(one, one, one, one, one, 1, 128, 64, 4, 64, 0, 64, None, 0, 0, None) == -1
    assert "groupsize" in capi.last_error()
    assert lib.xbit_gemv_bf16(one, one, one, one, one, 1, 128, 64, 3, 128, 0, 64, None, 0, 0, None) == -1
    assert lib.xbit_gemv_bf16(one, one, one, one, one, 1, 128, 64, 4, 128, 0, 32, None, 0, 0, None) == -1


def test_ops_reject_cpu_tensors():
    import torch
    import xbitops_b200 as X
    qw = torch.zeros(16, 8, dtype=torch.int32)
    s = torch.zeros(1, 8, dtype=torch.float16)
    qz = torch.zeros(1, 1, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        X.dequant(qw, s, qz, 128, 4, 128, 0)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        X.gemv(torch.zeros(1, 128, dtype=torch.float16), qw, s, qz, 128, 4, 128, 0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "xbitops_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
                assert "xbit_oracle" not in text and "libxbit_refcpu" not in text, f
    hdr = open(os.path.join(ROOT, "include", "xbitops_b200.h")).read()
    assert "oracle" not in hdr


def test_auto_family_policy_without_gpu():
    """xbit_gemv_pick_family is pure host logic (148 SMs assumed without a device): the measured policy of DESIGN.md 4.2
    -- persistent kernel (5) on the Llama shapes, for 8- / 2-bit weights at groupsize 128 and for whole-tile batches of a
    short K; cluster split-K (2) for few tiles of a long K, K > 16384 and batches beyond 8 rows; the generic kernel (3)
    for odd widths, other group sizes and ragged N."""
    from xbitops_b200 import capi
    lib = capi.load()
    want = {(1, 4096, 4096, 4, 128): 5, (1, 4096, 11008, 4, 128): 5, (1, 11008, 4096, 4, 128): 5, (1, 8192, 28672, 4, 128): 5,
            (1, 8192, 1024, 4, 128): 2, (1, 28672, 8192, 4, 128): 2, (8, 4096, 11008, 4, 128): 5, (4, 8192, 8192, 4, 128): 2,
            (16, 8192, 8192, 4, 128): 2, (1, 4096, 4096, 8, 128): 5, (2, 4096, 4096, 2, 128): 5, (1, 4096, 4096, 3, 128): 3,
            (1, 4096, 4096, 4, 100): 3, (1, 4096, 4096, 8, 64): 3, (1, 4096, 4100, 4, 128): 3}
    for args, fam in want.items():
        assert lib.xbit_gemv_pick_family(*args) == fam, args
