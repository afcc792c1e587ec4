"""The compiled drop-in module `XbitOps` (csrc/dq_torch_ops.cc, same name / positional surface as the
reference extension, /root/reference/src/dq_torch_ops.cc:80-85) against the ctypes mirror: both go
through the same C ABI, so results must be bit-identical.

Runs in a fresh interpreter: the module is a drop-in REPLACEMENT for the reference extension of the
same name, and the parity tests of this suite load that reference extension into the pytest process
(two pybind11 modules called `XbitOps` in one process is not a supported configuration)."""
import os
import subprocess
import sys
import textwrap

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BODY = r'''
import importlib.util, sys
import numpy as np, torch
sys.path.insert(0, ROOT)
from xbitops_b200 import _build, synth
import xbitops_b200 as X

path = _build.torch_ext_path() or _build.build_torch_ext()
spec = importlib.util.spec_from_file_location("XbitOps", str(path))
ext = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ext)
assert hasattr(ext, "dequant") and hasattr(ext, "gemv")

def inputs(K, N, bits, g, M):
    qw, s, qz, a = synth.make_inputs(K, N, bits, g, M=M, seed=K + bits)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    return d(qw), d(s.view(np.int16)).view(torch.float16), d(qz), d(a.view(np.int16)).view(torch.float16)

for bits in (2, 3, 4, 8):
    K, N, g = 1024, 512, 128
    tq, ts, tz, ta = inputs(K, N, bits, g, 3)
    assert torch.equal(ext.dequant(tq, ts, tz, g, bits, K, 1), X.dequant(tq, ts, tz, g, bits, K, 1)), bits
    assert torch.equal(ext.gemv(ta, tq, ts, tz, g, bits, K, 1), X.gemv(ta, tq, ts, tz, g, bits, K, 1)), bits
    y3 = ext.gemv(ta.view(3, 1, K), tq, ts, tz, g, bits, K, 0)
    assert tuple(y3.shape) == (3, 1, N)
    yb = ext.gemv(ta, tq, ts.to(torch.bfloat16), tz, g, bits, K, 0)
    assert yb.dtype == torch.bfloat16
    wb = ext.dequant(tq, ts.to(torch.bfloat16), tz, g, bits, K, 0)
    assert wb.dtype == torch.bfloat16 and tuple(wb.shape) == (K, N)

# errors are RuntimeError, never process death; work follows the current stream
K, N, g = 512, 256, 128
tq, ts, tz, ta = inputs(K, N, 4, g, 1)
for bad in (lambda: ext.dequant(tq, ts, tz, 8, 4, K, 0), lambda: ext.gemv(ta.float(), tq, ts, tz, g, 4, K, 0),
            lambda: ext.dequant(tq.cpu(), ts, tz, g, 4, K, 0), lambda: ext.dequant(tq, ts, tz, g, 4, K + 8, 0)):
    try:
        bad()
    except RuntimeError:
        pass
    else:
        raise AssertionError("expected RuntimeError")
want = ext.gemv(ta, tq, ts, tz, g, 4, K, 0)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    got = ext.gemv(ta, tq, ts, tz, g, 4, K, 0)
side.synchronize()
assert torch.equal(got, want)
print("shim ok")
'''


def test_pybind_shim_in_fresh_process():
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test without a CUDA device")
    code = "ROOT = %r\n" % ROOT + textwrap.dedent(BODY)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "shim ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
