"""The compiled drop-in module `XbitOps` (csrc/dq_torch_ops.cc, same name / positional surface as the
reference extension, /root/reference/src/dq_torch_ops.cc:80-85) against the ctypes mirror: both go
through the same C ABI, so results must be bit-identical."""
import importlib.util

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from xbitops_b200 import _build, synth  # noqa: E402
import xbitops_b200 as X  # noqa: E402


@pytest.fixture(scope="module")
def ext():
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test without a CUDA device")
    path = _build.torch_ext_path()
    if path is None:
        path = _build.build_torch_ext()
    spec = importlib.util.spec_from_file_location("XbitOps", str(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _inputs(K, N, bits, g, M):
    qw, s, qz, a = synth.make_inputs(K, N, bits, g, M=M, seed=K + bits)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()  # noqa: E731
    return d(qw), d(s.view(np.int16)).view(torch.float16), d(qz), d(a.view(np.int16)).view(torch.float16)


def test_same_surface_as_reference(ext):
    assert hasattr(ext, "dequant") and hasattr(ext, "gemv")
    assert "add_zero_bias" in ext.dequant.__doc__ or "int" in ext.dequant.__doc__


@pytest.mark.parametrize("bits", (2, 3, 4, 8))
def test_pybind_equals_ctypes(bits, ext):
    K, N, g = 1024, 512, 128
    tq, ts, tz, ta = _inputs(K, N, bits, g, 3)
    assert torch.equal(ext.dequant(tq, ts, tz, g, bits, K, 1), X.dequant(tq, ts, tz, g, bits, K, 1))
    assert torch.equal(ext.gemv(ta, tq, ts, tz, g, bits, K, 1), X.gemv(ta, tq, ts, tz, g, bits, K, 1))
    y3 = ext.gemv(ta.view(3, 1, K), tq, ts, tz, g, bits, K, 0)
    assert tuple(y3.shape) == (3, 1, N)
    yb = ext.gemv(ta, tq, ts.to(torch.bfloat16), tz, g, bits, K, 0)
    assert yb.dtype == torch.bfloat16


def test_pybind_errors_and_stream(ext):
    K, N, g = 512, 256, 128
    tq, ts, tz, ta = _inputs(K, N, 4, g, 1)
    with pytest.raises(RuntimeError):
        ext.dequant(tq, ts, tz, 8, 4, K, 0)
    with pytest.raises(RuntimeError):
        ext.gemv(ta.float(), tq, ts, tz, g, 4, K, 0)
    want = ext.gemv(ta, tq, ts, tz, g, 4, K, 0)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        got = ext.gemv(ta, tq, ts, tz, g, 4, K, 0)
    side.synchronize()
    assert torch.equal(got, want)
