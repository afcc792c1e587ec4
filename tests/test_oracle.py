"""The oracle against (a) its independent numpy restatement, (b) the golden vectors produced by
executing the reference's CPU simulator and GPU kernels, (c) the reference CPU simulator itself when
/root/reference is present (build container only).  No GPU."""
import os
import re

import numpy as np
import pytest

from oracle import oracle as O
from xbitops_b200 import synth


def _cases(npz, suffix):
    pat = re.compile(r"b(\d)_g(\d+)_K(\d+)_N(\d+)(?:_z(\d))?_" + suffix + "$")
    for k in npz.files:
        m = pat.match(k)
        if m:
            yield k[: -len(suffix) - 1], tuple(int(x) if x is not None else None for x in m.groups())


def test_c_oracle_equals_numpy_restatement(c_oracle):
    for bits in range(2, 9):
        for (K, N, g) in ((96, 16, 32), (200, 24, 64), (130, 10, 48), (64, 8, 16)):
            for bias in (0, 1):
                for mode in ("gptq", "bits"):
                    qw, s, qz, a = synth.make_inputs(K, N, bits, g, M=3, seed=bits + K, scale_mode=mode)
                    assert (c_oracle.unpack_qweight(qw, K, bits) == O.np_unpack_qweight(qw, K, bits)).all()
                    assert (c_oracle.unpack_qzeros(qz, N, bits) == O.np_unpack_qzeros(qz, N, bits)).all()
                    dc = c_oracle.dequant(qw, s, qz, g, bits, K, bias)
                    dn = O.np_dequant(qw, s, qz, g, bits, K, bias)
                    assert (dc.view(np.uint16) == dn.view(np.uint16)).all(), (bits, K, N, g, bias, mode)
    qw, s, qz, a = synth.make_inputs(256, 32, 4, 128, M=3)
    w = c_oracle.dequant(qw, s, qz, 128, 4, 256, 1)
    y64, y16 = c_oracle.gemv_from_dq(a, w)
    n64, n16 = O.np_gemv_truth(a, w)
    assert np.allclose(y64, n64, rtol=1e-12, atol=1e-12)
    assert (np.abs(y16.astype(np.float64) - n16.astype(np.float64)) <= np.spacing(np.abs(n16)).astype(np.float64)).all()


def test_f16_conversion_exhaustive(c_oracle):
    # every fp16 bit pattern round-trips; midpoints round to even
    for h in range(0, 0x7C00):
        d = c_oracle.lib.xo_f16_to_f64(h)
        assert c_oracle.lib.xo_f64_to_f16(d) == h
        assert c_oracle.lib.xo_f64_to_f16(-d) == (h | 0x8000)
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(20000) * 10.0 ** rng.integers(-9, 5, 20000), [65504.0, 65520.0, 65519.9, 2.0 ** -25, 2.0 ** -24 * 1.5]])
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).view(np.uint16)
    got = np.array([c_oracle.lib.xo_f64_to_f16(float(v)) for v in x], np.uint16)
    assert (got == want).all()


def test_pack_unpack_roundtrip(c_oracle):
    rng = np.random.default_rng(1)
    for bits in range(2, 9):
        for (K, N) in ((7, 3), (96, 16), (33, 9)):
            w = rng.integers(0, 1 << bits, (K, N), dtype=np.uint8)
            q = synth.pack_qweight(w, bits)
            assert q.shape == ((K * bits + 31) // 32, N)
            assert (q == c_oracle.pack_qweight(w, bits)).all()
            assert (c_oracle.unpack_qweight(q, K, bits) == w).all()
            z = rng.integers(0, 1 << bits, (4, N), dtype=np.uint8)
            qz = synth.pack_qzeros(z, bits)
            assert (qz == c_oracle.pack_qzeros(z, bits)).all()
            assert (c_oracle.unpack_qzeros(qz, N, bits) == z).all()


def test_oracle_vs_golden_refcpu(c_oracle, golden_cpu):
    """Golden = outputs of the reference's own CPU simulator. Integers must be exact; the fp16
    outputs must equal the oracle formula evaluated with the simulator's tie-away rounding
    (cpp_simulate.cc:47) bit for bit -- the only difference to the GPU's RN-even arithmetic."""
    n = 0
    for tag, (bits, g, K, N, _) in _cases(golden_cpu, "ref_cpu"):
        qw, s, qz = golden_cpu[tag + "_qweight"], golden_cpu[tag + "_scales"].view(np.float16), golden_cpu[tag + "_qzeros"]
        assert (c_oracle.unpack_qweight(qw, K, bits) == golden_cpu[tag + "_ref_ints"]).all(), tag
        assert (c_oracle.unpack_qzeros(qz, N, bits) == golden_cpu[tag + "_ref_zints"]).all(), tag
        sim = O.np_dequant(qw, s, qz, g, bits, K, 0, rounding="sim")
        assert (sim.view(np.uint16) == golden_cpu[tag + "_ref_cpu"]).all(), tag
        # and RN-even differs from the simulator by at most one rounding of sz and one of out
        rne = c_oracle.dequant(qw, s, qz, g, bits, K, 0).astype(np.float64)
        ref = golden_cpu[tag + "_ref_cpu"].view(np.float16).astype(np.float64)
        z = c_oracle.unpack_qzeros(qz, N, bits).astype(np.float64)
        sz = np.abs(z * s.astype(np.float64))[np.arange(K) // g]
        tol = np.spacing(sz.astype(np.float16)).astype(np.float64) + np.spacing(np.maximum(np.abs(rne), np.abs(ref)).astype(np.float16)).astype(np.float64)
        assert (np.abs(rne - ref) <= tol).all(), tag
        n += 1
    assert n == 21


def test_oracle_vs_golden_refgpu(c_oracle, golden_gpu):
    """Golden = outputs of the reference's GPU kernels on a B200: the oracle must match bit for bit."""
    n = 0
    for tag, (bits, g, K, N, bias) in _cases(golden_gpu, "ref_gpu"):
        qw, s, qz = golden_gpu[tag + "_qweight"], golden_gpu[tag + "_scales"].view(np.float16), golden_gpu[tag + "_qzeros"]
        want = golden_gpu[tag + "_ref_gpu"]
        got = c_oracle.dequant(qw, s, qz, g, bits, K, bias).view(np.uint16)[: want.shape[0]]
        assert (got == want).all(), tag
        n += 1
    assert n > 0


@pytest.mark.skipif(not os.path.exists("/root/reference/src/cpp_simulate.cc"), reason="reference tree not present")
def test_oracle_vs_live_reference_cpu(c_oracle):
    rc = O.RefCpu()
    for bits in range(2, 9):
        for (K, N, g) in ((512, 64, 128), (416, 32, 32)):
            qw, s, qz, _ = synth.make_inputs(K, N, bits, g, seed=77 + bits)
            ints = rc.dequant(qw, np.ones_like(s), np.zeros_like(qz), g, bits, K)
            assert (ints.astype(np.int32) == c_oracle.unpack_qweight(qw, K, bits)).all()
            live = rc.dequant(qw, s, qz, g, bits, K)
            sim = O.np_dequant(qw, s, qz, g, bits, K, 0, rounding="sim")
            assert (live.view(np.uint16) == sim.view(np.uint16)).all()


def test_reference_arith_gemv_is_close_to_truth(c_oracle):
    """The reference's shipped fp16-chain arithmetic sits within 1e-2 of the truth (context for the
    tolerance the GPU tests use)."""
    K, N = 1024, 64
    qw, s, qz, a = synth.make_inputs(K, N, 4, 128, seed=5)
    y64, _ = c_oracle.gemv(a, qw, s, qz, 128, 4, K, 0)
    yr = c_oracle.gemv_w4_ref_arith(a, qw, s, qz, 128, K, 0).astype(np.float64)
    assert np.abs(yr - y64).max() / np.abs(y64).max() < 5e-3
