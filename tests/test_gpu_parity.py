"""Parity tests proper: the CUDA path (through the C ABI) against the oracle on the same seeded
inputs.  DQ: bit-exact.  GEMV: the north star's fp16 tolerance (max relative error <= 1e-2) against
the fp64-accumulated truth over the bit-exact dequantised weights, written as
    max|y - y_ref| / max|y_ref| <= 1e-2                      (normalised max error)
    |y - y_ref| <= 1e-2 * |y_ref| + 2e-3 * max|y_ref|         (element-wise: 1 % relative plus an absolute
        floor so that near-zero outputs do not blow the relative error up)
    The truth is defined over the fp16-ROUNDED dequantised weights (the reference semantics: every
    weight is rounded to fp16 before the dot product), which alone puts any exact-arithmetic kernel
    ~3e-4 (normalised) away from it; the reference's own shipped kernel sits at ~5e-4
    (tests/test_oracle.py::test_reference_arith_gemv_is_close_to_truth), ours at 4e-4 .. 9e-4.
Run on the B200 box: pytest -m gpu."""
import ctypes

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from xbitops_b200 import capi, synth  # noqa: E402
import xbitops_b200 as X  # noqa: E402

GEMV_TOL = 1e-2


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test without a CUDA device (no CPU fallback exists)")
    return torch.device("cuda:0")


def t16(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).to(dev).view(torch.float16)


def ti(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def assert_gemv_close(y, y64, what="", floor=0.1):
    y = np.asarray(y, np.float64)
    mx = np.abs(y64).max()
    err = np.abs(y - y64)
    assert err.max() / mx <= GEMV_TOL, f"{what}: normalised max error {err.max() / mx:.3e}"
    assert (err <= GEMV_TOL * np.abs(y64) + 2e-3 * mx).all(), f"{what}: element-wise bound violated"


def floor_of(family):
    return 0.1


# ------------------------------------------------------------------ dequant

@pytest.mark.parametrize("bits", range(2, 9))
@pytest.mark.parametrize("g", (32, 64, 128))
def test_dequant_bit_exact_vs_oracle(bits, g, dev, c_oracle):
    for (K, N) in ((256, 64), (416, 136), (1024, 512), (2048 + 32, 1000)):
        for bias in (0, 1):
            for mode in ("gptq", "bits"):
                qw, s, qz, _ = synth.make_inputs(K, N, bits, g, seed=bits * 100 + g, scale_mode=mode)
                want = c_oracle.dequant(qw, s, qz, g, bits, K, bias)
                got = X.dequant(ti(qw, dev), t16(s, dev), ti(qz, dev), g, bits, K, bias).cpu().numpy()
                assert got.shape == (K, N) and got.dtype == np.float16
                assert (got.view(np.uint16) == want.view(np.uint16)).all(), (bits, g, K, N, bias, mode)


@pytest.mark.parametrize("bits", range(2, 9))
def test_dequant_ragged_and_fallback_shapes(bits, dev, c_oracle):
    # ragged K (not a multiple of 32 or of the group), N not a multiple of 8, groupsize not a multiple of 32
    for (K, N, g) in ((100, 12, 48), (333, 34, 16), (7, 2, 16), (33, 8, 32), (95, 16, 64), (4097, 24, 128)):
        qw, s, qz, _ = synth.make_inputs(K, N, bits, g, seed=bits)
        want = c_oracle.dequant(qw, s, qz, g, bits, K, 1)
        got = X.dequant(ti(qw, dev), t16(s, dev), ti(qz, dev), g, bits, K, 1).cpu().numpy()
        assert (got.view(np.uint16) == want.view(np.uint16)).all(), (bits, K, N, g)


@pytest.mark.parametrize("bits", range(2, 9))
def test_unpacked_integers_through_the_op(bits, dev, c_oracle):
    """scales == 1, zeros == 0 -> the op returns half(w): integer-exact unpack check;
    qweight == 0, scales == -1 -> the unpacked zero points."""
    K, N, g = 512, 256, 128
    qw, s, qz, _ = synth.make_inputs(K, N, bits, g, seed=9)
    ones = np.ones_like(s)
    got = X.dequant(ti(qw, dev), t16(ones, dev), ti(np.zeros_like(qz), dev), g, bits, K, 0).cpu().numpy()
    assert (got.astype(np.int32) == c_oracle.unpack_qweight(qw, K, bits)).all()
    gz = X.dequant(ti(np.zeros_like(qw), dev), t16(-ones, dev), ti(qz, dev), g, bits, K, 1).cpu().numpy()
    assert (gz[::g].astype(np.int32) == c_oracle.unpack_qzeros(qz, N, bits).astype(np.int32) + 1).all()


def test_dequant_golden_refcpu_integers(dev, golden_cpu):
    import re
    for k in golden_cpu.files:
        m = re.match(r"(b(\d)_g(\d+)_K(\d+)_N(\d+))_ref_ints$", k)
        if not m:
            continue
        tag, bits, g, K, N = m.group(1), *map(int, m.groups()[1:])
        qw, s, qz = golden_cpu[tag + "_qweight"], golden_cpu[tag + "_scales"], golden_cpu[tag + "_qzeros"]
        one = np.ones(s.shape, np.float16)
        got = X.dequant(ti(qw, dev), t16(one, dev), ti(np.zeros_like(qz), dev), g, bits, K, 0).cpu().numpy()
        assert (got.astype(np.int32) == golden_cpu[k]).all(), tag


def test_dequant_golden_refgpu(dev, golden_gpu):
    import re
    n = 0
    for k in golden_gpu.files:
        m = re.match(r"(b(\d)_g(\d+)_K(\d+)_N(\d+)_z(\d))_ref_gpu$", k)
        if not m:
            continue
        tag, bits, g, K, N, bias = m.group(1), *map(int, m.groups()[1:])
        qw, s, qz = golden_gpu[tag + "_qweight"], golden_gpu[tag + "_scales"].view(np.float16), golden_gpu[tag + "_qzeros"]
        want = golden_gpu[k]
        got = X.dequant(ti(qw, dev), t16(s, dev), ti(qz, dev), g, bits, K, bias).cpu().numpy().view(np.uint16)
        assert (got[: want.shape[0]] == want).all(), tag
        n += 1
    assert n > 0


def test_dequant_bf16_scales_and_full_size_properties(dev):
    """BASELINE config 3 size (4096 x 11008): size-independent properties instead of the oracle:
    (a) integer trick, checked against a torch restatement of the bit stream on the GPU;
    (b) bf16 scales give exactly fp16-path output cast to bf16 (dq_torch_ops.cc:33-42);
    (c) linearity in the scale for power-of-two factors (exact in fp16 away from the range ends)."""
    K, N, bits, g = 4096, 11008, 4, 128
    gen = torch.Generator(device=dev).manual_seed(0)
    qw = torch.randint(-2**31, 2**31 - 1, (K * bits // 32, N), dtype=torch.int32, device=dev, generator=gen)
    qz = torch.randint(-2**31, 2**31 - 1, (K // g, N * bits // 32), dtype=torch.int32, device=dev, generator=gen)
    s = (torch.rand((K // g, N), device=dev, generator=gen) * 0.018 + 0.002).to(torch.float16)
    ints = X.dequant(qw, torch.ones_like(s), torch.zeros_like(qz), g, bits, K, 0)
    shifts = torch.arange(0, 32, bits, device=dev, dtype=torch.int32).view(1, -1, 1)
    want = ((qw.unsqueeze(1) >> shifts) & ((1 << bits) - 1)).reshape(K, N).to(torch.float16)
    assert torch.equal(ints, want)
    w = X.dequant(qw, s, qz, g, bits, K, 1)
    wb = X.dequant(qw, s.to(torch.bfloat16), qz, g, bits, K, 1)
    assert wb.dtype == torch.bfloat16
    assert torch.equal(wb, X.dequant(qw, s.to(torch.bfloat16).to(torch.float16), qz, g, bits, K, 1).to(torch.bfloat16))
    w4 = X.dequant(qw, s * 4, qz, g, bits, K, 1)
    assert torch.equal(w4, w * 4)


# ------------------------------------------------------------------ gemv

FAMILIES = [(capi.GEMV_SIMT, (1, 3)), (capi.GEMV_MMA, (1, 2, 7, 8, 9, 16)), (capi.GEMV_GENERIC, (1, 5)),
            (capi.GEMV_PERSIST, (1, 2, 3, 5, 8))]


@pytest.mark.parametrize("family,Ms", FAMILIES)
def test_gemv_w4_families_vs_truth(family, Ms, dev, c_oracle):
    for (K, N, g) in ((4096, 4096, 128), (1024, 256, 32), (11008, 512, 128), (2048, 8192 + 64, 64), (4096, 96, 128), (128, 32, 32), (384, 160, 64)):
        for bias in (0, 1):
            qw, s, qz, a = synth.make_inputs(K, N, 4, g, M=max(Ms), seed=K + N)
            w = c_oracle.dequant(qw, s, qz, g, 4, K, bias)
            tq, ts, tz = ti(qw, dev), t16(s, dev), ti(qz, dev)
            for M in Ms:
                if family == capi.GEMV_PERSIST and M * K > 8 * 8192:
                    continue                     # M * K too large to stage in one SM's shared memory: AUTO uses the cluster kernel
                y64 = a[:M].astype(np.float64) @ w.astype(np.float64)
                got = X.gemv(t16(a[:M], dev), tq, ts, tz, g, 4, K, bias, family=family).cpu().numpy()
                assert got.shape == (M, N)
                assert_gemv_close(got, y64, f"family={family} M={M} K={K} N={N} g={g} bias={bias}", floor_of(family))


@pytest.mark.parametrize("bits", (2, 3, 5, 6, 7, 8))
def test_gemv_other_bit_widths(bits, dev, c_oracle):
    """The reference aborts for bits != 4 (gemv_w4a16_pt.cu:152-155); the intended A16Wx semantics are
    y = a @ DQ for every width."""
    for (M, K, N, g) in ((1, 1024, 256, 128), (2, 777, 100, 48), (17, 512, 64, 32), (1, 4096, 512, 64)):
        qw, s, qz, a = synth.make_inputs(K, N, bits, g, M=M, seed=bits)
        y64, _ = c_oracle.gemv(a, qw, s, qz, g, bits, K, 1)
        got = X.gemv(t16(a, dev), ti(qw, dev), t16(s, dev), ti(qz, dev), g, bits, K, 1).cpu().numpy()
        assert_gemv_close(got, y64, f"bits={bits} M={M} K={K} N={N} g={g}")


def test_gemv_streamk_schedule(dev, c_oracle, monkeypatch):
    """The opt-in persistent stream-K schedule (XBIT_GEMV_STREAMK=1 + workspace): same results, and the
    workspace is left zeroed (flags cleared) so that consecutive calls can share it."""
    from xbitops_b200 import ops
    capi.set_option("XBIT_GEMV_STREAMK", 1)
    # (K, N, M, g, bias): whole tiles per CTA and tiles shared by several CTAs, a half-empty last stage
    # (K = 4224 = 33 blocks), a partial last tile (N = 4128), fewer units than SMs (N = 96), every
    # groupsize of the fast path, both tensor-core tile heights (M <= 8, M <= 16)
    cases = ((4096, 4096, 1, 128, 1), (11008, 4096, 1, 128, 1), (4096, 11008, 2, 128, 0), (8192, 1024, 5, 128, 1),
             (4224, 4128, 1, 128, 1), (1024, 96, 3, 64, 0), (2048, 2048, 8, 32, 1), (4096, 512, 13, 128, 1),
             (8192, 8192, 16, 64, 0))
    for (K, N, M, g, bias) in cases:
        qw, s, qz, a = synth.make_inputs(K, N, 4, g, M=M, seed=K + N + M)
        w = c_oracle.dequant(qw, s, qz, g, 4, K, bias)
        y64 = a.astype(np.float64) @ w.astype(np.float64)
        tq, ts, tz, ta = ti(qw, dev), t16(s, dev), ti(qz, dev), t16(a, dev)
        y1 = X.gemv(ta, tq, ts, tz, g, 4, K, bias, family=capi.GEMV_MMA)
        y2 = X.gemv(ta, tq, ts, tz, g, 4, K, bias, family=capi.GEMV_MMA)
        assert torch.equal(y1, y2)                           # deterministic, workspace reusable
        assert_gemv_close(y1.cpu().numpy(), y64, f"stream-K {K}x{N} M={M} g={g}")
        capi.set_option("XBIT_GEMV_STREAMK", 0)         # the cluster split-K kernel on the same inputs
        y3 = X.gemv(ta, tq, ts, tz, g, 4, K, bias, family=capi.GEMV_MMA)
        capi.set_option("XBIT_GEMV_STREAMK", 1)
        assert_gemv_close(y3.cpu().numpy(), y64, f"cluster {K}x{N} M={M} g={g}")
        assert float((y1.double() - y3.double()).abs().max()) <= 2e-3 * float(np.abs(y64).max())
    capi.set_option("XBIT_GEMV_STREAMK")
    torch.cuda.synchronize()
    ws = ops.gemv_workspace(dev)
    assert int(ws[: 148 * 4].view(torch.int32).abs().sum()) == 0      # ready flags cleared


def test_gemv_persistent_schedule(dev, c_oracle, monkeypatch):
    """The default W4 schedule (one persistent CTA per SM, per-warp rings): block-granular CTA boundaries with the
    cross-CTA fix-up through the workspace (XBIT_W4P_FINE=1) and tile-aligned boundaries (=0) give results within
    fp32 summation-order noise of each other, are deterministic, and leave the workspace zeroed."""
    from xbitops_b200 import ops
    # whole tiles per CTA, tiles shared by 2..3 CTAs (long K), fewer blocks than warps (128 x 32), more tiles than
    # SMs, every groupsize of the fast path, odd block counts (K = 4224: 33 blocks)
    cases = ((4096, 4096, 1, 128, 1), (11008, 4096, 1, 128, 1), (4096, 11008, 2, 128, 0), (28672, 1024, 1, 128, 1),
             (4224, 4128, 1, 128, 1), (1024, 96, 3, 64, 0), (2048, 2048, 8, 32, 1), (128, 32, 1, 32, 0),
             (256, 64, 2, 128, 1), (8192, 8192, 4, 64, 0), (384, 4736, 1, 128, 1),
             # tall N-split shards: a CTA's range is shorter than a tile, so only a window of the activations is staged
             (8192, 1024, 2, 128, 1), (8192, 1024, 3, 64, 0), (16384, 512, 1, 128, 1))
    for (K, N, M, g, bias) in cases:
        qw, s, qz, a = synth.make_inputs(K, N, 4, g, M=M, seed=K + N + M)
        w = c_oracle.dequant(qw, s, qz, g, 4, K, bias)
        y64 = a.astype(np.float64) @ w.astype(np.float64)
        tq, ts, tz, ta = ti(qw, dev), t16(s, dev), ti(qz, dev), t16(a, dev)
        ys = {}
        for fine in ("1", "0"):
            capi.set_option("XBIT_W4P_FINE", int(fine))
            y1 = X.gemv(ta, tq, ts, tz, g, 4, K, bias, family=capi.GEMV_PERSIST)
            y2 = X.gemv(ta, tq, ts, tz, g, 4, K, bias, family=capi.GEMV_PERSIST)
            assert torch.equal(y1, y2), f"persist fine={fine} {K}x{N} M={M}: not deterministic"
            assert_gemv_close(y1.cpu().numpy(), y64, f"persist fine={fine} {K}x{N} M={M} g={g}")
            ys[fine] = y1
        assert float((ys["1"].double() - ys["0"].double()).abs().max()) <= 2e-3 * float(np.abs(y64).max())
    capi.set_option("XBIT_W4P_FINE")
    torch.cuda.synchronize()
    ws = ops.gemv_workspace(dev)
    lib = capi.load()
    lo = lib.xbit_gemv_workspace_bytes(16, 0, 0, 4, 128) - 148 * 8 * 32 * 8       # the persistent kernel's region is the tail
    assert int(ws[lo:].view(torch.int32).abs().sum()) == 0      # partial slots and flags cleared


def test_gemv_multi_projection_is_bit_identical(dev):
    """xbit_gemv_f16_multi: Q/K/V and gate + up in one launch give, bit for bit, what separate gemv
    calls give (the reference issues one op per projection, dq_torch_ops.cc:46-78); unequal widths, M = 2, a group size
    of the fp16 block math, and a combination that cannot be fused (bits 3) fall back to separate launches."""
    gen = torch.Generator(device=dev).manual_seed(99)

    def rand_proj(K, N, bits, g):
        qw = torch.randint(-2**31, 2**31 - 1, ((K * bits + 31) // 32, N), dtype=torch.int32, device=dev, generator=gen)
        qz = torch.randint(-2**31, 2**31 - 1, (K // g, (N * bits + 31) // 32), dtype=torch.int32, device=dev, generator=gen)
        s = (torch.rand((K // g, N), device=dev, generator=gen) * 0.018 + 0.002).to(torch.float16)
        return qw, s, qz

    cases = ((4096, (4096, 4096, 4096), 1, 4, 128),        # Q, K, V of Llama-2-7B
             (4096, (11008, 11008), 1, 4, 128),            # gate, up
             (8192, (8192, 1024, 1024), 2, 4, 128),        # grouped-query attention: unequal widths
             (2048, (2048, 512, 2048, 64), 3, 4, 64),      # four matrices, fp16 block math
             (1024, (256, 512), 1, 3, 128))                # not fusable: the generic kernel, one launch per matrix
    for (K, Ns, M, bits, g) in cases:
        a = torch.randn((M, K), device=dev, generator=gen).to(torch.float16)
        projs = [rand_proj(K, N, bits, g) for N in Ns]
        fused = X.gemv_multi(a, projs, g, bits, K, 1)
        for (q, s, z), y in zip(projs, fused):
            sep = X.gemv(a, q, s, z, g, bits, K, 1)
            assert torch.equal(y, sep), f"multi K={K} Ns={Ns} M={M} g={g}: differs from the separate call"
            truth = a.double() @ X.dequant(q, s, z, g, bits, K, 1).double()
            assert_gemv_close(y.cpu().numpy(), truth.cpu().numpy(), f"multi K={K} N={q.shape[1]}")


def test_gemv_shapes_dtypes_and_large_m(dev, c_oracle):
    K, N, g = 1024, 512, 128
    qw, s, qz, a = synth.make_inputs(K, N, 4, g, M=40, seed=3)
    w = c_oracle.dequant(qw, s, qz, g, 4, K, 0)
    y64 = a.astype(np.float64) @ w.astype(np.float64)
    tq, ts, tz = ti(qw, dev), t16(s, dev), ti(qz, dev)
    a3 = t16(a, dev).view(5, 8, K)                       # [B, S, K] -> [B, S, N]   (dq_torch_ops.cc:59-64)
    y = X.gemv(a3, tq, ts, tz, g, 4, K, 0)
    assert tuple(y.shape) == (5, 8, N) and y.dtype == torch.float16
    assert_gemv_close(y.view(40, N).cpu().numpy(), y64, "M=40")
    yb = X.gemv(t16(a[:2], dev), tq, ts.to(torch.bfloat16), tz, g, 4, K, 0)
    assert yb.dtype == torch.bfloat16 and tuple(yb.shape) == (2, N)


def test_gemv_golden_refgpu(dev, golden_gpu, c_oracle):
    """Outputs of the reference's own GPU gemv: ours must agree with them at least as well as both
    agree with the truth."""
    import re
    n = 0
    for k in golden_gpu.files:
        m = re.match(r"(gemv_K(\d+)_N(\d+)_z(\d))_ref_gpu$", k)
        if not m:
            continue
        tag, K, N, bias = m.group(1), *map(int, m.groups()[1:])
        qw, s, qz = golden_gpu[tag + "_qweight"], golden_gpu[tag + "_scales"].view(np.float16), golden_gpu[tag + "_qzeros"]
        a = golden_gpu[tag + "_a"].view(np.float16)
        ref = golden_gpu[k].view(np.float16).astype(np.float64)
        y64, _ = c_oracle.gemv(a, qw, s, qz, 128, 4, K, bias)
        for fam in (capi.GEMV_SIMT, capi.GEMV_MMA):
            got = X.gemv(t16(a, dev), ti(qw, dev), t16(s, dev), ti(qz, dev), 128, 4, K, bias, family=fam).cpu().numpy()
            assert_gemv_close(got, y64, tag)
            assert np.abs(got.astype(np.float64) - ref).max() / np.abs(y64).max() <= GEMV_TOL
        n += 1
    assert n > 0


@pytest.mark.parametrize("K,N", [(4096, 4096), (4096, 11008), (11008, 4096), (8192, 8192)])
def test_gemv_full_size_properties(K, N, dev):
    """BASELINE sizes, size-independent properties (no CPU oracle at this size):
    (a) gemv == activations @ dequant (our own bit-exact DQ, fp32 matmul on the GPU) within tolerance;
    (b) column sharding: gemv on a column slice agrees with the slice of the truth (the K-split chosen
        by the planner depends on N, so the fp32 summation order may differ from the unsharded call);
    (c) row m of an M-row call == the M=1 call on that row (same family), bit for bit;
    (d) linearity: gemv(2a) == 2*gemv(a) (power-of-two scaling is exact in fp16/fp32 except for subnormals)."""
    g, bits = 128, 4
    gen = torch.Generator(device=dev).manual_seed(K + N)
    qw = torch.randint(-2**31, 2**31 - 1, (K // 8, N), dtype=torch.int32, device=dev, generator=gen)
    qz = torch.randint(-2**31, 2**31 - 1, (K // g, N // 8), dtype=torch.int32, device=dev, generator=gen)
    s = (torch.rand((K // g, N), device=dev, generator=gen) * 0.018 + 0.002).to(torch.float16)
    a = torch.randn((4, K), device=dev, generator=gen).to(torch.float16)
    w = X.dequant(qw, s, qz, g, bits, K, 1)
    truth = (a.double() @ w.double()).cpu().numpy()
    for fam in (capi.GEMV_SIMT, capi.GEMV_MMA, capi.GEMV_PERSIST):
        y = X.gemv(a, qw, s, qz, g, bits, K, 1, family=fam)
        assert_gemv_close(y.cpu().numpy(), truth, f"{K}x{N} family {fam}", floor_of(fam))
        if fam == capi.GEMV_PERSIST:
            # M <= 2 runs the integer block math, M >= 3 the fp16 exact-product math: rows agree bit for bit within each
            # -- under the same CTA-boundary mode (the planner may pick block-granular ranges for one row and tile-aligned
            # ones for two: another fp32 summation order, i.e. fp16 rounding of the same sums)
            for mode in (0, 1):
                capi.set_option("XBIT_W4P_FINE", mode)
                try:
                    y1 = X.gemv(a[1:2], qw, s, qz, g, bits, K, 1, family=fam)
                    y12 = X.gemv(a[:2], qw, s, qz, g, bits, K, 1, family=fam)
                finally:
                    capi.set_option("XBIT_W4P_FINE")
                assert torch.equal(y1[0], y12[1]), f"{K}x{N} fine={mode}"
            y1 = X.gemv(a[1:2], qw, s, qz, g, bits, K, 1, family=fam)[0].double()
            y12 = X.gemv(a[:2], qw, s, qz, g, bits, K, 1, family=fam)[1].double()
            assert float((y1 - y12).abs().max()) <= 2.0 ** -9 * float(y12.abs().max())
            # (3 and 4 rows may get different warp counts / rings, i.e. another fp32 summation order: fp16 rounding of the
            # same sums, not bit identity)
            y3 = X.gemv(a[1:4], qw, s, qz, g, bits, K, 1, family=fam)[1].double()
            assert float((y3 - y[2].double()).abs().max()) <= 2.0 ** -9 * float(y[2].double().abs().max())
        else:
            y1 = X.gemv(a[2:3], qw, s, qz, g, bits, K, 1, family=fam)
            assert torch.equal(y1[0], y[2])
        y2 = X.gemv(a[:1] * 2, qw, s, qz, g, bits, K, 1, family=fam).double()
        yd = X.gemv(a[:1], qw, s, qz, g, bits, K, 1, family=fam).double() * 2
        # exact up to fp16 subnormal effects in the staged activations (a/16 and the low half of
        # sum_k a_k lose bits below 2^-24, which does not scale): at most one fp16 ulp, and rarely
        assert float((y2 - yd).abs().max()) <= 2.0 ** -10 * float(yd.abs().max())
        assert float((y2 != yd).double().mean()) <= 0.02
        half = N // 2
        ysl = X.gemv(a[:1], qw[:, half:].contiguous(), s[:, half:].contiguous(), qz[:, half // 8:].contiguous(),
                     g, bits, K, 1, family=fam)
        assert_gemv_close(ysl.cpu().numpy(), truth[:1, half:], f"{K}x{N} column shard, family {fam}", floor_of(fam))


@pytest.mark.parametrize("bits", range(2, 9))
def test_dequant_full_size_bands_vs_oracle(bits, dev, c_oracle):
    """BASELINE config 3 at its real size (4096 x 11008), every bit width and group size: three 256-row bands of the
    output (first, one that straddles group and word-row boundaries of every width, last) bit for bit against the C
    oracle (oracle.dequant with k_range restates only the band, so the CPU side stays a few seconds)."""
    K, N = 4096, 11008
    for g in (32, 64, 128):
        qw, s, qz, _ = synth.make_inputs(K, N, bits, g, seed=100 + bits + g)
        got = X.dequant(ti(qw, dev), t16(s, dev), ti(qz, dev), g, bits, K, 1).cpu().numpy().view(np.uint16)
        for k0 in (0, 1888, K - 256):
            want = c_oracle.dequant(qw, s, qz, g, bits, K, 1, k_range=(k0, k0 + 256)).view(np.uint16)
            assert (got[k0:k0 + 256] == want[k0:k0 + 256]).all(), f"bits={bits} g={g} rows {k0}..{k0 + 256}"


def test_gemv_auto_never_fails_for_large_m_times_k(dev, c_oracle):
    """M = 16 / 17 with K = 32768 has no K split that stages the activations in shared memory for the
    tensor-core kernels; AUTO must fall back to smaller row slabs (or the generic kernel), not fail."""
    K, N, g = 32768, 64, 128
    qw, s, qz, a = synth.make_inputs(K, N, 4, g, M=17, seed=11)
    w = c_oracle.dequant(qw, s, qz, g, 4, K, 1)
    tq, ts, tz = ti(qw, dev), t16(s, dev), ti(qz, dev)
    for M in (16, 17, 8):
        y64 = a[:M].astype(np.float64) @ w.astype(np.float64)
        got = X.gemv(t16(a[:M], dev), tq, ts, tz, g, 4, K, 1).cpu().numpy()
        assert_gemv_close(got, y64, f"AUTO M={M} K={K}")


def test_gemv_persistent_schedule_is_race_free_under_repetition(dev):
    """compute-sanitizer is closed on the GPU pool (tools/sanitize.sh documents the attempt), so the shared-memory and
    cross-CTA hand-offs of the persistent kernel are checked by repetition instead: 60 back-to-back launches per shape
    (programmatic dependent launch, the workspace and the rings reused by overlapping launches) must reproduce the
    first result bit for bit -- any race in the partial-tile protocol shows up as a differing sum."""
    gen = torch.Generator(device=dev).manual_seed(5)
    for (K, N, M, g, fine) in ((4096, 4096, 1, 128, 0), (4096, 11008, 1, 128, 1), (11008, 4096, 2, 128, 1), (2048, 4736, 3, 64, 1),
                              (8192, 8192, 1, 128, 1)):
        qw = torch.randint(-2**31, 2**31 - 1, (K // 8, N), dtype=torch.int32, device=dev, generator=gen)
        qz = torch.randint(-2**31, 2**31 - 1, (K // g, N // 8), dtype=torch.int32, device=dev, generator=gen)
        s = (torch.rand((K // g, N), device=dev, generator=gen) * 0.018 + 0.002).to(torch.float16)
        a = torch.randn((M, K), device=dev, generator=gen).to(torch.float16)
        capi.set_option("XBIT_W4P_FINE", fine)
        X.set_static_weights(True)
        try:
            outs = torch.empty((60, M, N), dtype=torch.float16, device=dev)
            for i in range(60):
                X.gemv(a, qw, s, qz, g, 4, K, 1, family=capi.GEMV_PERSIST, out=outs[i])
            torch.cuda.synchronize()
        finally:
            X.set_static_weights(False)
            capi.set_option("XBIT_W4P_FINE")
        assert bool((outs == outs[0]).all()), f"{K}x{N} M={M} fine={fine}: repeated launches differ"
        truth = a.double() @ X.dequant(qw, s, qz, g, 4, K, 1).double()
        assert_gemv_close(outs[0].cpu().numpy(), truth.cpu().numpy(), f"{K}x{N} M={M}")


def test_gemv_runs_on_current_stream_and_is_graph_capturable(dev):
    """The reference launches gemv on the legacy default stream (gemv_w4a16_pt.cu:162); ours must
    follow torch's current stream and be capturable (no sync, no allocation inside the C ABI)."""
    K, N, g = 4096, 4096, 128
    qw = torch.randint(-2**31, 2**31 - 1, (K // 8, N), dtype=torch.int32, device=dev)
    qz = torch.randint(-2**31, 2**31 - 1, (K // g, N // 8), dtype=torch.int32, device=dev)
    s = (torch.rand((K // g, N), device=dev) * 0.018 + 0.002).to(torch.float16)
    a = torch.randn((1, K), device=dev).to(torch.float16)
    eager = X.gemv(a, qw, s, qz, g, 4, K, 0)
    out = torch.empty_like(eager)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        X.gemv(a, qw, s, qz, g, 4, K, 0, out=out)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    out2 = torch.empty_like(eager)
    with torch.cuda.graph(graph):
        for _ in range(3):
            X.gemv(a, qw, s, qz, g, 4, K, 0, out=out2)
    out2.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager) and torch.equal(out2, eager)


def test_errors_are_runtime_errors_not_aborts(dev):
    qw = torch.zeros(16, 8, dtype=torch.int32, device=dev)
    s = torch.ones(1, 8, dtype=torch.float16, device=dev)
    qz = torch.zeros(1, 1, dtype=torch.int32, device=dev)
    with pytest.raises(RuntimeError, match="in_features"):
        X.dequant(qw, s, qz, 128, 4, 64, 0)
    with pytest.raises(RuntimeError, match="groupsize"):
        X.dequant(qw, s, qz, 8, 4, 128, 0)
    with pytest.raises(RuntimeError, match="bits"):
        X.dequant(qw, s, qz, 128, 9, 128, 0)
    with pytest.raises(RuntimeError, match="contiguous"):
        X.dequant(qw.t(), s, qz, 128, 4, 128, 0)
    with pytest.raises(RuntimeError):
        X.dequant(torch.zeros(4, 8, dtype=torch.int32, device=dev), s, qz, 128, 1, 128, 0)   # bits=1 unsupported
    with pytest.raises(RuntimeError, match="float16"):
        X.gemv(torch.zeros(1, 128, device=dev), qw, s, qz, 128, 4, 128, 0)
    # still alive and correct afterwards
    y = X.gemv(torch.ones(1, 128, dtype=torch.float16, device=dev), qw, s, qz, 128, 4, 128, 0)
    assert torch.equal(y, torch.zeros_like(y))


def test_host_buffer_entry_point(dev, c_oracle):
    """xbit_gemv_f16_host: activations from pinned host memory, result back to host (the e2e leg)."""
    K, N, g = 1024, 256, 128
    qw, s, qz, a = synth.make_inputs(K, N, 4, g, M=1, seed=1)
    y64, _ = c_oracle.gemv(a, qw, s, qz, g, 4, K, 0)
    tq, ts, tz = ti(qw, dev), t16(s, dev), ti(qz, dev)
    ha = torch.from_numpy(a.view(np.int16)).pin_memory()
    ho = torch.empty((1, N), dtype=torch.int16).pin_memory()
    da = torch.empty((1, K), dtype=torch.float16, device=dev)
    do = torch.empty((1, N), dtype=torch.float16, device=dev)
    lib = capi.load()
    st = torch.cuda.current_stream().cuda_stream
    capi.check(lib.xbit_gemv_f16_host(ha.data_ptr(), ho.data_ptr(), da.data_ptr(), do.data_ptr(), tq.data_ptr(),
                                      ts.data_ptr(), tz.data_ptr(), 1, K, N, 4, g, 0, None, 0, st))
    torch.cuda.synchronize()
    assert_gemv_close(ho.numpy().view(np.float16), y64, "host entry")          # pinned result: stored by the kernel itself
    assert not bool(torch.equal(do.view(torch.int16).cpu(), ho))                  # ... the staging buffer was not used
    hp = np.zeros((1, N), dtype=np.int16)                                          # pageable result: staged + D2H copy
    capi.check(lib.xbit_gemv_f16_host(ha.data_ptr(), hp.ctypes.data, da.data_ptr(), do.data_ptr(), tq.data_ptr(),
                                      ts.data_ptr(), tz.data_ptr(), 1, K, N, 4, g, 0, None, 0, st))
    torch.cuda.synchronize()
    assert np.array_equal(hp, ho.numpy())


def test_flag_in_data_chain_single_gpu(dev):
    """xbit_gemv_f16_peers_ll / xbit_ll_unpack_f16 with world = 1: two dependent calls, the second one reads the
    first one's {results, call number} slots while staging its activations (the 2-GPU version is in
    test_multigpu.py); and the signal form with its wait kernel."""
    import ctypes
    K = N = 2048
    M = 3
    qw, s, qz, a = synth.make_inputs(K, N, 4, 128, M=M, seed=5)
    tq, ts, tz, ta = ti(qw, dev), t16(s, dev) * 0.05, ti(qz, dev), t16(a, dev)
    y1 = X.gemv(ta, tq, ts, tz, 128, 4, K, 1)
    y2 = X.gemv(y1, tq, ts, tz, 128, 4, K, 1)
    lib = capi.load()
    st = torch.cuda.current_stream().cuda_stream
    ll = torch.zeros((2, M, N), dtype=torch.int32, device=dev)
    state = torch.zeros(4, dtype=torch.int32, device=dev)
    out = torch.empty((M, N), dtype=torch.float16, device=dev)
    for call, (src, flag) in enumerate(((ta.data_ptr(), 0), (ll[0].data_ptr(), capi.GEMV_FLAG_A_IS_LL))):
        outs = (ctypes.c_void_p * 1)(ll[call].data_ptr())
        capi.check(lib.xbit_gemv_f16_peers_ll(src, tq.data_ptr(), ts.data_ptr(), tz.data_ptr(), outs, state.data_ptr(), call, 1, 0, M, K, N,
                                              4, 128, 1, N, 0, capi.GEMV_AUTO | flag, st))
    capi.check(lib.xbit_ll_unpack_f16(ll[1].data_ptr(), out.data_ptr(), M * N, state.data_ptr(), 2, state.data_ptr() + 12, st))
    torch.cuda.synchronize()
    assert state.tolist() == [0, 2, 2, 0]                        # two calls, chain base advanced by two, no timeout
    slots = ll[0].view(M, N // 2, 2)
    assert bool((slots[..., 1] == 1).all()) and bool((ll[1].view(M, N // 2, 2)[..., 1] == 2).all())
    # (the chain planner may pick another K split than the plain call: same block math, other fp32 summation order)
    got1 = slots[..., 0].contiguous().view(torch.float16).view(M, N)
    assert float((got1.double() - y1.double()).abs().max()) <= 2e-3 * float(y1.double().abs().max())
    assert float((out.double() - y2.double()).abs().max()) <= 4e-3 * float(y2.double().abs().max())
    # and against the oracle-defined truth (fp64 dot over the bit-exact dequantised weights), call by call
    w64 = X.dequant(tq, ts, tz, 128, 4, K, 1).double()
    t1 = ta.double() @ w64
    assert_gemv_close(got1.cpu().numpy(), t1.cpu().numpy(), "LL chain call 0")
    t2 = got1.double() @ w64                                    # the second call's input is the first call's fp16 output
    assert_gemv_close(out.cpu().numpy(), t2.cpu().numpy(), "LL chain call 1")
    # restrictions are reported, not executed
    rc = lib.xbit_gemv_f16_peers_ll(ta.data_ptr(), tq.data_ptr(), ts.data_ptr(), tz.data_ptr(), (ctypes.c_void_p * 1)(ll[0].data_ptr()),
                                    state.data_ptr(), 0, 1, 0, M, K, N, 3, 128, 1, N, 0, capi.GEMV_AUTO, st)
    assert rc == -1 and "bits" in capi.last_error()          # XBIT_EINVAL
    # signal form, world = 1: the flag carries the call count, the wait kernel returns at once
    flags = torch.zeros(8, dtype=torch.int32, device=dev)
    state2 = torch.zeros(4, dtype=torch.int32, device=dev)
    y = torch.empty((M, N), dtype=torch.float16, device=dev)
    for _ in range(2):
        capi.check(lib.xbit_gemv_f16_peers_signal(ta.data_ptr(), tq.data_ptr(), ts.data_ptr(), tz.data_ptr(),
                                                  (ctypes.c_void_p * 1)(y.data_ptr()), (ctypes.c_void_p * 1)(flags.data_ptr()),
                                                  state2.data_ptr(), 1, 0, M, K, N, 4, 128, 1, N, 0,
                                                  capi.GEMV_AUTO | capi.GEMV_FLAG_WAIT_PEERS, st))
    capi.check(lib.xbit_peers_wait(flags.data_ptr(), 1, 0, state2.data_ptr() + 12, st))
    torch.cuda.synchronize()
    assert int(flags[0]) == 2 and state2.tolist() == [0, 2, 0, 0]
    # (the plain call may run the persistent schedule, the signal form is raised by the cluster kernel: other summation order)
    assert float((y.double() - y1.double()).abs().max()) <= 2e-3 * float(y1.double().abs().max())


def test_sharded_chain_wrapper_single_gpu(dev):
    """ShardedQChain (world = 1): 2048 -> 4096 -> 2048 -> 2048, repeated calls (chain base advances), M = 2."""
    from xbitops_b200.sharded import ShardedQChain
    dims = (2048, 4096, 2048, 2048)
    layers, ref_layers = [], []
    for i in range(3):
        K, N = dims[i], dims[i + 1]
        qw, s, qz, _ = synth.make_inputs(K, N, 4, 128, M=1, seed=40 + i)
        tq, ts, tz = ti(qw, dev), t16(s, dev) * 0.05, ti(qz, dev)
        layers.append((tq, ts, tz, K, N))
    a = synth.make_inputs(dims[0], 32, 4, 128, M=2, seed=9)[3]
    ta = t16(a, dev)
    ref = ta
    for (tq, ts, tz, K, N) in layers:
        ref = X.gemv(ref, tq, ts, tz, 128, 4, K, 1)
    chain = ShardedQChain(layers, 128, 4, 1, max_rows=2)
    for _ in range(8):                                                                     # back-to-back chains
        y = chain(ta)
    torch.cuda.synchronize()
    assert tuple(y.shape) == (2, dims[-1]) and chain._bufs[2].tolist()[2:] == [24, 0]     # 8 chains of 3 calls, no timeout
    assert float((y.double() - ref.double()).abs().max()) <= 6e-3 * float(ref.double().abs().max())
    with pytest.raises(ValueError):
        ShardedQChain([layers[0], layers[2]], 128, 4, 1)                                   # 4096 != 2048: not a chain


def test_peer_entry_point_single_process(dev, c_oracle):
    """xbit_gemv_f16_peers with world=2 emulated inside one process: two 'rank' buffers on the same
    device, each shard's epilogue stores its slice into both (multi-process NVLink is in test_multigpu)."""
    K, N, g, M = 2048, 512, 128, 2
    qw, s, qz, a = synth.make_inputs(K, N, 4, g, M=M, seed=2)
    y64, _ = c_oracle.gemv(a, qw, s, qz, g, 4, K, 0)
    ta = t16(a, dev)
    bufs = [torch.zeros((M, N), dtype=torch.float16, device=dev) for _ in range(2)]
    arr = (ctypes.c_void_p * 2)(*[b.data_ptr() for b in bufs])
    lib = capi.load()
    half = N // 2
    for rank in range(2):
        sl = slice(rank * half, (rank + 1) * half)
        tq, ts = ti(qw[:, sl], dev), t16(s[:, sl], dev)
        tz = ti(qz[:, rank * half // 8:(rank + 1) * half // 8], dev)
        capi.check(lib.xbit_gemv_f16_peers(ta.data_ptr(), tq.data_ptr(), ts.data_ptr(), tz.data_ptr(), arr, 2, M, K, half,
                                           4, g, 0, N, rank * half, None, 0, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(bufs[0], bufs[1])
    assert_gemv_close(bufs[0].cpu().numpy(), y64, "peers")


# ------------------------------------------------------------------ bf16-native

def _bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)


@pytest.mark.parametrize("bits", range(2, 9))
def test_dequant_bf16_native_bit_exact_vs_oracle(bits, dev):
    """xbit_dequant_bf16: out = RN_bf16((w - z) * s), bit for bit against the numpy restatement, block kernel and element
    fallback, incl. scales beyond fp16's range (where the reference's fp16 round trip returns inf)."""
    from oracle import oracle as orc
    X.set_native_bf16(True)
    try:
        for (K, N, g) in ((256, 64, 128), (416, 136, 32), (1024 + 32, 1000, 64), (300, 50, 100)):
            for bias in (0, 1):
                qw, s, qz, _ = synth.make_inputs(K, N, bits, g, seed=bits * 7 + g)
                sb = torch.from_numpy(s.astype(np.float32)).to(torch.bfloat16)
                if bias:
                    sb = sb * 2.0 ** 20                                   # far outside fp16
                want = orc.np_dequant_bf16_native(qw, _bf16_bits(sb), qz, g, bits, K, bias)
                got = X.dequant(ti(qw, dev), sb.to(dev), ti(qz, dev), g, bits, K, bias)
                assert got.dtype == torch.bfloat16 and tuple(got.shape) == (K, N)
                assert (_bf16_bits(got) == want).all(), (bits, K, N, g, bias)
                assert torch.isfinite(got.float()).all()
    finally:
        X.set_native_bf16(False)


@pytest.mark.parametrize("M", (1, 2, 5))
def test_gemv_bf16_native_vs_fp64_truth(M, dev):
    """xbit_gemv_bf16 (bits 4, groupsize 128): against the fp64 product of the bf16 inputs.  Tolerance: one bf16 rounding
    of the result (2^-9 relative) plus the kernel's accumulation error -> 4e-3 of the largest |y|; and the
    same call with scales far outside fp16's range stays finite where the reference's fp16 arithmetic overflows."""
    from oracle import oracle as orc
    X.set_native_bf16(True)
    try:
        for (K, N, bits) in ((512, 256, 4), (4096, 4096, 4), (4096, 11008, 4), (11008, 4096, 4), (512, 256, 8), (4096, 4096, 8), (4224, 4128, 8), (512, 256, 2), (4096, 11008, 2)):
            for boost in (1.0, 2.0 ** 18):
                qw, s, qz, _ = synth.make_inputs(K, N, bits, 128, seed=K + N + M)
                sb = (torch.from_numpy(s.astype(np.float32)) * boost).to(torch.bfloat16)
                gen = torch.Generator().manual_seed(K)
                a = torch.randn((M, K), generator=gen).to(torch.bfloat16)
                y = X.gemv(a.to(dev), ti(qw, dev), sb.to(dev), ti(qz, dev), 128, bits, K, 1)
                assert y.dtype == torch.bfloat16 and tuple(y.shape) == (M, N)
                w = orc.np_unpack_qweight(qw, K, bits).astype(np.float64)
                z = orc.np_unpack_qzeros(qz, N, bits).astype(np.float64) + 1
                grp = np.arange(K) // 128
                truth = a.double().numpy() @ ((w - z[grp]) * sb.double().numpy()[grp])
                got = y.double().cpu().numpy()
                assert np.isfinite(got).all()
                err = np.abs(got - truth).max() / np.abs(truth).max()
                assert err <= 4e-3, (K, N, bits, M, boost, err)
    finally:
        X.set_native_bf16(False)


# ------------------------------------------------------------------ A16W8 on the persistent kernel

@pytest.mark.parametrize("wbits", (8, 2))
def test_gemv_w8_persistent_integer_math(wbits, dev, c_oracle):
    """8-bit weights, groupsize 128 (the reference aborts on everything but bits 4, gemv_w4a16_pt.cu:152-155): AUTO routes
    M <= 2 (and larger batches two rows at a time) to the persistent kernel, whose integer block math takes the packed
    words as MMA operands unchanged.  Against the fp64 product over the oracle's dequantised weights; both CTA-boundary
    modes; deterministic; the generic kernel agrees within the tolerance."""
    lib = capi.load()
    cases = ((512, 256, 1, 1), (4096, 4096, 1, 0), (4096, 4096, 2, 1), (4096, 11008, 1, 1), (11008, 4096, 2, 0), (1024, 96, 5, 1),
             (4224, 4128, 3, 0))
    for (K, N, M, bias) in cases:
        qw, s, qz, a = synth.make_inputs(K, N, wbits, 128, M=M, seed=K + N + M + wbits)
        w = c_oracle.dequant(qw, s, qz, 128, wbits, K, bias)
        y64 = a.astype(np.float64) @ w.astype(np.float64)
        tq, ts, tz, ta = ti(qw, dev), t16(s, dev), ti(qz, dev), t16(a, dev)
        assert lib.xbit_gemv_pick_family(min(M, 2), K, N, wbits, 128) == capi.GEMV_PERSIST
        y_auto = X.gemv(ta, tq, ts, tz, 128, wbits, K, bias)
        assert_gemv_close(y_auto.cpu().numpy(), y64, f"W{wbits} AUTO {K}x{N} M={M}")
        for fine in (1, 0):
            capi.set_option("XBIT_W4P_FINE", fine)
            try:
                y1 = X.gemv(ta, tq, ts, tz, 128, wbits, K, bias, family=capi.GEMV_PERSIST)
                y2 = X.gemv(ta, tq, ts, tz, 128, wbits, K, bias, family=capi.GEMV_PERSIST)
            finally:
                capi.set_option("XBIT_W4P_FINE")
            assert torch.equal(y1, y2)
            assert_gemv_close(y1.cpu().numpy(), y64, f"W{wbits} persist fine={fine} {K}x{N} M={M}")
        yg = X.gemv(ta, tq, ts, tz, 128, wbits, K, bias, family=capi.GEMV_GENERIC)
        assert float((yg.double() - y_auto.double()).abs().max()) <= 2e-3 * float(np.abs(y64).max())
        # back-to-back launches (programmatic dependent launch, rings and workspace reused) reproduce the result bit for bit
        X.set_static_weights(True)
        try:
            outs = torch.empty((30,) + tuple(y_auto.shape), dtype=torch.float16, device=dev)
            for i in range(30):
                X.gemv(ta, tq, ts, tz, 128, wbits, K, bias, out=outs[i])
            torch.cuda.synchronize()
        finally:
            X.set_static_weights(False)
        assert bool((outs == outs[0]).all()), f"W{wbits} {K}x{N} M={M}: repeated launches differ"
