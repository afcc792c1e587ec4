"""QLinear and the device-side packer.  CPU: pack / unpack against the numpy packer and the C
oracle's unpacker.  GPU: forward against the dequantised dense layer, both dispatch branches, bias, bf16, act-order."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from xbitops_b200 import synth  # noqa: E402
from xbitops_b200 import qlinear as Q  # noqa: E402


@pytest.mark.parametrize("bits", range(2, 9))
def test_torch_packer_matches_numpy_packer_and_oracle(bits, c_oracle):
    rng = np.random.default_rng(bits)
    K, N, g = 200, 72, 40
    w = rng.integers(0, 1 << bits, size=(K, N), dtype=np.uint8)
    z = rng.integers(0, 1 << bits, size=(K // g, N), dtype=np.uint8)
    qw, qz = Q.pack_qweight(torch.from_numpy(w), bits), Q.pack_qzeros(torch.from_numpy(z), bits)
    assert np.array_equal(qw.numpy(), synth.pack_qweight(w, bits)) and np.array_equal(qz.numpy(), synth.pack_qzeros(z, bits))
    assert np.array_equal(c_oracle.unpack_qweight(qw.numpy(), K, bits), w) and np.array_equal(c_oracle.unpack_qzeros(qz.numpy(), N, bits), z)
    assert np.array_equal(Q.unpack_stream(qw, K, bits, 0).numpy(), w) and np.array_equal(Q.unpack_stream(qz, N, bits, 1).numpy(), z)


def test_rtn_quantiser_matches_numpy_quantiser():
    w = np.random.default_rng(0).standard_normal((256, 48)).astype(np.float32)
    for bits in (3, 4, 8):
        qw, s, qz = Q.quantize_rtn(torch.from_numpy(w), bits, 64, 1)
        nqw, ns, nqz = synth.quantize(w, bits, 64, 1)
        assert np.array_equal(s.numpy().view(np.uint16), ns.view(np.uint16))
        assert np.array_equal(qw.numpy(), nqw) and np.array_equal(qz.numpy(), nqz)


@pytest.mark.gpu
def test_qlinear_forward_matches_dense(c_oracle):
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test without a CUDA device (no CPU fallback exists)")
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(3)
    K, N = 1024, 768
    lin = torch.nn.Linear(K, N, bias=True, device=dev, dtype=torch.float16)
    for bits, g in ((4, 128), (3, 64), (8, 32), (8, 128), (2, 128)):       # the last two: the fast 8- / 2-bit kernels
        q = Q.QLinear.from_linear(lin, bits, g)
        w = q.dequantized_weight()                                   # [K, N] fp16
        # the packed tensors restated by the C oracle give the same dense weight, bit for bit
        want = c_oracle.dequant(q.qweight.cpu().numpy(), q.scales.cpu().numpy(), q.qzeros.cpu().numpy(), g, bits, K, 1)
        assert np.array_equal(w.cpu().numpy().view(np.uint16), want.view(np.uint16))
        for rows in (1, 5, 16, 40):                                  # gemv branch up to 16 rows, dequant + matmul above
            x = torch.randn((rows, K), device=dev, generator=gen).to(torch.float16)
            ref = x.double() @ w.double() + lin.bias.double()
            y = q(x)
            assert y.shape == (rows, N) and y.dtype == torch.float16
            assert float((y.double() - ref).abs().max() / ref.abs().max()) < 1e-2
        y3 = q(torch.randn((2, 3, K), device=dev, generator=gen).to(torch.float16))
        assert y3.shape == (2, 3, N)
    # bf16 layer: output dtype follows the scales (dq_torch_ops.cc:33-42)
    qb = Q.QLinear.from_linear(lin.to(torch.bfloat16), 4, 128)
    xb = torch.randn((2, K), device=dev, generator=gen).to(torch.bfloat16)
    yb = qb(xb)
    assert yb.dtype == torch.bfloat16
    # ... and with the bf16-native kernels switched on the layer stays in bf16 end to end
    from xbitops_b200 import ops
    ops.set_native_bf16(True)
    try:
        yn = qb(xb)
        wn = qb.dequantized_weight()
    finally:
        ops.set_native_bf16(False)
    assert yn.dtype == torch.bfloat16 and wn.dtype == torch.bfloat16
    refb = xb.double() @ wn.double() + qb.bias.double()
    assert float((yn.double() - refb).abs().max() / refb.abs().max()) < 1e-2
    assert float((yn.double() - yb.double()).abs().max() / refb.abs().max()) < 2e-2


@pytest.mark.gpu
def test_qlinear_act_order():
    """g_idx of a desc_act checkpoint: channels of a group scattered over k.  The layer re-packs once and gathers x."""
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test without a CUDA device (no CPU fallback exists)")
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(4)
    K, N, bits, g = 512, 256, 4, 128
    w = torch.randn((K, N), device=dev, generator=gen)
    order = torch.randperm(K, device=dev, generator=gen)               # quantisation order: group j = channels order[j*g:(j+1)*g]
    qw_sorted, s, qz = Q.quantize_rtn(w[order], bits, g, 1)             # what GPTQ computes, groups contiguous in ITS order
    g_idx = torch.empty(K, dtype=torch.long, device=dev)
    g_idx[order] = torch.arange(K, device=dev) // g
    # the checkpoint stores the rows in ORIGINAL channel order with g_idx beside them
    rows_sorted = Q.unpack_stream(qw_sorted, K, bits, 0)
    rows_orig = torch.empty_like(rows_sorted)
    rows_orig[order] = rows_sorted
    q = Q.QLinear.from_packed(Q.pack_qweight(rows_orig, bits), s, qz, bits, g, K, g_idx=g_idx)
    assert q.perm is not None
    # dense restatement: row k uses the scale / zero of group g_idx[k]
    zi = Q.unpack_stream(qz, N, bits, 1).float() + 1
    dense = ((rows_orig.float() - zi[g_idx]) * s.float()[g_idx])
    x = torch.randn((3, K), device=dev, generator=gen).to(torch.float16)
    ref = x.double() @ dense.double()
    assert float((q(x).double() - ref).abs().max() / ref.abs().max()) < 1e-2
    assert float((q.dequantized_weight().double() - dense.double()).abs().max()) < 2e-3 * float(dense.abs().max())
