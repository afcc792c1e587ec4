// dq_torch_ops.cc -- the PyTorch operator surface, re-registered over the torch-free C ABI.
//
// Mirrors /root/reference/src/dq_torch_ops.cc: module name "XbitOps" (:80, setup.py:100), ops
//   dequant(qweight, scales, qzeros, groupsize, bits, in_features, add_zero_bias) -> Tensor[K, N]   (:23-44)
//   gemv(input_a, qweight, scales, qzeros, groupsize, bits, in_features, add_zero_bias) -> Tensor    (:46-78)
// positional-only, same argument order, same output shapes/dtypes (output dtype follows `scales`;
// bf16 scales are computed in fp16 and the result is cast back, :33-42, :65-76; 3-D activations
// [B, S, K] flatten to M = B*S and the output regains [B, S, N], :59-64).
// Differences, all deliberate:
//   * outputs are at::empty (every element is written) instead of at::zeros (:38, :71);
//   * the device is guarded and restored (the reference calls cudaSetDevice and leaves it, :32, :58);
//   * both ops run on the CURRENT torch stream (the reference's gemv uses the legacy default
//     stream, gemv_w4a16_pt.cu:162);
//   * dtypes and the scales/qzeros shapes are validated explicitly; failures are RuntimeError
//     (never exit()/abort());
//   * gemv accepts every bits in [2, 8] and any groupsize >= 16 (the reference aborts unless
//     bits == 4 && groupsize == 128, gemv_w4a16_pt.cu:152-155);
//   * one extra module function, set_static_weights(bool) (default False, or env XBIT_STATIC_WEIGHTS=1 at import):
//     the caller's promise that qweight / scales / qzeros are resident model weights, never written by the kernel
//     that precedes a gemv call in its stream.  gemv then prefetches them under programmatic dependent launch while
//     that kernel drains (XBIT_GEMV_FLAG_STATIC_WEIGHTS); activations are still read only after it has finished.
// No kernel lives here: this file only validates, allocates and calls include/xbitops_b200.h.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <atomic>
#include <map>
#include <mutex>
#include <utility>

#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../../include/xbitops_b200.h"

#define CHECK_CUDA(x) TORCH_CHECK(x.device().is_cuda(), #x " must be a CUDA tensor")
#define CHECK_CONTIGUOUS(x) TORCH_CHECK(x.is_contiguous(), #x " must be contiguous")
#define CHECK_INPUT(x) \
  CHECK_CUDA(x);       \
  CHECK_CONTIGUOUS(x)

namespace {

void check_quant_args(const torch::Tensor& qweight, const torch::Tensor& scales, const torch::Tensor& qzeros,
                      int groupsize, int bits, int in_features) {
  CHECK_INPUT(qweight);
  CHECK_INPUT(scales);
  CHECK_INPUT(qzeros);
  TORCH_CHECK(qweight.dim() == 2, "qweight must be 2-dimensional");
  TORCH_CHECK(groupsize >= 16, "groupsize must be >= 16");
  TORCH_CHECK(bits >= 1 && bits <= 8, "bits must be >= 1 and <= 8");
  TORCH_CHECK(((int64_t)in_features * bits + 31) / 32 == qweight.size(0), "in_features must be >= 1");
  TORCH_CHECK(qweight.scalar_type() == torch::kInt32, "qweight must be int32");
  TORCH_CHECK(qzeros.scalar_type() == torch::kInt32, "qzeros must be int32");
  TORCH_CHECK(scales.scalar_type() == torch::kFloat16 || scales.scalar_type() == torch::kBFloat16,
              "scales must be float16 or bfloat16");
  const int64_t n = qweight.size(1);
  const int64_t groups = (in_features + groupsize - 1) / groupsize;
  TORCH_CHECK(scales.dim() == 2 && scales.size(0) >= groups && scales.size(1) == n,
              "scales must be [ceil(in_features/groupsize), out_features]");
  TORCH_CHECK(qzeros.dim() == 2 && qzeros.size(0) >= groups && qzeros.size(1) == (n * bits + 31) / 32,
              "qzeros must be [ceil(in_features/groupsize), ceil(out_features*bits/32)]");
  TORCH_CHECK(scales.device() == qweight.device() && qzeros.device() == qweight.device(),
              "qweight, scales and qzeros must be on the same device");
}

std::atomic<bool> g_static_weights{[] {
  const char* v = std::getenv("XBIT_STATIC_WEIGHTS");
  return v && *v && *v != '0';
}()};

// bf16 without the fp16 round trip (default off = the reference's behaviour, :33-42, :65-76):
// dequant with bf16 scales -> xbit_dequant_bf16; gemv with bf16 activations and scales -> xbit_gemv_bf16 where it exists
std::atomic<bool> g_native_bf16{[] {
  const char* v = std::getenv("XBIT_NATIVE_BF16");
  return v && *v && *v != '0';
}()};

void raise_if(int rc) { TORCH_CHECK(rc == XBIT_OK, "xbitops_b200: ", xbit_last_error()); }

// Scratch for the persistent stream-K schedule (include/xbitops_b200.h: xbit_gemv_workspace_bytes):
// zero-initialised once, left zeroed by every call, one buffer per (device, stream) because calls on
// one stream are ordered and calls on different streams must not share flags.  The map is leaked on
// purpose: tensors must not be destroyed after the CUDA context at interpreter exit.
at::Tensor gemv_workspace(const at::Device& device, cudaStream_t stream, size_t nbytes) {
  static std::mutex mu;
  static auto* cache = new std::map<std::pair<int, cudaStream_t>, at::Tensor>();
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair((int)device.index(), stream);
  auto it = cache->find(key);
  if (it == cache->end() || (size_t)it->second.numel() < nbytes) {
    at::Tensor t = at::zeros({(int64_t)nbytes}, at::TensorOptions().dtype(at::kByte).device(device));
    it = cache->insert_or_assign(key, t).first;
  }
  return it->second;
}

// (internal linkage: the reference extension exports functions of the same names, and both modules
// can live in one process -- the parity tests load them side by side)

torch::Tensor dequant_any_bit(const torch::Tensor& qweight, const torch::Tensor& scales, const torch::Tensor& qzeros,
                              int groupsize, int bits, int in_features, uint8_t add_zero_bias) {
  check_quant_args(qweight, scales, qzeros, groupsize, bits, in_features);
  const c10::cuda::CUDAGuard guard(qweight.device());
  auto stream = at::cuda::getCurrentCUDAStream().stream();
  if (g_native_bf16.load() && scales.scalar_type() == torch::kBFloat16) {
    at::Tensor out = at::empty({in_features, qweight.size(1)}, scales.options());
    raise_if(xbit_dequant_bf16(qweight.data_ptr<int32_t>(), scales.data_ptr(), qzeros.data_ptr<int32_t>(), out.data_ptr(),
                               in_features, (int)qweight.size(1), bits, groupsize, add_zero_bias,
                               reinterpret_cast<xbit_stream_t>(stream)));
    return out;
  }
  auto f16_scale = scales;
  const auto ori_dtype = scales.scalar_type();
  if (ori_dtype == torch::kBFloat16) f16_scale = scales.to(torch::kFloat16);
  at::Tensor output = at::empty({in_features, qweight.size(1)}, f16_scale.options());
  raise_if(xbit_dequant_f16(qweight.data_ptr<int32_t>(), f16_scale.data_ptr(), qzeros.data_ptr<int32_t>(),
                            output.data_ptr(), in_features, (int)qweight.size(1), bits, groupsize, add_zero_bias,
                            reinterpret_cast<xbit_stream_t>(stream)));
  if (ori_dtype == torch::kBFloat16) output = output.to(torch::kBFloat16);
  return output;
}

torch::Tensor op_gemv(const torch::Tensor& input_a, const torch::Tensor& qweight, const torch::Tensor& scales,
                      const torch::Tensor& qzeros, int groupsize, int bits, int in_features, uint8_t add_zero_bias) {
  CHECK_INPUT(input_a);
  check_quant_args(qweight, scales, qzeros, groupsize, bits, in_features);
  TORCH_CHECK(qweight.device().index() == input_a.device().index(), "input and weight must be on the same device");
  const bool native = g_native_bf16.load() && input_a.scalar_type() == torch::kBFloat16 && scales.scalar_type() == torch::kBFloat16 &&
                      (bits == 2 || bits == 4 || bits == 8) && groupsize == 128 && in_features % 128 == 0 && qweight.size(1) % 32 == 0 && in_features <= 16384;
  TORCH_CHECK(native || input_a.scalar_type() == torch::kFloat16, "input_a must be float16");
  TORCH_CHECK(input_a.dim() >= 2 && input_a.size(-1) == in_features, "input_a must be [..., in_features]");
  const c10::cuda::CUDAGuard guard(qweight.device());
  std::vector<int64_t> outputshape = {input_a.size(0), qweight.size(1)};
  int64_t mat_m = input_a.size(0);
  if (input_a.dim() > 2) {
    outputshape.insert(outputshape.begin() + 1, input_a.size(1));
    mat_m *= input_a.size(1);
  }
  if (native) {
    at::Tensor out = at::empty(outputshape, scales.options());
    auto st = at::cuda::getCurrentCUDAStream().stream();
    if (mat_m > 0) {
      const size_t ws_bytes = xbit_gemv_workspace_bytes(2, in_features, (int)qweight.size(1), bits, groupsize);
      at::Tensor ws = gemv_workspace(qweight.device(), st, ws_bytes);
      raise_if(xbit_gemv_bf16(input_a.data_ptr(), qweight.data_ptr<int32_t>(), scales.data_ptr(), qzeros.data_ptr<int32_t>(),
                              out.data_ptr(), (int)mat_m, in_features, (int)qweight.size(1), bits, groupsize, add_zero_bias,
                              qweight.size(1), ws.data_ptr(), ws_bytes, g_static_weights.load() ? XBIT_GEMV_FLAG_STATIC_WEIGHTS : 0,
                              reinterpret_cast<xbit_stream_t>(st)));
    }
    return out;
  }
  auto f16_scale = scales;
  const auto ori_dtype = scales.scalar_type();
  if (ori_dtype == torch::kBFloat16) f16_scale = scales.to(torch::kFloat16);
  at::Tensor output = at::empty(outputshape, f16_scale.options());
  auto stream = at::cuda::getCurrentCUDAStream().stream();
  if (mat_m > 0) {
    at::Tensor ws;
    void* ws_ptr = nullptr;
    const size_t ws_bytes = xbit_gemv_workspace_bytes((int)(mat_m > 16 ? 16 : mat_m), in_features, (int)qweight.size(1), bits, groupsize);
    if (ws_bytes > 0) {
      ws = gemv_workspace(qweight.device(), stream, ws_bytes);
      ws_ptr = ws.data_ptr();
    }
    raise_if(xbit_gemv_f16_ex(input_a.data_ptr(), qweight.data_ptr<int32_t>(), f16_scale.data_ptr(),
                              qzeros.data_ptr<int32_t>(), output.data_ptr(), (int)mat_m, in_features, (int)qweight.size(1),
                              bits, groupsize, add_zero_bias, qweight.size(1), ws_ptr, ws_bytes,
                              XBIT_GEMV_AUTO | (g_static_weights.load() ? XBIT_GEMV_FLAG_STATIC_WEIGHTS : 0),
                              reinterpret_cast<xbit_stream_t>(stream)));
  }
  if (ori_dtype == torch::kBFloat16) output = output.to(torch::kBFloat16);
  return output;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("dequant", &dequant_any_bit,
        "dequantize qweight to fp16, \nfunction type: const torch::Tensor& qweight, "
        "const torch::Tensor& scales, const torch::Tensor& qzeros, int groupsize, int bits, int in_features, "
        "int add_zero_bias");
  m.def("gemv", &op_gemv,
        "gemv, \nfunction type: const torch::Tensor& input_a, const torch::Tensor& qweight, "
        "const torch::Tensor& scales, const torch::Tensor& qzeros, int groupsize, int bits, int in_features, "
        "int add_zero_bias");
  m.def("set_static_weights", [](bool on) { g_static_weights.store(on); },
        "promise that the weight tensors passed to gemv are never written by the kernel preceding the call in its "
        "stream (resident model weights): gemv then prefetches them while that kernel drains");
  m.def("get_static_weights", []() { return g_static_weights.load(); });
  m.def("set_native_bf16", [](bool on) { g_native_bf16.store(on); },
        "bf16 scales (dequant) / bf16 activations and scales (gemv) without the reference's fp16 round trip");
  m.def("get_native_bf16", []() { return g_native_bf16.load(); });
}
