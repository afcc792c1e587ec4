// xbit_capi.cu -- the torch-free C ABI (include/xbitops_b200.h): argument validation, family
// selection and launch.  Replaces the reference's host launchers
//   lauch_deqantize_cuda_pt_kernel  /root/reference/src/cu/unpack_weight_2_to_7.cu:426-441
//   lauch_Gemv_kernel               /root/reference/src/cu/gemv_w4a16_pt.cu:149-173
// and keeps the preconditions of the op layer (/root/reference/src/dq_torch_ops.cc:25-31,49-57),
// turned into error codes instead of TORCH_CHECK / exit(-1) / abort().
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/xbitops_b200.h"
#include "xbit_internal.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  // clear the sticky-less last error so the next call starts clean
  (void)cudaGetLastError();
  return fail(XBIT_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

int check_common(const void* qweight, const void* scales, const void* qzeros, int K, int N, int bits, int groupsize,
                 int add_zero_bias) {
  if (!qweight || !scales || !qzeros) return fail(XBIT_EINVAL, "null tensor pointer");
  if (bits < 2 || bits > 8) return fail(XBIT_EINVAL, "bits must be in [2, 8], got %d", bits);
  if (groupsize < 16) return fail(XBIT_EINVAL, "groupsize must be >= 16, got %d", groupsize);
  if (K < 1 || N < 1) return fail(XBIT_EINVAL, "in_features (K=%d) and out_features (N=%d) must be >= 1", K, N);
  if ((long long)K * bits > 0x7fffffffLL || (long long)N * bits > 0x7fffffffLL)
    return fail(XBIT_EINVAL, "K*bits / N*bits overflow int32");
  if (add_zero_bias != 0 && add_zero_bias != 1) return fail(XBIT_EINVAL, "add_zero_bias must be 0 or 1, got %d", add_zero_bias);
  return XBIT_OK;
}

}  // namespace

extern "C" {

int xbit_version(void) { return 200; /* 0.2.0 */ }

int xbit_set_option(const char* name, int value) {
  g_err[0] = 0;
  if (!name || !xbit::set_option(name, value)) return fail(XBIT_EINVAL, "unknown option %s", name ? name : "(null)");
  return XBIT_OK;
}

const char* xbit_last_error(void) { return g_err; }

static int dequant_any(const int32_t* qweight, const void* scales_f16, const int32_t* qzeros, void* out_f16, int K, int N,
                       int bits, int groupsize, int add_zero_bias, xbit_stream_t stream, int bf16);

int xbit_dequant_f16(const int32_t* qweight, const void* scales_f16, const int32_t* qzeros, void* out_f16, int K, int N,
                     int bits, int groupsize, int add_zero_bias, xbit_stream_t stream) {
  return dequant_any(qweight, scales_f16, qzeros, out_f16, K, N, bits, groupsize, add_zero_bias, stream, 0);
}

int xbit_dequant_bf16(const int32_t* qweight, const void* scales_bf16, const int32_t* qzeros, void* out_bf16, int K, int N,
                      int bits, int groupsize, int add_zero_bias, xbit_stream_t stream) {
  return dequant_any(qweight, scales_bf16, qzeros, out_bf16, K, N, bits, groupsize, add_zero_bias, stream, 1);
}

static int dequant_any(const int32_t* qweight, const void* scales_f16, const int32_t* qzeros, void* out_f16, int K, int N,
                       int bits, int groupsize, int add_zero_bias, xbit_stream_t stream, int bf16) {
  g_err[0] = 0;
  if (int rc = check_common(qweight, scales_f16, qzeros, K, N, bits, groupsize, add_zero_bias)) return rc;
  if (!out_f16) return fail(XBIT_EINVAL, "null output pointer");
  if ((reinterpret_cast<uintptr_t>(out_f16) | reinterpret_cast<uintptr_t>(scales_f16)) & 1u)
    return fail(XBIT_EINVAL, "16-bit float pointers must be 2-byte aligned");
  if ((reinterpret_cast<uintptr_t>(qweight) | reinterpret_cast<uintptr_t>(qzeros)) & 3u)
    return fail(XBIT_EINVAL, "int32 pointers must be 4-byte aligned");
  xbit::DqArgs a;
  a.bf16 = bf16;
  a.qweight = reinterpret_cast<const uint32_t*>(qweight);
  a.scales = reinterpret_cast<const __half*>(scales_f16);
  a.qzeros = reinterpret_cast<const uint32_t*>(qzeros);
  a.out = reinterpret_cast<__half*>(out_f16);
  a.K = K; a.N = N; a.bits = bits; a.groupsize = groupsize; a.zero_bias = add_zero_bias;
  a.qrows = ceil_div((long long)K * bits, 32);
  a.zwords = ceil_div((long long)N * bits, 32);
  cudaError_t e = xbit::launch_dequant(a, reinterpret_cast<cudaStream_t>(stream), nullptr);
  if (e != cudaSuccess) return cuda_fail(e, bf16 ? "xbit_dequant_bf16 launch" : "xbit_dequant_f16 launch");
  return XBIT_OK;
}

static size_t persist_ws_offset(int m);

size_t xbit_gemv_workspace_bytes(int M, int, int, int bits, int) {
  // Optional: with this much zero-initialised scratch the W4 path runs the persistent, perfectly
  // balanced stream-K schedule (partial tiles + ready flags; left zeroed after every call).
  // Without it (NULL / too small) split-K is reduced through cluster shared memory instead.
  if (bits != 4 && bits != 8 && bits != 2) return 0;
  const int m = M > 16 ? 16 : (M < 1 ? 1 : M);
  // two disjoint regions: [0, sk) the stream-K kernel's flags + partial tiles, [sk, sk + pp) the persistent kernel's
  // {partial, flag} slots (the former leaves its partial tiles behind, which must never be read as slots)
  return persist_ws_offset(m) + 4 * xbit::gemv_w4p_workspace_bytes(m);
}

static size_t persist_ws_offset(int m) { return (xbit::gemv_w4_streamk_workspace_bytes(m) + 255) / 256 * 256; }

static int pick_family(const xbit::GemvArgs& a) {
  // 8-bit weights, groupsize 128, M <= 2: the persistent kernel's integer block math (the packed words are the MMA
  // operands as they are); every other width / group size outside the W4 kernels: the generic kernel
  if ((a.bits == 8 || a.bits == 2) && a.M <= 2 && xbit::env_int("XBIT_GEMV_FAMILY", 0) != XBIT_GEMV_GENERIC && xbit::gemv_w4p_preferred(a))
    return XBIT_GEMV_PERSIST;
  if (!xbit::gemv_w4_supported(a)) return XBIT_GEMV_GENERIC;
  // Crossover measured on B200 (BASELINE.json configs[4]; profiles/): with the nibble
  // bits fed to the tensor core as fp16 subnormals the mma.sync kernel needs one ALU op per weight
  // pair and accumulates in fp32, so it is at least as fast as the SIMT half2-FMA kernel already
  // at M = 1 and strictly more accurate.  SIMT stays selectable (family / XBIT_GEMV_FAMILY).
  const int forced = xbit::env_int("XBIT_GEMV_FAMILY", 0);
  if (forced == XBIT_GEMV_SIMT && a.M == 1) return XBIT_GEMV_SIMT;
  if (forced == XBIT_GEMV_MMA) return XBIT_GEMV_MMA;
  if (forced == XBIT_GEMV_GENERIC) return XBIT_GEMV_GENERIC;
  if (a.M <= 8 && xbit::gemv_w4p_preferred(a)) return XBIT_GEMV_PERSIST;
  return XBIT_GEMV_MMA;
}

static bool use_streamk(const xbit::GemvArgs& g, int family, void* workspace, size_t workspace_bytes) {
  // The persistent stream-K schedule needs the caller's workspace.  XBIT_GEMV_STREAMK=1 forces it on
  // wherever it applies, =0 off; otherwise the measured policy of gemv_w4_prefers_streamk decides.
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return false;
  if (workspace_bytes < xbit::gemv_w4_streamk_workspace_bytes(g.M)) return false;
  const int v = xbit::env_int("XBIT_GEMV_STREAMK", -1);
  if (v == 0) return false;
  if (v == 1) return xbit::gemv_w4_streamk_applicable(g, family);
  return xbit::gemv_w4_prefers_streamk(g, family);
}

struct PeerSignal {
  void* const* flags;   // world peer-mapped flag arrays (null in the LL form)
  void* state;          // local uint32 counters
  int rank;
  bool ll;              // flag-in-data form: outs are LL buffers
  int chain_index;      // LL form: position of the call in its chain
};

static int gemv_impl(const void* a_f16, const int32_t* qweight, const void* scales_f16, const int32_t* qzeros,
                     void* const* outs, int world, int M, int K, int N, int bits, int groupsize, int add_zero_bias,
                     int64_t out_row_stride, int64_t col_offset, int family_and_flags, void* workspace,
                     size_t workspace_bytes, xbit_stream_t stream, const PeerSignal* sig = nullptr) {
  g_err[0] = 0;
  if (int rc = check_common(qweight, scales_f16, qzeros, K, N, bits, groupsize, add_zero_bias)) return rc;
  if (!a_f16) return fail(XBIT_EINVAL, "null activation pointer");
  if (M < 1) return fail(XBIT_EINVAL, "M must be >= 1, got %d", M);
  if (world < 1 || world > xbit::kMaxPeers) return fail(XBIT_EINVAL, "world must be in [1, %d], got %d", xbit::kMaxPeers, world);
  if (col_offset < 0 || out_row_stride < col_offset + N)
    return fail(XBIT_EINVAL, "out_row_stride (%lld) must be >= col_offset + N (%lld)", (long long)out_row_stride,
                (long long)(col_offset + N));
  xbit::GemvArgs g = {};
  memset(&g, 0, sizeof(g));
  for (int p = 0; p < world; ++p) {
    if (!outs || !outs[p]) return fail(XBIT_EINVAL, "null output pointer (rank %d)", p);
    g.out[p] = reinterpret_cast<__half*>(outs[p]);
  }
  g.world = world;
  g.qweight = reinterpret_cast<const uint32_t*>(qweight);
  g.scales = reinterpret_cast<const __half*>(scales_f16);
  g.qzeros = reinterpret_cast<const uint32_t*>(qzeros);
  g.K = K; g.N = N; g.bits = bits; g.groupsize = groupsize; g.zero_bias = add_zero_bias;
  g.ldo = out_row_stride; g.col_offset = col_offset;
  g.qrows = ceil_div((long long)K * bits, 32);
  g.zwords = ceil_div((long long)N * bits, 32);
  g.groups = ceil_div(K, groupsize);
  g.static_weights = (family_and_flags & XBIT_GEMV_FLAG_STATIC_WEIGHTS) ? 1 : 0;
  int family_req = family_and_flags & XBIT_GEMV_FAMILY_MASK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (sig) {
    // the signal is raised by the cluster split-K kernel of the tensor-core family, one launch per call
    if ((!sig->ll && !sig->flags) || !sig->state || sig->rank < 0 || sig->rank >= world) return fail(XBIT_EINVAL, "bad peer signal arguments");
    if (sig->ll && ((out_row_stride | col_offset) & 1)) return fail(XBIT_EINVAL, "the LL form needs even out_row_stride and col_offset");
    if (family_req != XBIT_GEMV_AUTO && family_req != XBIT_GEMV_MMA) return fail(XBIT_EINVAL, "the fused signal needs the AUTO or MMA family");
    g.M = M;
    if (M > 16 || !xbit::gemv_w4_supported(g))
      return fail(XBIT_EINVAL, "the fused signal needs the W4 fast path (bits 4, groupsize 32/64/128, K%%128=0, N%%32=0) and M <= 16");
    // The flag-in-data form is also spoken by the persistent kernel (option XBIT_LL_PERSIST=1; tile-aligned CTA ranges:
    // consecutive calls of a chain overlap, so they must not share workspace slots), but measured behind the cluster kernel
    // there: 6.02 vs 7.85 TB/s aggregate on 4 GPUs, 6.42 vs 6.49 on 2 (profiles/r02_bench_n4_ll_*.json) -- with one CTA per
    // SM the next call of a chain cannot spin on its slots while this one computes.  The completion-flag form is raised by
    // the cluster kernel only.
    family_req = (sig->ll && M <= 8 && xbit::env_int("XBIT_LL_PERSIST", 0) != 0 && pick_family(g) == XBIT_GEMV_PERSIST) ? XBIT_GEMV_PERSIST : XBIT_GEMV_MMA;
    workspace = nullptr;          // no stream-K here: its tiles are not stored by one CTA each
    workspace_bytes = 0;
    for (int p = 0; p < world && !sig->ll; ++p) {
      if (!sig->flags[p]) return fail(XBIT_EINVAL, "null flag pointer (rank %d)", p);
      g.sig_flags[p] = reinterpret_cast<unsigned int*>(sig->flags[p]);
    }
    g.sig_state = reinterpret_cast<unsigned int*>(sig->state);
    g.sig_rank = sig->rank;
    g.sig_wait = (!sig->ll && (family_and_flags & XBIT_GEMV_FLAG_WAIT_PEERS)) ? 1 : 0;
    g.ll_out = sig->ll ? 1 : 0;
    g.ll_chain_index = sig->chain_index;
    g.a_is_ll = (sig->ll && (family_and_flags & XBIT_GEMV_FLAG_A_IS_LL)) ? 1 : 0;
  }

  // rows are processed in slabs the chosen family can take (weights are re-read per slab only
  // beyond M = 16; the reference re-reads them for every row, gemv_w4a16_pt.cu:158)
  for (int m0 = 0; m0 < M;) {
    g.a = reinterpret_cast<const __half*>(a_f16) + (size_t)m0 * K;
    g.M = M - m0;
    xbit::GemvArgs probe = g;
    int family = family_req;
    int auto_slab = 0;
    if (family == XBIT_GEMV_AUTO) {
      // the largest slab of rows (16, 8, 4, 2, 1) one of the W4 kernels can stage; the generic kernel otherwise
      // (e.g. M = 16 with K = 32768 has no K split that fits the activations in shared memory)
      family = XBIT_GEMV_GENERIC;
      for (int slab_try = g.M > 16 ? 16 : g.M; slab_try >= 1; slab_try = slab_try > 1 ? (slab_try + 1) / 2 : 0) {
        probe.M = slab_try;
        const int f = pick_family(probe);
        if (f == XBIT_GEMV_GENERIC) {
          if ((g.bits == 8 || g.bits == 2) && slab_try > 2) continue;   // 2- / 8-bit weights: the persistent kernel takes two rows per launch
          break;
        }
        if (f != XBIT_GEMV_MMA || xbit::gemv_w4_mma_has_plan(probe) || use_streamk(probe, XBIT_GEMV_MMA, workspace, workspace_bytes)) {
          family = f;
          auto_slab = slab_try;
          break;
        }
      }
    }
    int slab;
    cudaError_t e;
    switch (family) {
      case XBIT_GEMV_SIMT:
        if (!xbit::gemv_w4_supported(g)) return fail(XBIT_EINVAL, "SIMT family needs bits=4, groupsize in {32, 64, 128}, K%%128=0, N%%32=0, 16-byte aligned pointers");
        slab = 1; g.M = 1;   // one activation row per launch (the GEMV kernel proper)
        if (use_streamk(g, XBIT_GEMV_SIMT, workspace, workspace_bytes)) e = xbit::launch_gemv_w4_streamk(g, XBIT_GEMV_SIMT, workspace, workspace_bytes, st);
        else e = xbit::launch_gemv_w4_simt(g, st);
        break;
      case XBIT_GEMV_MMA:
        if (!xbit::gemv_w4_supported(g)) return fail(XBIT_EINVAL, "MMA family needs bits=4, groupsize in {32, 64, 128}, K%%128=0, N%%32=0, 16-byte aligned pointers");
        slab = auto_slab ? auto_slab : (g.M > 16 ? 16 : g.M); g.M = slab;
        if (use_streamk(g, XBIT_GEMV_MMA, workspace, workspace_bytes)) e = xbit::launch_gemv_w4_streamk(g, XBIT_GEMV_MMA, workspace, workspace_bytes, st);
        else e = xbit::launch_gemv_w4_mma(g, st);
        break;
      case XBIT_GEMV_PERSIST:
        slab = auto_slab ? auto_slab : (g.M > 8 ? 8 : g.M);
        if ((g.bits == 8 || g.bits == 2) && slab > 2) slab = 2;
        g.M = slab;
        if (!xbit::gemv_w4p_applicable(g)) return fail(XBIT_EINVAL, "PERSIST family needs bits=4 (groupsize in {32, 64, 128}) or bits=8 (groupsize 128), K%%128=0, N%%32=0, 16-byte aligned pointers and M*K small enough to stage");
        {
          // the persistent kernel's region of the workspace lies behind the stream-K kernel's (xbit_gemv_workspace_bytes)
          const size_t off = persist_ws_offset(g.M);
          const bool has = workspace && workspace_bytes > off;
          e = xbit::launch_gemv_w4p(g, has ? static_cast<unsigned char*>(workspace) + off : nullptr, has ? workspace_bytes - off : 0, st);
        }
        break;
      case XBIT_GEMV_GENERIC:
        slab = g.M;
        e = xbit::launch_gemv_generic(g, st);
        break;
      default:
        return fail(XBIT_EINVAL, "unknown gemv family %d", family);
    }
    if (e != cudaSuccess) return cuda_fail(e, "xbit_gemv_f16 launch");
    for (int p = 0; p < world; ++p) g.out[p] += (size_t)slab * out_row_stride;
    m0 += slab;
  }
  return XBIT_OK;
}

int xbit_gemv_f16_ex(const void* a_f16, const int32_t* qweight, const void* scales_f16, const int32_t* qzeros,
                     void* out_f16, int M, int K, int N, int bits, int groupsize, int add_zero_bias,
                     int64_t out_row_stride, void* workspace, size_t workspace_bytes, int family, xbit_stream_t stream) {
  void* outs[1] = {out_f16};
  return gemv_impl(a_f16, qweight, scales_f16, qzeros, outs, 1, M, K, N, bits, groupsize, add_zero_bias, out_row_stride, 0,
                   family, workspace, workspace_bytes, stream);
}

int xbit_gemv_f16(const void* a_f16, const int32_t* qweight, const void* scales_f16, const int32_t* qzeros, void* out_f16,
                  int M, int K, int N, int bits, int groupsize, int add_zero_bias, int64_t out_row_stride, void* workspace,
                  size_t workspace_bytes, xbit_stream_t stream) {
  return xbit_gemv_f16_ex(a_f16, qweight, scales_f16, qzeros, out_f16, M, K, N, bits, groupsize, add_zero_bias,
                          out_row_stride, workspace, workspace_bytes, XBIT_GEMV_AUTO, stream);
}

int xbit_gemv_bf16(const void* a_bf16, const int32_t* qweight, const void* scales_bf16, const int32_t* qzeros, void* out_bf16,
                   int M, int K, int N, int bits, int groupsize, int add_zero_bias, int64_t out_row_stride, void* workspace,
                   size_t workspace_bytes, int flags, xbit_stream_t stream) {
  g_err[0] = 0;
  if (int rc = check_common(qweight, scales_bf16, qzeros, K, N, bits, groupsize, add_zero_bias)) return rc;
  if (!a_bf16 || !out_bf16 || M < 1 || out_row_stride < N) return fail(XBIT_EINVAL, "bad activation / output arguments");
  if ((bits != 4 && bits != 8 && bits != 2) || groupsize != 128)
    return fail(XBIT_EINVAL, "the bf16-native GEMV covers bits 2, 4 and 8, groupsize=128 (got %d, %d): convert to fp16 as the reference does", bits, groupsize);
  const size_t off = persist_ws_offset(2);
  const bool has = workspace && workspace_bytes > off;
  for (int m0 = 0; m0 < M; m0 += 2) {                 // the integer block math takes two activation rows per launch
    xbit::GemvArgs g;
    memset(&g, 0, sizeof(g));
    g.bf16 = 1;
    g.a = reinterpret_cast<const __half*>(a_bf16) + (size_t)m0 * K;
    g.out[0] = reinterpret_cast<__half*>(out_bf16) + (size_t)m0 * out_row_stride;
    g.world = 1;
    g.qweight = reinterpret_cast<const uint32_t*>(qweight);
    g.scales = reinterpret_cast<const __half*>(scales_bf16);
    g.qzeros = reinterpret_cast<const uint32_t*>(qzeros);
    g.M = M - m0 < 2 ? M - m0 : 2; g.K = K; g.N = N; g.bits = bits; g.groupsize = groupsize; g.zero_bias = add_zero_bias;
    g.ldo = out_row_stride; g.col_offset = 0;
    g.qrows = ceil_div((long long)K * bits, 32);
    g.zwords = ceil_div((long long)N * bits, 32);
    g.groups = ceil_div(K, groupsize);
    g.static_weights = (flags & XBIT_GEMV_FLAG_STATIC_WEIGHTS) ? 1 : 0;
    if (!xbit::gemv_w4p_applicable(g))
      return fail(XBIT_EINVAL, "the bf16-native GEMV needs K%%128=0, N%%32=0, 16-byte aligned pointers and K small enough to stage");
    const cudaError_t e = xbit::launch_gemv_w4p(g, has ? static_cast<unsigned char*>(workspace) + off : nullptr, has ? workspace_bytes - off : 0,
                                                reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "xbit_gemv_bf16 launch");
  }
  return XBIT_OK;
}

int xbit_gemv_f16_multi(const void* a_f16, const xbit_gemv_problem* problems, int count, int M, int K, int bits, int groupsize,
                        int add_zero_bias, void* workspace, size_t workspace_bytes, int family, xbit_stream_t stream) {
  g_err[0] = 0;
  if (!problems || count < 1 || count > 4) return fail(XBIT_EINVAL, "count must be in [1, 4] with a non-null problem array, got %d", count);
  const int family_req = family & XBIT_GEMV_FAMILY_MASK;
  // the fused launch: every matrix on the persistent schedule (AUTO would pick it, or it was asked for)
  bool fuse = count > 1 && M >= 1 && M <= 8 && a_f16 && (family_req == XBIT_GEMV_AUTO || family_req == XBIT_GEMV_PERSIST);
  xbit::GemvArgs gs[4];
  for (int i = 0; i < count && fuse; ++i) {
    const xbit_gemv_problem& pr = problems[i];
    if (check_common(pr.qweight, pr.scales_f16, pr.qzeros, K, pr.N, bits, groupsize, add_zero_bias) != XBIT_OK || !pr.out_f16 ||
        pr.out_row_stride < pr.N) {
      fuse = false;                                  // let the per-matrix call report the error
      break;
    }
    xbit::GemvArgs& g = gs[i];
    memset(&g, 0, sizeof(g));
    g.a = reinterpret_cast<const __half*>(a_f16);
    g.out[0] = reinterpret_cast<__half*>(pr.out_f16);
    g.world = 1;
    g.qweight = reinterpret_cast<const uint32_t*>(pr.qweight);
    g.scales = reinterpret_cast<const __half*>(pr.scales_f16);
    g.qzeros = reinterpret_cast<const uint32_t*>(pr.qzeros);
    g.M = M; g.K = K; g.N = pr.N; g.bits = bits; g.groupsize = groupsize; g.zero_bias = add_zero_bias;
    g.ldo = pr.out_row_stride; g.col_offset = 0;
    g.qrows = ceil_div((long long)K * bits, 32);
    g.zwords = ceil_div((long long)pr.N * bits, 32);
    g.groups = ceil_div(K, groupsize);
    g.static_weights = (family & XBIT_GEMV_FLAG_STATIC_WEIGHTS) ? 1 : 0;
    if (family_req == XBIT_GEMV_AUTO ? pick_family(g) != XBIT_GEMV_PERSIST : !xbit::gemv_w4p_applicable(g)) fuse = false;
  }
  if (fuse) {
    const size_t off = persist_ws_offset(M);
    const bool has = workspace && workspace_bytes > off;
    const cudaError_t e = xbit::launch_gemv_w4p_multi(gs, count, has ? static_cast<unsigned char*>(workspace) + off : nullptr,
                                                      has ? workspace_bytes - off : 0, reinterpret_cast<cudaStream_t>(stream));
    if (e == cudaSuccess) return XBIT_OK;
    (void)cudaGetLastError();
    if (e != cudaErrorNotSupported && e != cudaErrorInvalidValue) return cuda_fail(e, "xbit_gemv_f16_multi launch");
  }
  for (int i = 0; i < count; ++i) {
    const xbit_gemv_problem& pr = problems[i];
    if (int rc = xbit_gemv_f16_ex(a_f16, pr.qweight, pr.scales_f16, pr.qzeros, pr.out_f16, M, K, pr.N, bits, groupsize, add_zero_bias,
                                  pr.out_row_stride, workspace, workspace_bytes, family, stream))
      return rc;
  }
  return XBIT_OK;
}

int xbit_gemv_pick_family(int M, int K, int N, int bits, int groupsize) {
  xbit::GemvArgs g = {};
  memset(&g, 0, sizeof(g));
  g.M = M > 16 ? 16 : M; g.K = K; g.N = N; g.bits = bits; g.groupsize = groupsize;
  return pick_family(g);
}

int xbit_gemv_f16_peers_ex(const void* a_f16, const int32_t* qweight, const void* scales_f16, const int32_t* qzeros,
                           void* const* peer_out_host_array, int world, int M, int K, int N_local, int bits, int groupsize,
                           int add_zero_bias, int64_t out_row_stride, int64_t col_offset, void* workspace,
                           size_t workspace_bytes, int family, xbit_stream_t stream) {
  return gemv_impl(a_f16, qweight, scales_f16, qzeros, peer_out_host_array, world, M, K, N_local, bits, groupsize,
                   add_zero_bias, out_row_stride, col_offset, family, workspace, workspace_bytes, stream);
}

int xbit_gemv_f16_peers(const void* a_f16, const int32_t* qweight, const void* scales_f16, const int32_t* qzeros,
                        void* const* peer_out_host_array, int world, int M, int K, int N_local, int bits, int groupsize,
                        int add_zero_bias, int64_t out_row_stride, int64_t col_offset, void* workspace,
                        size_t workspace_bytes, xbit_stream_t stream) {
  return xbit_gemv_f16_peers_ex(a_f16, qweight, scales_f16, qzeros, peer_out_host_array, world, M, K, N_local, bits,
                                groupsize, add_zero_bias, out_row_stride, col_offset, workspace, workspace_bytes,
                                XBIT_GEMV_AUTO, stream);
}

int xbit_gemv_f16_peers_signal(const void* a_f16, const int32_t* qweight, const void* scales_f16, const int32_t* qzeros,
                               void* const* peer_out_host_array, void* const* peer_flags_host_array, void* local_state,
                               int world, int rank, int M, int K, int N_local, int bits, int groupsize, int add_zero_bias,
                               int64_t out_row_stride, int64_t col_offset, int family, xbit_stream_t stream) {
  const PeerSignal sig = {peer_flags_host_array, local_state, rank, false, 0};
  return gemv_impl(a_f16, qweight, scales_f16, qzeros, peer_out_host_array, world, M, K, N_local, bits, groupsize,
                   add_zero_bias, out_row_stride, col_offset, family, nullptr, 0, stream, &sig);
}

int xbit_gemv_f16_peers_ll(const void* a_f16_or_ll, const int32_t* qweight, const void* scales_f16, const int32_t* qzeros,
                           void* const* peer_ll_out_host_array, void* local_state, int chain_index, int world, int rank, int M,
                           int K, int N_local, int bits, int groupsize, int add_zero_bias, int64_t out_row_stride,
                           int64_t col_offset, int family, xbit_stream_t stream) {
  if (chain_index < 0) return fail(XBIT_EINVAL, "chain_index must be >= 0, got %d", chain_index);
  if (((family & XBIT_GEMV_FLAG_A_IS_LL) != 0) != (chain_index > 0))
    return fail(XBIT_EINVAL, "XBIT_GEMV_FLAG_A_IS_LL must be set exactly for chain_index > 0 (got index %d)", chain_index);
  const PeerSignal sig = {nullptr, local_state, rank, true, chain_index};
  return gemv_impl(a_f16_or_ll, qweight, scales_f16, qzeros, peer_ll_out_host_array, world, M, K, N_local, bits, groupsize,
                   add_zero_bias, out_row_stride, col_offset, family, nullptr, 0, stream, &sig);
}

int xbit_ll_unpack_f16(const void* ll_in, void* out_f16, int64_t n_elems, void* local_state, int chain_len, void* timeout_flag,
                       xbit_stream_t stream) {
  g_err[0] = 0;
  if (!ll_in || !out_f16 || !local_state) return fail(XBIT_EINVAL, "null pointer");
  if (n_elems < 2 || (n_elems & 1)) return fail(XBIT_EINVAL, "n_elems must be even and >= 2, got %lld", (long long)n_elems);
  if (chain_len < 1) return fail(XBIT_EINVAL, "chain_len must be >= 1, got %d", chain_len);
  const cudaError_t e = xbit::launch_ll_unpack(ll_in, out_f16, n_elems / 2, reinterpret_cast<unsigned int*>(local_state), chain_len,
                                               reinterpret_cast<unsigned int*>(timeout_flag), reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "xbit_ll_unpack_f16 launch");
  return XBIT_OK;
}

int xbit_peers_wait(const void* local_flags, int world, int rank, void* timeout_flag, xbit_stream_t stream) {
  g_err[0] = 0;
  if (!local_flags) return fail(XBIT_EINVAL, "null flag pointer");
  if (world < 1 || world > xbit::kMaxPeers || rank < 0 || rank >= world)
    return fail(XBIT_EINVAL, "need 1 <= world <= %d and 0 <= rank < world, got world %d rank %d", xbit::kMaxPeers, world, rank);
  const cudaError_t e = xbit::launch_peers_wait(reinterpret_cast<const unsigned int*>(local_flags), world, rank,
                                                reinterpret_cast<unsigned int*>(timeout_flag), reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "xbit_peers_wait launch");
  return XBIT_OK;
}

int xbit_gemv_f16_host(const void* a_f16_host, void* out_f16_host, void* d_a_staging, void* d_out_staging,
                       const int32_t* qweight, const void* scales_f16, const int32_t* qzeros, int M, int K, int N, int bits,
                       int groupsize, int add_zero_bias, void* workspace, size_t workspace_bytes, xbit_stream_t stream) {
  g_err[0] = 0;
  if (!a_f16_host || !out_f16_host || !d_a_staging || !d_out_staging) return fail(XBIT_EINVAL, "null host/staging pointer");
  if (M < 1 || K < 1 || N < 1) return fail(XBIT_EINVAL, "bad shape M=%d K=%d N=%d", M, K, N);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e;
  const size_t a_bytes = (size_t)M * K * 2;
  const void* a_dev = nullptr;
  {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, a_f16_host) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
      a_dev = attr.devicePointer;
    else
      cudaGetLastError();
  }
  if (a_dev && a_bytes % 16 == 0 && ((reinterpret_cast<uintptr_t>(a_dev) | reinterpret_cast<uintptr_t>(d_a_staging)) & 15u) == 0) {
    // page-locked activations: pulled over PCIe by a small kernel (keeps the chain on programmatic launches)
    e = xbit::launch_pull_rows(a_dev, d_a_staging, a_bytes, st);
    if (e != cudaSuccess) return cuda_fail(e, "xbit_gemv_f16_host pull");
  } else {
    e = cudaMemcpyAsync(d_a_staging, a_f16_host, a_bytes, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "xbit_gemv_f16_host H2D");
  }
  // Pinned (page-locked, device-mapped) result buffer: the kernel's epilogue stores straight into it over
  // PCIe -- M*N*2 bytes of posted writes, no copy node behind the kernel.  Pageable memory: staged + D2H copy.
  void* out_dev = nullptr;
  {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, out_f16_host) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
      out_dev = attr.devicePointer;
    else
      cudaGetLastError();                           // unregistered host memory is reported as an error by older runtimes
  }
  // neither the copy nor the pull kernel touches the weights: the static-weights prefetch is safe
  int rc = xbit_gemv_f16_ex(d_a_staging, qweight, scales_f16, qzeros, out_dev ? out_dev : d_out_staging, M, K, N, bits, groupsize,
                            add_zero_bias, N, workspace, workspace_bytes, XBIT_GEMV_AUTO | XBIT_GEMV_FLAG_STATIC_WEIGHTS, stream);
  if (rc != XBIT_OK) return rc;
  if (!out_dev) {
    e = cudaMemcpyAsync(out_f16_host, d_out_staging, (size_t)M * N * 2, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return cuda_fail(e, "xbit_gemv_f16_host D2H");
  }
  return XBIT_OK;
}

}  // extern "C"
