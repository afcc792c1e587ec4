// xbit_internal.h -- argument blocks and launcher prototypes shared by the C ABI and the kernels.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace xbit {

constexpr int kMaxPeers = 8;

struct DqArgs {
  const uint32_t* qweight;
  const __half* scales;
  const uint32_t* qzeros;
  __half* out;
  int K, N, bits, groupsize, zero_bias;
  int qrows;   // ceil(K*bits/32)
  int zwords;  // ceil(N*bits/32)
  int bf16;    // scales and out are bf16, arithmetic out = RN_bf16((w - z) * s) (dq_sm100.cu)
};

struct GemvArgs {
  const __half* a;          // [M, K]
  const uint32_t* qweight;  // [qrows, N]
  const __half* scales;     // [G, N]
  const uint32_t* qzeros;   // [G, zwords]
  __half* out[kMaxPeers];   // world output buffers ([M, ldo] each); out[0] only when world == 1
  int world;
  int M, K, N, bits, groupsize, zero_bias;
  long long ldo;            // output row stride (elements)
  long long col_offset;     // first output column of this shard
  int qrows, zwords, groups;
  int static_weights;       // XBIT_GEMV_FLAG_STATIC_WEIGHTS: weights may be read before griddepcontrol.wait
  int bf16;                 // a, scales and out are bf16 (xbit_gemv_bf16: persistent kernel, integer block math only)
  // fused completion signal of the N-split epilogue (xbit_gemv_f16_peers_signal); null = none
  unsigned int* sig_flags[kMaxPeers];   // rank r's flag array [world] as mapped here; this rank writes slot sig_rank
  unsigned int* sig_state;              // local: [0] tiles stored so far, [1] calls published
  int sig_rank;
  int sig_wait;                         // XBIT_GEMV_FLAG_WAIT_PEERS: wait for the previous call before reading activations
  // flag-in-data ("LL") form of the N-split exchange (xbit_gemv_f16_peers_ll): every pair of fp16 results
  // travels as one 8-byte {half2, call number} store, the consumer spins on the slots it needs: no fences
  int ll_out;                           // out[p] are LL buffers ([M][ldo/2] 8-byte slots); call number = sig_state[2] + ll_chain_index + 1
  int ll_chain_index;                   // position of this call in its chain (0 = first)
  int a_is_ll;                          // a is the LL buffer the previous call filled (K/2 slots per row)
  // decomposition (filled by the planner)
  int splits;               // K splits = cluster size along grid.y
  int units_per_split;      // 128-k blocks per split
  // persistent stream-K schedule (gemv_w4_streamk_kernel)
  float* sk_partials;       // [grid][M * 128] fp32 partial tiles (workspace)
  unsigned int* sk_flags;   // [grid] "partial ready" flags, zero outside a launch (workspace)
  int sk_stages_per_tile;   // ceil((K/128) / 2)
  int sk_total_stages;      // tiles * stages_per_tile
  int sk_ring;              // pipeline depth (stages)
  int sk_act_units;         // 256-k activation chunks staged per CTA
  int sk_act_abs;           // 1: chunk index = stage inside the tile (whole K staged); 0: the CTA's own unit index
  int ring;                 // pipeline depth of gemv_w4_kernel (stages, <= 8)
  unsigned long long* trace;   // tools/trace.py only: per-CTA globaltimer stamps [cta][16], null in production
  int debug_skip;           // tools/sweep.py only: 1 = consumers skip the math (feed ceiling), results are garbage
};

cudaError_t launch_dequant(const DqArgs& a, cudaStream_t stream, int* path_taken);

// W4 fast kernels (bits == 4, groupsize % 32 == 0, K % 8 == 0, N % 8 == 0, 16-byte aligned pointers)
bool gemv_w4_supported(const GemvArgs& a);
cudaError_t launch_gemv_w4_simt(GemvArgs a, cudaStream_t stream);   // M <= 4
cudaError_t launch_gemv_w4_mma(GemvArgs a, cudaStream_t stream);    // M <= 16
bool gemv_w4_mma_has_plan(GemvArgs a);                              // false: no K split stages M * K in shared memory
// persistent, balanced stream-K variant of the two (needs workspace); false = not applicable here
bool gemv_w4_streamk_applicable(const GemvArgs& a, int family);
bool gemv_w4_prefers_streamk(const GemvArgs& a, int family);      // AUTO policy: cluster grid fills the machine badly
size_t gemv_w4_streamk_workspace_bytes(int M);
cudaError_t launch_gemv_w4_streamk(GemvArgs a, int family, void* workspace, size_t workspace_bytes, cudaStream_t stream);
// persistent per-SM schedule with per-warp TMA rings (gemv_w4p_sm100.cu): the default for M <= 8
bool gemv_w4p_applicable(const GemvArgs& a);
bool gemv_w4p_preferred(const GemvArgs& a);    // AUTO policy: measured ahead of the cluster split-K kernel here
size_t gemv_w4p_workspace_bytes(int M);
cudaError_t launch_gemv_w4p(const GemvArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t stream);
// `count` (<= 4) matrices sharing the activations in one launch; cudaErrorNotSupported = launch them one by one
cudaError_t launch_gemv_w4p_multi(const GemvArgs* gs, int count, void* workspace, size_t workspace_bytes, cudaStream_t stream);
// any bits / groupsize / M / N
cudaError_t launch_gemv_generic(GemvArgs a, cudaStream_t stream);

int device_sm_count();
int env_int(const char* name, int dflt);     // policy switch: environment at load time, xbit_set_option afterwards
bool set_option(const char* name, int value);
void apply_debug_knobs(GemvArgs& a);   // no-op unless built with -DXBIT_DEVTOOLS
constexpr size_t kMaxDynSmem = 220 * 1024;
// cached cuTensorMapEncodeTiled (a pure function of its arguments) and the one-time dynamic shared memory opt-in
cudaError_t encode_2d(CUtensorMap* map, CUtensorMapDataType dt, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t row_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle sw,
                      CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
cudaError_t ensure_max_dyn_smem(const void* kern);
cudaError_t launch_ll_unpack(const void* ll_in, void* out_f16, long long n_pairs, unsigned int* state, int chain_len, unsigned int* timeout_flag, cudaStream_t stream);
cudaError_t launch_pull_rows(const void* src_host_devptr, void* dst, size_t bytes, cudaStream_t stream);
cudaError_t launch_peers_wait(const unsigned int* flags, int world, int rank, unsigned int* timeout_flag, cudaStream_t stream);

}  // namespace xbit
