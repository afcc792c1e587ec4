/*
 * xbitops_b200.h -- torch-free C ABI of the B200-native (sm_100a) XbitOps hot path.
 *
 * This is the drop-in boundary.  It replaces the reference's L2 host
 * launchers, which take torch::Tensor& and therefore cannot be bound from anything but ATen:
 *
 *   xbit_dequant_f16      <-  lauch_deqantize_cuda_pt_kernel   /root/reference/src/cu/unpack_weight_2_to_7.cu:426-441
 *                             (declared at src/dq_torch_ops.cc:11-13, called at :39)
 *   xbit_gemv_f16         <-  lauch_Gemv_kernel                /root/reference/src/cu/gemv_w4a16_pt.cu:149-173
 *                             (declared at src/dq_torch_ops.cc:15-17, called at :72)
 *   xbit_gemv_f16_peers   <-  no reference counterpart (the reference is single-GPU):
 *                             N-split GEMV whose epilogue stores each finished output slice into
 *                             every rank's output buffer over NVLink peer mappings.
 *   xbit_gemv_f16_peers_signal / xbit_peers_wait  <-  none either: the same with the rank
 *                             synchronisation fused into the kernel (flag per rank, no barrier launch).
 *   xbit_gemv_f16_peers_ll / xbit_ll_unpack_f16   <-  none either: flag-in-data exchange (8-byte
 *                             {results, call number} stores), consumed by the next call's staging.
 *
 * The reference's own `extern "C" int QbitGemv(SampleData*)` (src/gemv.cuh:22) is a benchmark
 * harness entry, not an operator ABI, and is deliberately not mirrored.
 *
 * Conventions (same meaning as the reference's op arguments, src/dq_torch_ops.cc:23-24,46-48):
 *   K = in_features, N = out_features = qweight.size(1), M = activation rows, G = ceil(K/groupsize)
 *   qweight  int32 [ceil(K*bits/32), N]  LSB-first bit stream along K per column
 *   scales   fp16  [G, N]
 *   qzeros   int32 [G, ceil(N*bits/32)]  LSB-first bit stream along N per group row;
 *                                        effective zero = stored + add_zero_bias
 *   out      dequant: fp16 [K, N] row-major; gemv: fp16 [M, N] with row stride out_row_stride
 * All pointers are DEVICE pointers on the current CUDA device unless a name says "host".
 * Every entry point only enqueues work on `stream` (no allocation, no synchronisation, no
 * global state): re-entrant, thread-safe and CUDA-graph capturable.
 *
 * Error convention: 0 = XBIT_OK; otherwise a negative XBIT_E* code and a message retrievable
 * with xbit_last_error() (thread-local).  Nothing in this library calls exit()/abort()
 * (the reference does: unpack_weight_2_to_7.cu:436-440, gemv_w4a16_pt.cu:152-155,168-172).
 * There is no CPU fallback: without a CUDA device every compute entry returns XBIT_ECUDA.
 */
#ifndef XBITOPS_B200_H_
#define XBITOPS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define XBIT_API
#else
#define XBIT_API __attribute__((visibility("default")))
#endif

/* cudaStream_t without pulling in cuda_runtime.h */
typedef struct CUstream_st* xbit_stream_t;

enum {
  XBIT_OK = 0,
  XBIT_EINVAL = -1,   /* bad argument (shape, bits, groupsize, null pointer, alignment) */
  XBIT_ECUDA = -2,    /* CUDA runtime error (launch failure, no device) */
  XBIT_EWORKSPACE = -3 /* workspace too small */
};

/* GEMV kernel families; XBIT_GEMV_AUTO lets the library pick by (M, bits, groupsize, shape). */
enum {
  XBIT_GEMV_AUTO = 0,
  XBIT_GEMV_SIMT = 1,     /* W4 SIMT GEMV: LOP3 magic-number unpack, half2 FMA, one row per launch  */
  XBIT_GEMV_MMA = 2,      /* W4 tensor-core kernel: register-level unpack straight into mma.sync
                             m16n8k16 fragments, fp32 accumulation, M <= 16                          */
  XBIT_GEMV_GENERIC = 3,  /* any bits 2..8, any groupsize >= 16, any M: SIMT, fp32 accumulation      */
  /* 4: the tcgen05 / TMEM family of round 1 -- removed (measured behind the mma.sync kernels at every M) */
  XBIT_GEMV_PERSIST = 5   /* W4, M <= 8: one persistent CTA per SM, per-warp TMA rings, block-granular
                             stream-K (same tensor-core block math as XBIT_GEMV_MMA); AUTO's choice
                             wherever it applies                                                       */
};

/* Flags OR-ed into the `family` argument of xbit_gemv_f16_ex / xbit_gemv_f16_peers_ex. */
enum {
  /* The weight tensors (qweight, scales, qzeros) were completely written before the kernel that
   * precedes this call in `stream` was launched (true for resident model weights in a decode
   * loop).  The kernel may then prefetch weights while that previous kernel is still running
   * (programmatic dependent launch); activations are still read only after it has finished. */
  XBIT_GEMV_FLAG_STATIC_WEIGHTS = 0x100,
  /* xbit_gemv_f16_peers_signal only: before it reads its activations the kernel waits until
   * every rank has published as many calls as this rank has (i.e. the previous N-split call is
   * complete everywhere): the gather of call i is awaited inside call i+1, no wait launch. */
  XBIT_GEMV_FLAG_WAIT_PEERS = 0x200,
  /* xbit_gemv_f16_peers_ll only: a_f16 is not a plain fp16 matrix but the LL buffer that the
   * previous xbit_gemv_f16_peers_ll call of every rank filled ([M][K/2] 8-byte slots). */
  XBIT_GEMV_FLAG_A_IS_LL = 0x400,
  XBIT_GEMV_FAMILY_MASK = 0xFF
};

XBIT_API int xbit_version(void);                 /* major*10000 + minor*100 + patch */
XBIT_API const char* xbit_last_error(void);      /* thread-local; "" when no error */

/* Policy switches for tests and the developer tools (kernel family, decomposition, ring depth ...: XBIT_GEMV_FAMILY,
 * XBIT_GEMV_STREAMK, XBIT_W4P_FINE, XBIT_W4P_RING, XBIT_DQ_SMEM_KB, ... -- the list is in csrc/gemv_sm100.cu).
 * Each is read from the environment ONCE, when the library first needs one, and can be changed afterwards only here;
 * value INT_MIN restores the built-in policy.  Results never depend on them beyond fp32 summation order.
 * Thread-safe (atomics); XBIT_EINVAL for an unknown name. */
XBIT_API int xbit_set_option(const char* name, int value);

/* Dequantise to fp16.  bits in [2, 8]; groupsize >= 16; N even.  out is fully overwritten
 * (no pre-zeroing needed, unlike the reference's at::zeros, src/dq_torch_ops.cc:38). */
XBIT_API int xbit_dequant_f16(const int32_t* qweight, const void* scales_f16, const int32_t* qzeros,
                              void* out_f16, int K, int N, int bits, int groupsize, int add_zero_bias,
                              xbit_stream_t stream);

/* bf16-native dequantisation (no reference counterpart: the reference converts bf16 scales to
 * fp16 and the fp16 result back, /root/reference/src/dq_torch_ops.cc:33-42, losing bf16's range):
 * scales and out are bf16, out = RN_bf16((w - z) * s) with (w - z) * s exact in fp32 -- one rounding. */
XBIT_API int xbit_dequant_bf16(const int32_t* qweight, const void* scales_bf16, const int32_t* qzeros,
                               void* out_bf16, int K, int N, int bits, int groupsize, int add_zero_bias,
                               xbit_stream_t stream);

/* Bytes of OPTIONAL scratch for xbit_gemv_f16 (0 when none is useful).  With a workspace of at
 * least this size the W4 path MAY run a persistent, perfectly balanced stream-K schedule (fp32
 * partial tiles + ready flags) where that is measured to be faster (large matrices whose column
 * tiles fill the SMs badly; env XBIT_GEMV_STREAMK=0/1 forces it off/on); without one (NULL /
 * smaller) split-K is always reduced through cluster shared memory.  The scratch must be 256-byte aligned, ZERO-INITIALISED before its first use
 * (every call leaves it zeroed again), and must not be shared by calls that can run concurrently
 * (calls ordered on one stream may share it). */
XBIT_API size_t xbit_gemv_workspace_bytes(int M, int K, int N, int bits, int groupsize);
/* (two disjoint regions: the stream-K schedule's flags and partial tiles, then the persistent schedule's
 * {partial, flag} slots for up to 4 matrices per launch) */

/* y[m, n] = RN16( sum_k a[m, k] * DQ[k, n] ), fp32 accumulation.
 * a_f16 is [M, K] row-major contiguous.  out_row_stride is in elements (>= N).
 * bits in [2, 8] (the reference aborts unless bits == 4 && groupsize == 128). */
XBIT_API int xbit_gemv_f16(const void* a_f16, const int32_t* qweight, const void* scales_f16,
                           const int32_t* qzeros, void* out_f16, int M, int K, int N, int bits,
                           int groupsize, int add_zero_bias, int64_t out_row_stride,
                           void* workspace, size_t workspace_bytes, xbit_stream_t stream);

/* As xbit_gemv_f16, with an explicit kernel family (XBIT_GEMV_*) -- used by the crossover
 * sweep (BASELINE.json configs[4]) and the tests.  Returns XBIT_EINVAL if the family cannot
 * run the problem. */
XBIT_API int xbit_gemv_f16_ex(const void* a_f16, const int32_t* qweight, const void* scales_f16,
                              const int32_t* qzeros, void* out_f16, int M, int K, int N, int bits,
                              int groupsize, int add_zero_bias, int64_t out_row_stride,
                              void* workspace, size_t workspace_bytes, int family,
                              xbit_stream_t stream);

/* One weight matrix of a multi-projection call: the arguments of xbit_gemv_f16 that differ per matrix. */
typedef struct xbit_gemv_problem {
  const int32_t* qweight;    /* [ceil(K*bits/32), N] */
  const void* scales_f16;    /* [G, N]               */
  const int32_t* qzeros;     /* [G, ceil(N*bits/32)] */
  void* out_f16;             /* [M, N], row stride out_row_stride */
  int N;
  int64_t out_row_stride;
} xbit_gemv_problem;

/* bf16-native GEMV (no reference counterpart: the reference computes in fp16 whatever the scales'
 * type, /root/reference/src/dq_torch_ops.cc:65-76): activations, scales and output bf16, fp32 accumulation, ONE rounding
 * of the result to bf16 -- nothing passes through fp16, so scales and activations keep bf16's range.
 * Covers bits = 4 and 8, groupsize = 128, K % 128 = 0, N % 32 = 0 (the persistent kernel's integer block math; rows are taken
 * two per launch); XBIT_EINVAL otherwise -- the caller then converts to fp16 as the reference does.
 * `flags`: XBIT_GEMV_FLAG_STATIC_WEIGHTS or 0.  Workspace as for xbit_gemv_f16. */
XBIT_API int xbit_gemv_bf16(const void* a_bf16, const int32_t* qweight, const void* scales_bf16,
                            const int32_t* qzeros, void* out_bf16, int M, int K, int N, int bits,
                            int groupsize, int add_zero_bias, int64_t out_row_stride,
                            void* workspace, size_t workspace_bytes, int flags, xbit_stream_t stream);

/* Multi-projection GEMV: `count` (1..4) weight matrices applied to ONE activation matrix, e.g. the Q, K and V
 * projections or gate + up of a decoder layer:  out_p[m, n] = RN16( sum_k a[m, k] * DQ_p[k, n] )  for every p.
 * Semantically -- and, where the fused path applies, BIT FOR BIT -- the same as `count` xbit_gemv_f16_ex calls
 * with the same family | flags word, one per matrix (/root/reference/src/dq_torch_ops.cc:46-78 is one op call per
 * projection).  Where all matrices take the persistent W4 schedule (bits 4, group size 32/64/128, K % 128 == 0,
 * N % 32 == 0, M <= 8) they run as ONE launch: one kernel boundary, one activation staging, the weight rings
 * running ahead from one matrix into the next.  Otherwise the matrices are launched one after the other.
 * workspace: as for xbit_gemv_f16 (xbit_gemv_workspace_bytes covers 4 matrices). */
XBIT_API int xbit_gemv_f16_multi(const void* a_f16, const xbit_gemv_problem* problems, int count, int M, int K,
                                 int bits, int groupsize, int add_zero_bias, void* workspace,
                                 size_t workspace_bytes, int family, xbit_stream_t stream);

/* The family xbit_gemv_f16 would pick (XBIT_GEMV_*), for introspection and the bench log. */
XBIT_API int xbit_gemv_pick_family(int M, int K, int N, int bits, int groupsize);

/* N-split multi-GPU GEMV with a fused all-gather epilogue.  This rank owns output columns
 * [col_offset, col_offset + N_local) of an [M, N_total] result; qweight/scales/qzeros are the
 * rank-local column shards (N = N_local).  peer_out[r] (r in [0, world)) is rank r's full
 * [M, N_total] fp16 output buffer as mapped in THIS process (peer_out[rank] is the local one;
 * the others are NVLink peer mappings, e.g. from cudaIpcOpenMemHandle or torch symmetric
 * memory).  The kernel's epilogue stores this rank's slice into all `world` buffers; the caller
 * still has to synchronise ranks (a barrier) before anyone reads its buffer. */
XBIT_API int xbit_gemv_f16_peers(const void* a_f16, const int32_t* qweight, const void* scales_f16,
                                 const int32_t* qzeros, void* const* peer_out_host_array, int world,
                                 int M, int K, int N_local, int bits, int groupsize, int add_zero_bias,
                                 int64_t out_row_stride, int64_t col_offset,
                                 void* workspace, size_t workspace_bytes, xbit_stream_t stream);

/* As xbit_gemv_f16_peers with an explicit family | flags word (see xbit_gemv_f16_ex). */
XBIT_API int xbit_gemv_f16_peers_ex(const void* a_f16, const int32_t* qweight, const void* scales_f16,
                                    const int32_t* qzeros, void* const* peer_out_host_array, int world,
                                    int M, int K, int N_local, int bits, int groupsize, int add_zero_bias,
                                    int64_t out_row_stride, int64_t col_offset,
                                    void* workspace, size_t workspace_bytes, int family,
                                    xbit_stream_t stream);

/* As xbit_gemv_f16_peers_ex, with the rank synchronisation fused into the kernel: no barrier
 * launch.  peer_flags[r] is rank r's flag array (world x uint32, zero-initialised once, in the
 * same kind of peer-mapped memory as the outputs) as mapped in THIS process; local_state is two
 * zero-initialised uint32 in ordinary device memory of this rank.  When the last column tile of
 * this rank's slice has been stored into every buffer, the kernel publishes its call number
 * (local_state[1] + 1) in slot `rank` of every rank's flag array (system-scope fence, then the
 * flag stores).  A rank may read its output buffer after xbit_peers_wait on the same stream, or
 * inside the next xbit_gemv_f16_peers_signal call issued with XBIT_GEMV_FLAG_WAIT_PEERS (whose
 * activations are then typically that buffer).  Call numbers live in device memory, so a captured
 * CUDA graph can be replayed.  Every rank must issue the same sequence of calls.  Restrictions:
 * the W4 fast path (bits 4, groupsize 32/64/128, K%128 = 0, N_local%32 = 0) and M <= 16;
 * XBIT_EINVAL otherwise (use the barrier form). */
XBIT_API int xbit_gemv_f16_peers_signal(const void* a_f16, const int32_t* qweight, const void* scales_f16,
                                        const int32_t* qzeros, void* const* peer_out_host_array,
                                        void* const* peer_flags_host_array, void* local_state,
                                        int world, int rank, int M, int K, int N_local, int bits,
                                        int groupsize, int add_zero_bias, int64_t out_row_stride,
                                        int64_t col_offset, int family, xbit_stream_t stream);

/* Enqueues a one-warp kernel that returns once every rank has published as many calls as this
 * rank (local_flags = this rank's own flag array).  timeout_flag (device uint32, may be NULL) is
 * set to 1 if a peer did not arrive within about 30 s of SM clocks -- the kernel never hangs. */
XBIT_API int xbit_peers_wait(const void* local_flags, int world, int rank, void* timeout_flag,
                             xbit_stream_t stream);

/* Flag-in-data ("LL") form of the N-split exchange, for chains of dependent GEMVs (a decode
 * step): no barrier, no fence, no wait launch, no wait for the previous grid.  peer_ll_out[r] is
 * rank r's LL result buffer as mapped in THIS process: M * out_row_stride / 2 slots of 8 bytes,
 * slot = {two fp16 results, call number}, in peer-mapped memory, zero-initialised once.  The
 * epilogue stores every pair of this rank's results into every rank's buffer with one 8-byte
 * store each.  A chain is a sequence of calls with chain_index 0, 1, 2, ...: call 0 reads plain
 * fp16 activations, call i > 0 (XBIT_GEMV_FLAG_A_IS_LL) reads the LL buffer call i-1 filled -- its
 * activation staging spins on exactly the slots it needs, as they arrive from the ranks, and does
 * not wait for call i-1's grid to complete -- and xbit_ll_unpack_f16 with chain_len = number of
 * calls ends the chain.  Call numbers are local_state[2] (the chain base; four zero-initialised
 * uint32 in ordinary device memory of this rank, advanced by xbit_ll_unpack_f16; local_state[3] is
 * set to 1 if a slot did not arrive within about 30 s) + chain_index + 1,
 * so a captured CUDA graph can be replayed.  Rotate over nbuf >= 2 LL buffers along a chain (call i
 * writes buffer i % nbuf and reads buffer (i - 1) % nbuf) with (chain_len - 1) % nbuf != 0: a
 * repeated or replayed chain starts again on buffer 0 while a slower rank may still be unpacking
 * the last call's buffer, so the two must never coincide (a single-call chain alternates two
 * buffers between repetitions instead).  Every rank must issue the same sequence of calls.  Same restrictions as xbit_gemv_f16_peers_signal;
 * out_row_stride and col_offset must be even. */
XBIT_API int xbit_gemv_f16_peers_ll(const void* a_f16_or_ll, const int32_t* qweight, const void* scales_f16,
                                    const int32_t* qzeros, void* const* peer_ll_out_host_array,
                                    void* local_state, int chain_index, int world, int rank, int M, int K,
                                    int N_local, int bits, int groupsize, int add_zero_bias,
                                    int64_t out_row_stride, int64_t col_offset, int family,
                                    xbit_stream_t stream);

/* End of a chain of chain_len calls: LL buffer (filled by the last call) -> plain fp16 [n_elems]
 * once every slot carries the last call's number; then advances the chain base in local_state.
 * n_elems must be even.  timeout_flag as in xbit_peers_wait. */
XBIT_API int xbit_ll_unpack_f16(const void* ll_in, void* out_f16, int64_t n_elems, void* local_state,
                                int chain_len, void* timeout_flag, xbit_stream_t stream);

/* Host-buffer convenience used for end-to-end measurement: activations come from (pinned) host
 * memory and the result goes back to host memory; weights stay resident on the device.
 * d_a_staging / d_out_staging are device scratch of M*K*2 and M*N*2 bytes.  Enqueues
 * H2D copy -> gemv -> D2H copy on `stream`; the caller synchronises.  With page-locked,
 * device-mapped host memory (cudaHostAlloc / cudaHostRegister) no copy nodes are enqueued: a
 * small kernel pulls the activation rows over PCIe and the GEMV stores its result straight into
 * out_f16_host.  The weights must not be written by the operation that precedes this call in
 * `stream` (they are prefetched under programmatic dependent launch). */
XBIT_API int xbit_gemv_f16_host(const void* a_f16_host, void* out_f16_host, void* d_a_staging,
                                void* d_out_staging, const int32_t* qweight, const void* scales_f16,
                                const int32_t* qzeros, int M, int K, int N, int bits, int groupsize,
                                int add_zero_bias, void* workspace, size_t workspace_bytes,
                                xbit_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* XBITOPS_B200_H_ */
