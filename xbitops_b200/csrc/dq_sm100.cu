// dq_sm100.cu -- group-wise 2..8-bit dequantisation to fp16 for sm_100a.
//
// Replaces /root/reference/src/cu/unpack_weight_2_to_7.cu: DequantizeAndUnpackWeight248 (:44-85),
// iterator_qweight_v2 (:196-217), DequantizeAndUnpackWeight3567_v2 (:219-330) and their launchers
// (:353-422).  Output is bit-identical to the reference arithmetic
//     sz = hmul2(half(z + bias), s);  out = hfma2(half(w), s, -sz)       (:58-61, :72-75, :284-286, :305)
// for every width, with the bit-stream semantics the reference's b=3/5/7 path and CPU simulator
// implement (the reference's GPU b=6 path is broken beyond row 31; we follow the stream).
//
// Design (HBM-bound, write dominated: K*N*2 output bytes vs K*N*b/8 input bytes):
//   * ONE kernel template for all widths: a thread owns a 32-row x 8-column block = B word-rows x
//     two 128-bit loads; the warp reads 1 KiB contiguous per word-row and writes 512 B contiguous
//     per output row with 16-byte streaming stores (the reference: 8-byte loads, 4-byte stores,
//     and a shared-memory round trip for b=3/5/6/7).
//   * no shared memory, no run-time indexed word arrays: every field position is a compile-time
//     constant, so extraction is PRMT (byte window) + LOP3 (mask|magic) and, only where a field
//     straddles a 32-bit word or sits too high in its byte window, one SHF funnel shift.
//   * int -> fp16 is exact by construction (magic number, then HSUB2 of the base), so the final
//     HFMA2 sees exactly half(w) as the reference does.
//   * out is fully overwritten: callers allocate with empty(), not zeros() (reference: at::zeros,
//     src/dq_torch_ops.cc:38 -> an extra K*N*2-byte memset).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <type_traits>
#include <utility>

#include "unpack.cuh"
#include "xbit_internal.h"

namespace xbit {

template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(std::forward<F>(f));
  }
}

// exact fp16 pair (half(field_I of column a), half(field_I of column b)) from two columns' blocks
template <int B, int I>
__device__ __forceinline__ __half2 field_pair_exact(const uint32_t (&wa)[B], const uint32_t (&wb)[B]) {
  constexpr int pos = I * B, wi = pos >> 5, sh = pos & 31, byte = sh >> 3, p = sh & 7;
  constexpr uint32_t mask = (1u << B) - 1u;
  constexpr bool one_byte = (p + B <= 8);
  constexpr bool window_ok = (p + B <= 10) && (one_byte || byte <= 2);
  if constexpr (window_ok) {
    // bytes [byte, byte+1] of each column's word into the two 16-bit lanes; the field then sits at
    // mantissa bits [p, p+B) of each lane
    constexpr int b1 = one_byte ? byte : byte + 1;
    constexpr uint32_t sel = (uint32_t)byte | ((uint32_t)b1 << 4) | ((uint32_t)(4 + byte) << 8) | ((uint32_t)(4 + b1) << 12);
    const uint32_t lanes = prmt(wa[wi], wb[wi], sel);
    const uint32_t x = and_or(lanes, dup16(mask << p), magic2(p));
    return __hsub2(u2h2(x), u2h2(magic2(p)));          // (2^(10-p) + w) - 2^(10-p) = w, exact
  } else {
    const uint32_t va = block_field<B, I>(wa);
    const uint32_t vb = block_field<B, I>(wb);
    const uint32_t lanes = prmt(va, vb, 0x5410);       // low 16 bits of each
    const uint32_t x = and_or(lanes, dup16(mask), magic2(0));
    return __hsub2(u2h2(x), u2h2(magic2(0)));
  }
}

// Eight consecutive b-bit zero points starting at column n0 (multiple of 8) of one qzeros row:
// they occupy exactly b bytes starting at byte (n0/8)*b of the row.
template <int B>
__device__ __forceinline__ void load_zero_octet(const uint32_t* __restrict__ zrow, int zwords, int octet,
                                                int zero_bias, uint32_t (&z)[8]) {
  const int byte_off = octet * B;
  const int w0 = byte_off >> 2;
  const int boff = (byte_off & 3) * 8;
  const uint32_t a = __ldg(zrow + w0);
  const uint32_t b = (w0 + 1 < zwords) ? __ldg(zrow + w0 + 1) : 0u;
  const uint32_t c = (w0 + 2 < zwords) ? __ldg(zrow + w0 + 2) : 0u;
  unsigned long long v = ((unsigned long long)b << 32) | a;
  v >>= boff;
  if (boff) v |= (unsigned long long)c << (64 - boff);
#pragma unroll
  for (int j = 0; j < 8; ++j) z[j] = (uint32_t)((v >> (j * B)) & ((1u << B) - 1u)) + (uint32_t)zero_bias;
}

// -(half(z0), half(z1)) * s  as the reference computes it: hmul2 then negate (exact)
__device__ __forceinline__ __half2 neg_scaled_zero(uint32_t z0, uint32_t z1, __half2 s) {
  // z <= 256 < 1024: (1024 + z) - 1024 is exact
  const __half2 hz = __hsub2(u2h2(magic2(0) + (z0 | (z1 << 16))), u2h2(magic2(0)));
  return __hneg2(__hmul2(hz, s));
}

// BF: bf16-native arithmetic (the reference converts bf16 scales to fp16 and the result back,
// src/dq_torch_ops.cc:33-42, which loses bf16's range): scales and output are bf16 and
//     out = RN_bf16((w - z) * s)      -- (w - z) * s is exact in fp32 (9 x 8 significant bits): ONE rounding.
template <int B, bool BF>
__global__ void __launch_bounds__(128)
dq_block32_kernel(const uint32_t* __restrict__ qweight, const __half* __restrict__ scales,
                  const uint32_t* __restrict__ qzeros, __half* __restrict__ out,
                  int K, int N, int groupsize, int zero_bias, int qrows, int zwords, int rblocks) {
  // programmatic dependent launch: the next kernel in the stream may become resident while this one
  // drains; nothing is read before the previous kernel has completed
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int octets = N >> 3;
  const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
  const int cx = (int)(t % octets);
  const int rb = (int)(t / octets);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (rb >= rblocks) return;
  const int k0 = rb * 32;
  const int n0 = cx * 8;

  // ---- packed weights: B word-rows x 8 columns, two 128-bit streaming loads per row
  uint32_t w[8][B];
#pragma unroll
  for (int j = 0; j < B; ++j) {
    const int row = rb * B + j;
    uint4 lo = make_uint4(0, 0, 0, 0), hi = make_uint4(0, 0, 0, 0);
    if (row < qrows) {
      const uint32_t* p = qweight + (size_t)row * N + n0;
      lo = ldg_stream_v4(p);
      hi = ldg_stream_v4(p + 4);
    }
    w[0][j] = lo.x; w[1][j] = lo.y; w[2][j] = lo.z; w[3][j] = lo.w;
    w[4][j] = hi.x; w[5][j] = hi.y; w[6][j] = hi.z; w[7][j] = hi.w;
  }

  // ---- per-group scale and zero (groupsize % 32 == 0 on this path: one group per block)
  const int g = k0 / groupsize;
  const uint4 sv = __ldg(reinterpret_cast<const uint4*>(scales + (size_t)g * N + n0));
  const __half2 s[4] = {u2h2(sv.x), u2h2(sv.y), u2h2(sv.z), u2h2(sv.w)};
  uint32_t z[8];
  load_zero_octet<B>(qzeros + (size_t)g * zwords, zwords, cx, zero_bias, z);
  __half* orow = out + (size_t)k0 * N + n0;          // (bf16 elements when BF: same size)
  const int rows_left = K - k0;
  if constexpr (BF) {
    // the 16 bits of each scale are bf16: fp32 = bits << 16;  -(z * s) is exact in fp32, and so is fma(w, s, -(z * s))
    float sf[8], nzs[8];
    const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      sf[2 * p] = __uint_as_float(sw[p] << 16);
      sf[2 * p + 1] = __uint_as_float(sw[p] & 0xffff0000u);
      nzs[2 * p] = -(float)z[2 * p] * sf[2 * p];
      nzs[2 * p + 1] = -(float)z[2 * p + 1] * sf[2 * p + 1];
    }
    auto pair_bf16 = [&](__half2 wh, int p) {
      const float2 wf = __half22float2(wh);
      const __nv_bfloat162 o = __floats2bfloat162_rn(fmaf(wf.x, sf[2 * p], nzs[2 * p]), fmaf(wf.y, sf[2 * p + 1], nzs[2 * p + 1]));
      return *reinterpret_cast<const uint32_t*>(&o);
    };
    static_for<0, 32>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      uint4 r;
      r.x = pair_bf16(field_pair_exact<B, i>(w[0], w[1]), 0);
      r.y = pair_bf16(field_pair_exact<B, i>(w[2], w[3]), 1);
      r.z = pair_bf16(field_pair_exact<B, i>(w[4], w[5]), 2);
      r.w = pair_bf16(field_pair_exact<B, i>(w[6], w[7]), 3);
      if (i < rows_left) stg_stream_v4(orow + (size_t)i * N, r);
    });
  } else {
    __half2 nsz[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) nsz[p] = neg_scaled_zero(z[2 * p], z[2 * p + 1], s[p]);

    // ---- 32 output rows, one 16-byte streaming store each
    static_for<0, 32>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      uint4 r;
      r.x = h22u(__hfma2(field_pair_exact<B, i>(w[0], w[1]), s[0], nsz[0]));
      r.y = h22u(__hfma2(field_pair_exact<B, i>(w[2], w[3]), s[1], nsz[1]));
      r.z = h22u(__hfma2(field_pair_exact<B, i>(w[4], w[5]), s[2], nsz[2]));
      r.w = h22u(__hfma2(field_pair_exact<B, i>(w[6], w[7]), s[3], nsz[3]));
      if (i < rows_left) stg_stream_v4(orow + (size_t)i * N, r);
    });
  }
}

// Fallback for shapes the block kernel cannot take (N % 8 != 0, groupsize % 32 != 0, unaligned
// pointers): one thread per output element, same arithmetic, still on the GPU.
__global__ void __launch_bounds__(256)
dq_element_kernel(const uint32_t* __restrict__ qweight, const __half* __restrict__ scales,
                  const uint32_t* __restrict__ qzeros, __half* __restrict__ out,
                  int K, int N, int bits, int groupsize, int zero_bias, int qrows, int zwords, int bf16) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)K * N) return;
  const int k = (int)(idx / N), n = (int)(idx % N);
  const uint32_t mask = (1u << bits) - 1u;
  // weight field
  const int pos = k * bits, wi = pos >> 5, sh = pos & 31;
  const uint32_t lo = __ldg(qweight + (size_t)wi * N + n);
  const uint32_t hi = (sh + bits > 32 && wi + 1 < qrows) ? __ldg(qweight + (size_t)(wi + 1) * N + n) : 0u;
  const uint32_t wv = __funnelshift_r(lo, hi, sh) & mask;
  // zero field
  const int g = k / groupsize;
  const int zpos = n * bits, zi = zpos >> 5, zsh = zpos & 31;
  const uint32_t zlo = __ldg(qzeros + (size_t)g * zwords + zi);
  const uint32_t zhi = (zsh + bits > 32 && zi + 1 < zwords) ? __ldg(qzeros + (size_t)g * zwords + zi + 1) : 0u;
  const uint32_t zv = (__funnelshift_r(zlo, zhi, zsh) & mask) + (uint32_t)zero_bias;
  if (bf16) {
    const float sf = __uint_as_float((uint32_t)reinterpret_cast<const unsigned short*>(scales)[(size_t)g * N + n] << 16);
    reinterpret_cast<__nv_bfloat16*>(out)[idx] = __float2bfloat16_rn(fmaf((float)wv, sf, -(float)zv * sf));
    return;
  }
  const __half s = scales[(size_t)g * N + n];
  const __half sz = __hmul(__ushort2half_rn((unsigned short)zv), s);
  out[idx] = __hfma(__ushort2half_rn((unsigned short)wv), s, __hneg(sz));
}

// Occupancy cap of the block kernel in KiB of (unused) dynamic shared memory per 128-thread block: 48 KiB = 4 blocks per
// SM, measured best on 4096 x 11008 (profiles/r02_pdq_occupancy_cap.log: 74-80 % -> 81-86 % of the copy peak);
// option XBIT_DQ_SMEM_KB overrides it (tools/pdq.py).
static int dq_smem_cap_kb() {
  const int v = env_int("XBIT_DQ_SMEM_KB", 48);
  return (v < 0 || v > 200) ? 48 : v;
}

template <int B, bool BF>
static cudaError_t launch_block32(const DqArgs& a, cudaStream_t stream) {
  const int rblocks = (a.K + 31) / 32;
  const long long threads = (long long)rblocks * (a.N >> 3);
  const unsigned grid = (unsigned)((threads + 127) / 128);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(128, 1, 1);
  cfg.stream = stream;
  // Occupancy cap (unused dynamic shared memory): with every block resident at once (10 per SM hold a whole 4096 x 11008
  // matrix in one wave) all threads read first and all write afterwards, DRAM is busy a third of the time
  // (profiles/r02_ncu_dq_*); a few waves of blocks stagger, and the loads of one wave overlap the stores of the previous.
  // (only when the grid still makes at least two such waves; small matrices keep every block resident)
  const int cap_kb = grid >= 8u * (unsigned)device_sm_count() ? dq_smem_cap_kb() : 0;
  if (cap_kb > 48) {
    static bool done[8] = {false, false, false, false, false, false, false, false};
    if (!done[B - 1]) {
      cudaError_t e = cudaFuncSetAttribute(dq_block32_kernel<B, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return e;
      done[B - 1] = true;
    }
  }
  cfg.dynamicSmemBytes = (size_t)cap_kb * 1024;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, dq_block32_kernel<B, BF>, a.qweight, a.scales, a.qzeros, a.out, a.K, a.N, a.groupsize,
                            a.zero_bias, a.qrows, a.zwords, rblocks);
}

cudaError_t launch_dequant(const DqArgs& a, cudaStream_t stream, int* path_taken) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(a.qweight) | reinterpret_cast<uintptr_t>(a.scales) |
                         reinterpret_cast<uintptr_t>(a.out)) & 15u) == 0;
  const bool block_ok = aligned && (a.N % 8 == 0) && (a.groupsize % 32 == 0);
  if (path_taken) *path_taken = block_ok ? 1 : 0;
  if (block_ok) {
    switch (a.bits) {
#define XBIT_DQ_CASE(B_) case B_: return a.bf16 ? launch_block32<B_, true>(a, stream) : launch_block32<B_, false>(a, stream);
      XBIT_DQ_CASE(2) XBIT_DQ_CASE(3) XBIT_DQ_CASE(4) XBIT_DQ_CASE(5) XBIT_DQ_CASE(6) XBIT_DQ_CASE(7) XBIT_DQ_CASE(8)
#undef XBIT_DQ_CASE
      default: return cudaErrorInvalidValue;
    }
  }
  const long long total = (long long)a.K * a.N;
  const unsigned grid = (unsigned)((total + 255) / 256);
  dq_element_kernel<<<grid, 256, 0, stream>>>(a.qweight, a.scales, a.qzeros, a.out, a.K, a.N, a.bits,
                                              a.groupsize, a.zero_bias, a.qrows, a.zwords, a.bf16);
  return cudaGetLastError();
}

}  // namespace xbit
