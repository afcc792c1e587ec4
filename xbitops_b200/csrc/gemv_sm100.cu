// gemv_sm100.cu -- fused dequant + A16Wx GEMV / skinny GEMM for sm_100a.
//
// Replaces /root/reference/src/cu/gemv_w4a16_pt.cu: gemv<T> (:35-145), warpReduceSum (:20-33) and
// lauch_Gemv_kernel (:149-173).  Semantics: y[m, n] = RN16( sum_k a[m, k] * DQ[k, n] ) where DQ is
// the dequantised weight of dq_sm100.cu.
//
// What the reference does and why it cannot approach the B200 roofline:
// CTA = 64 columns x all of K, so N = 4096 gives 64 CTAs on 148 SMs; 8 bytes of weights in flight
// per thread; one I2F per weight; M > 1 re-reads the weights M times; legacy default stream.
//
// Kernels in this file (DESIGN.md 4.2 has the measurements behind every choice):
//
//   gemv_w4_kernel<MT, UPG, WC, HYB>   the default W4 path (bits == 4, groupsize 32 / 64 / 128): cluster split-K.
//       A producer thread streams [rows x 128 B] boxes of packed weights, plus the matching scale
//       and zero rows, with 2-D tiled TMA loads (cp.async.bulk.tensor, 128-byte swizzle) through a
//       3..5-stage shared-memory ring guarded by full/empty mbarriers; 8 consumer warps unpack
//       straight out of shared memory with one conflict-free LDS.128 per 32-k unit.
//       MT >= 1  tensor-core family, the default for every M (M <= 8*MT per launch), block math
//                w4_consume_block_v2: the masked nibble / byte bits ARE fp16 subnormals (w * 2^-24), so
//                2 PRMT + 2 LOP3 per packed word produce the m16n8k16 A fragments; products are exact,
//                accumulation is fp32 in the tensor core; the zero point rides on one extra MMA per group.
//                Weights are read once for all M rows (the reference re-reads them M times, :158).
//       MT == 0  SIMT family (M == 1, selectable): LOP3 (mask|magic) + HSUB2 -> exact fp16 integers,
//                half2 FMA chains flushed to fp32 every 64 k, per-group fp32 scale.
//       Split-K: lanes -> warps (shared memory) -> thread-block cluster (DSMEM); deterministic.
//       Epilogue variants: plain fp16 stores (also into peer GPUs' buffers), + per-rank completion
//       flag (xbit_gemv_f16_peers_signal), or flag-in-data 8-byte {results, call number} stores
//       (xbit_gemv_f16_peers_ll) whose consumer is the next call's activation staging.
//   gemv_w4_streamk_kernel<MT, UPG>    persistent stream-K schedule of the same block math: one CTA per SM on
//       half an SM, contiguous unit ranges, fp32 partial tiles + flags in the caller's workspace.
//   gemv_generic_kernel                any bits 2..8, any groupsize >= 16, any M, any N: one column per
//       thread, bit-reader over the LSB-first stream, fp32 math with the zero point folded per
//       group.  Correctness path for the combinations the reference aborts on (:152-155).
//   peers_wait_kernel, ll_unpack_kernel, pull_rows_kernel   one-warp / small helpers of the multi-GPU and
//       host-buffer entry points.
//
// Programmatic dependent launch: weights do not depend on the previous kernel in a decode step, so
// (with XBIT_GEMV_FLAG_STATIC_WEIGHTS) the weight stream starts BEFORE griddepcontrol.wait and
// only the activation staging waits for the previous kernel.
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <limits.h>
#include <math.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <utility>

#include "unpack.cuh"
#include "gemv_prims.cuh"
#include "xbit_internal.h"
#include "../../include/xbitops_b200.h"

namespace cg = cooperative_groups;

namespace xbit {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;


// One packed W4 word (8 consecutive k of one column) -> four half2 of EXACT (w - z):
//   e[0] = (k0, k4)  e[1] = (k1, k5)  e[2] = (k2, k6)  e[3] = (k3, k7)
// zc_lo = half2(1024 + z), zc_hi = half2(64 + z) as bit patterns.
__device__ __forceinline__ void unpack_w4_minus_zero(uint32_t w, uint32_t zc_lo, uint32_t zc_hi, uint32_t (&e)[4]) {
  const uint32_t w8 = w >> 8;
  e[0] = h22u(__hsub2(u2h2(and_or(w, 0x000F000Fu, magic2(0))), u2h2(zc_lo)));
  e[1] = h22u(__hsub2(u2h2(and_or(w, 0x00F000F0u, magic2(4))), u2h2(zc_hi)));
  e[2] = h22u(__hsub2(u2h2(and_or(w8, 0x000F000Fu, magic2(0))), u2h2(zc_lo)));
  e[3] = h22u(__hsub2(u2h2(and_or(w8, 0x00F000F0u, magic2(4))), u2h2(zc_hi)));
}

// Same four pairs as raw fp16 SUBNORMALS: a nibble at mantissa bits [0,4) reads as w * 2^-24, at
// bits [4,8) as w * 2^-20.  No arithmetic at all: one AND per half2.  (Tensor-core path only: the
// products with fp16 activations are exact and are accumulated in fp32.)
__device__ __forceinline__ void unpack_w4_subnormal(uint32_t w, uint32_t (&e)[4]) {
  // w >> 8 as a multiply-high: IMAD.HI runs on the FMA pipe, which idles here, while the ALU pipe
  // (LOP3 / SHF) is the kernel's scarcest resource (profiles/r01_v3_ncu_gemv_mma_8192x28672.txt)
  const uint32_t w8 = __umulhi(w, 0x01000000u);
  e[0] = w & 0x000F000Fu;
  e[1] = w & 0x00F000F0u;
  e[2] = w8 & 0x000F000Fu;
  e[3] = w8 & 0x00F000F0u;
}

// acc += e.lo * a.lo + e.hi * a.hi with fp16 inputs and an fp32 accumulator: the sm_100 mixed-precision
// FMA (PTX fma.rn.f32.f16, SASS FHFMA with .H0/.H1 operand selectors).  Products of an fp16 activation
// with a subnormal nibble are exact in fp32.
__device__ __forceinline__ float fhfma_pair(uint32_t e, uint32_t a, float c) {
  float d;
  asm("{\n .reg .f16 el, eh, al, ah;\n .reg .f32 t;\n mov.b32 {el, eh}, %1;\n mov.b32 {al, ah}, %2;\n"
      " fma.rn.f32.f16 t, el, al, %3;\n fma.rn.f32.f16 %0, eh, ah, t;\n}"
      : "=f"(d) : "r"(e), "r"(a), "f"(c));
  return d;
}

// activations [8 consecutive k] (pairs (0,1)(2,3)(4,5)(6,7)) -> pairs (0,4)(1,5)(2,6)(3,7);
// kScaleOdd: the odd-k pairs (which meet high nibbles) are pre-multiplied by 2^-4.
template <bool kScaleOdd>
__device__ __forceinline__ uint4 permute_act8(uint4 v) {
  uint4 o;
  o.x = prmt(v.x, v.z, 0x5410);
  o.y = prmt(v.x, v.z, 0x7632);
  o.z = prmt(v.y, v.w, 0x5410);
  o.w = prmt(v.y, v.w, 0x7632);
  if (kScaleOdd) {
    const __half2 sixteenth = u2h2(0x2C002C00u);   // 0.0625
    o.y = h22u(__hmul2(u2h2(o.y), sixteenth));
    o.w = h22u(__hmul2(u2h2(o.w), sixteenth));
  }
  return o;
}


// CTA = 8 consumer warps (WC column chunks of 32 columns x WK = 8/WC K-slices) + 1 producer warp.
constexpr int kMaxStages = 8;
constexpr int kConsumerThreads = 8 * 32;
constexpr int kW4Threads = kConsumerThreads + 32;

template <int UPG, int WC>
struct W4Cfg {
  static constexpr int WK = 8 / WC;
  static constexpr int NT = 32 * WC;                 // columns per CTA
  static constexpr int GPB = 4 / UPG;                // scale groups per 128-k block
  static constexpr int kBoxRows = WK * 16;           // packed word-rows per stage (WK blocks of 128 k)
  static constexpr int kBoxBytes = kBoxRows * 128;   // one weight box: kBoxRows x 32 columns, SWIZZLE_128B
  static constexpr int kWeights = WC * kBoxBytes;    // = 16 KiB for every WC
  static constexpr int kGroupRows = WK * GPB;        // scale / zero rows per stage
  static constexpr int kScales = kGroupRows * NT * 2;
  static constexpr int kZeros = kGroupRows * (NT / 8) * 4;
  static constexpr int kStageBytes = (kWeights + kScales + kZeros + 1023) / 1024 * 1024;
  static constexpr uint32_t kTxBytes = kWeights + kScales + kZeros;   // TMA boxes always deliver their full size
};

// Per-lane constants of a consumer warp (wc = column chunk, wk = K-slice of the stage).
template <int MT>
struct W4Lane {
  uint32_t w_x0, w_x1;      // byte offsets of this lane's 16-byte chunk for units with u%2 == 0 / 1
  uint32_t s_off, z_off;    // byte offsets of this lane's scales / zero word inside a stage
  int zshift;               // 0 or 16: which half of the zero word holds this lane's 4 nibbles
  int r;                    // lane % 4
  int brow_off[MT > 0 ? MT : 1];   // activation row offset (halves) of this lane's batch column(s)
  float zbias;
  uint32_t zero_bias;
};

// One 128-k block (GPB scale groups of UPG 32-k units) of one stage: `st` = stage base, `ablk` = this
// lane's activations of the block (act_sm + block offset + lane word-row), `asum_blk` = group sums of
// the block ([GPB][MROWS], mma only).  Accumulates into tot.
// SIMT family only (MT == 0); the tensor-core families use w4_consume_block_v2 below.  (HYB is a leftover template
// parameter of the kernel: the FHFMA / HMMA hybrid it selected was measured in round 1, did not reduce issue slots, and
// has been removed together with the first tensor-core block math.)
template <int MT, int UPG, int WC, int HYB>
__device__ __forceinline__ void w4_consume_block(const unsigned char* __restrict__ st, const __half* __restrict__ ablk,
                                                 const float* __restrict__ asum_blk, const W4Lane<MT>& L,
                                                 float (&tot)[MT > 0 ? 2 * MT : 1][4], float (&tot_s)[4]) {
  using Cfg = W4Cfg<UPG, WC>;
  constexpr bool kMma = MT > 0;
  constexpr int MROWS = kMma ? 8 * MT : 1;
  constexpr int NT = Cfg::NT, GPB = Cfg::GPB;
  auto unit_row = [](int u) constexpr { return (UPG == 1) ? 4 * u : 8 * (u >> 1) + (u & 1); };
  const uint32_t s_off = L.s_off, z_off = L.z_off, w_x0 = L.w_x0, w_x1 = L.w_x1;
  const int zshift = L.zshift, r = L.r;
  const float zbias = L.zbias;
  (void)r; (void)zbias; (void)asum_blk; (void)MROWS;
#pragma unroll
  for (int q = 0; q < GPB; ++q) {
    const uint2 sraw = *reinterpret_cast<const uint2*>(st + s_off + q * (NT * 2));
    const uint32_t zraw = *reinterpret_cast<const uint32_t*>(st + z_off + q * (NT / 2)) >> zshift;
    float sf[4];
    {
      const float2 s01 = __half22float2(u2h2(sraw.x));
      const float2 s23 = __half22float2(u2h2(sraw.y));
      sf[0] = s01.x; sf[1] = s01.y; sf[2] = s23.x; sf[3] = s23.y;
    }
    // (the first tensor-core block math -- IMAD.HI + 4 LOP3 per word, zero point in the epilogue, and its FHFMA / HMMA
    // hybrid -- lived here in round 1; w4_consume_block_v2 replaced it everywhere, so only the SIMT family is left)
    static_assert(!kMma, "tensor-core families use w4_consume_block_v2");
    {
      uint32_t zlo[4], zhi[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t zb = ((zraw >> (4 * j)) & 0xFu) + L.zero_bias;   // <= 16
        zlo[j] = (magic_base_bits(0) + zb) * 0x00010001u;          // half2(1024 + z)
        zhi[j] = (magic_base_bits(4) + (zb << 4)) * 0x00010001u;   // half2(64 + z)
      }
      float grp[4] = {0.f, 0.f, 0.f, 0.f};
      __half2 acc[4];
#pragma unroll
      for (int uu = 0; uu < UPG; ++uu) {
        const int u = q * UPG + uu;
        const uint4 wv = *reinterpret_cast<const uint4*>(st + ((u & 1) ? w_x1 : w_x0) + unit_row(u) * 128);
        const uint4 av = *reinterpret_cast<const uint4*>(ablk + unit_row(u) * 8);
        const uint32_t w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t e[4];
          unpack_w4_minus_zero(w4[j], zlo[j], zhi[j], e);
          __half2 c;
          if (uu & 1) c = __hfma2(u2h2(e[0]), u2h2(av.x), acc[j]);
          else        c = __hmul2(u2h2(e[0]), u2h2(av.x));
          c = __hfma2(u2h2(e[1]), u2h2(av.y), c);
          c = __hfma2(u2h2(e[2]), u2h2(av.z), c);
          c = __hfma2(u2h2(e[3]), u2h2(av.w), c);
          acc[j] = c;
        }
        if ((uu & 1) || uu == UPG - 1) {      // flush the fp16 chains to fp32 every 2 units
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(acc[j]);
            grp[j] += f.x + f.y;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) tot[0][j] = fmaf(sf[j], grp[j], tot[0][j]);
    }
  }
}

template <int UPG>
__device__ __forceinline__ int w4_lane_row(int lane) { return ((UPG == 1) ? 1 : 2) * (lane & 3); }

// Word-row of (unit u, lane r) inside a 16-row block: UPG >= 2: 8*(u/2) + 2r + (u%2) -- conflict-free
// under the 128-byte swizzle; UPG == 1 (groupsize 32): 4u + r, so that a unit stays inside one scale
// group (2-way conflict).  16-byte chunk c8 of row rho lives at chunk c8 ^ (rho & 7).
template <int MT, int UPG, int WC>
__device__ __forceinline__ W4Lane<MT> make_w4_lane(int lane, int wc, int wk, int M, int pitch, int zero_bias) {
  using Cfg = W4Cfg<UPG, WC>;
  constexpr int kOddRowAdd = (UPG == 1) ? 4 : 1;
  W4Lane<MT> L;
  const int r = lane & 3, c8 = lane >> 2;
  const int lane_row = w4_lane_row<UPG>(lane);
  const int col_local = 32 * wc + 4 * c8;                         // this lane's first column inside the tile
  const uint32_t w_row = (uint32_t)(wc * Cfg::kBoxBytes + (wk * 16 + lane_row) * 128);
  L.w_x0 = w_row + (uint32_t)((c8 ^ lane_row) * 16);
  L.w_x1 = w_row + (uint32_t)((c8 ^ (lane_row + kOddRowAdd)) * 16);
  L.s_off = (uint32_t)(Cfg::kWeights + wk * Cfg::GPB * (Cfg::NT * 2) + col_local * 2);
  L.z_off = (uint32_t)(Cfg::kWeights + Cfg::kScales + wk * Cfg::GPB * (Cfg::NT / 2) + (col_local >> 3) * 4);
  L.zshift = 16 * (c8 & 1);
  L.r = r;
#pragma unroll
  for (int mt = 0; mt < (MT > 0 ? MT : 1); ++mt) L.brow_off[mt] = min(c8 + 8 * mt, M - 1) * pitch;
  L.zbias = (float)zero_bias;
  L.zero_bias = (uint32_t)zero_bias;
  return L;
}

// ------------------------------------------------------------------------------------------------
// v2 block math (tensor-core path of both W4 kernels).  Same exact-product idea as
// w4_consume_block, with fewer instructions per 128-k block -- instruction issue is what bounds an
// SM here (tools/pipe_probe2.cu):
//   * unpack = 2 PRMT + 2 LOP3 per word instead of IMAD.HI + 4 LOP3: the byte pairs (k0 + 16 k1) are
//     themselves exact fp16 subnormals, so  k0 a0 + k1 a1 = k0 (a0 - a1/16) + (k0 + 16 k1) (a1/16)
//     and only the low nibble needs a mask; the activations are staged as (a0 - a1/16, a1/16);
//   * the zero point rides on the tensor core: one extra m16n8k16 per column pair and scale group
//     with A = -(z + bias) * 64 * 2^-24 in k slots 0, 1 and B = (hi, lo) halves of sum_k a_k / 64, so the
//     group accumulator already holds 2^-24 * sum_k a_k (w_k - z) and the epilogue is one FFMA per
//     accumulator; the 2^24 is applied once per segment.
template <int MT>
struct W4Lane2 {
  uint32_t w_x0;            // byte offset of this lane's 16-byte chunk for units with u%2 == 0 (odd units: ^ kOddXor)
  uint32_t s_off, z_off;    // byte offsets of this lane's 4 scales / 4 zero nibbles (16 bits) inside a stage
  uint32_t zmul, zadd;      // (z * zmul + zadd) = half2 bits of -(z + bias) * 64 * 2^-24, twice, in lanes r == 0; 0 elsewhere
  int brow_off[MT];         // activation row offset (halves) of this lane's batch column(s)
  int zt_off[MT];           // byte offset of this lane's (hi, lo) group-sum word inside a group's table row
};

template <int MT, int UPG, int WC>
__device__ __forceinline__ W4Lane2<MT> make_w4_lane2(int lane, int wc, int wk, int M, int pitch, int zero_bias) {
  using Cfg = W4Cfg<UPG, WC>;
  W4Lane2<MT> L;
  const int r = lane & 3, c8 = lane >> 2;
  const int lane_row = w4_lane_row<UPG>(lane);
  const int col_local = 32 * wc + 4 * c8;
  L.w_x0 = (uint32_t)(wc * Cfg::kBoxBytes + (wk * 16 + lane_row) * 128 + ((c8 ^ lane_row) * 16));
  L.s_off = (uint32_t)(Cfg::kWeights + wk * Cfg::GPB * (Cfg::NT * 2) + col_local * 2);
  L.z_off = (uint32_t)(Cfg::kWeights + Cfg::kScales + wk * Cfg::GPB * (Cfg::NT / 2) + (col_local >> 2) * 2);
  L.zmul = r == 0 ? 0x00400040u : 0u;
  L.zadd = r == 0 ? (uint32_t)zero_bias * 0x00400040u + 0x80008000u : 0u;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int m = min(c8 + 8 * mt, M - 1);
    L.brow_off[mt] = m * pitch;
    L.zt_off[mt] = (m * 4 + r) * 4;
  }
  return L;
}


// One 128-k block: `st` = stage base, `ablk` = this lane's activations of the block, `zt_blk` = the
// block's group-sum table ([GPB][M][4] words).  tot accumulates 2^-24 * y.
template <int MT, int UPG, int WC>
__device__ __forceinline__ void w4_consume_block_v2(const unsigned char* __restrict__ st, const __half* __restrict__ ablk,
                                                    const unsigned char* __restrict__ zt_blk, int zt_group_bytes,
                                                    const W4Lane2<MT>& L, float (&tot)[2 * MT][4]) {
  using Cfg = W4Cfg<UPG, WC>;
  constexpr int NT = Cfg::NT, GPB = Cfg::GPB;
  constexpr uint32_t kOddXor = (UPG == 1) ? 64u : 16u;
  auto unit_row = [](int u) constexpr { return (UPG == 1) ? 4 * u : 8 * (u >> 1) + (u & 1); };
#pragma unroll
  for (int q = 0; q < GPB; ++q) {
    const uint2 sraw = *reinterpret_cast<const uint2*>(st + L.s_off + q * (NT * 2));
    const uint32_t zraw = *reinterpret_cast<const unsigned short*>(st + L.z_off + q * (NT / 2));
    uint32_t bz[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) bz[mt] = *reinterpret_cast<const uint32_t*>(zt_blk + q * zt_group_bytes + L.zt_off[mt]);

    // (hoisting all LDS.128 of the group in front of the math, or splitting the dependent HMMA chains over two
    // accumulator sets, was measured: no gain -- the loop is issue bound -- and 24 more live registers)
    float grp[2 * MT][4];
#pragma unroll
    for (int uu = 0; uu < UPG; ++uu) {
      const int u = q * UPG + uu;
      const uint4 wv = *reinterpret_cast<const uint4*>(st + ((u & 1) ? (L.w_x0 ^ kOddXor) : L.w_x0) + unit_row(u) * 128);
      uint4 bfrag[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
        bfrag[mt] = *reinterpret_cast<const uint4*>(ablk + L.brow_off[mt] + unit_row(u) * 8);
      const uint32_t w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int tt = 0; tt < 2; ++tt) {
        uint32_t ea[4], eb[4];
        unpack_w4_bytes(w4[2 * tt], ea);
        unpack_w4_bytes(w4[2 * tt + 1], eb);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          if (uu == 0) mma_m16n8k16_zero(grp[tt * MT + mt], ea[0], eb[0], ea[1], eb[1], bfrag[mt].x, bfrag[mt].y);
          else         mma_m16n8k16(grp[tt * MT + mt], ea[0], eb[0], ea[1], eb[1], bfrag[mt].x, bfrag[mt].y);
          mma_m16n8k16(grp[tt * MT + mt], ea[2], eb[2], ea[3], eb[3], bfrag[mt].z, bfrag[mt].w);
        }
      }
    }
    {
      // zero point last (its operands come out of the longest scalar chain): -(z + bias) * sum_k a_k on the tensor core
      uint32_t za[4];
      za[0] = (zraw & 0xFu) * L.zmul + L.zadd;
      za[1] = ((zraw >> 4) & 0xFu) * L.zmul + L.zadd;
      za[2] = ((zraw >> 8) & 0xFu) * L.zmul + L.zadd;
      za[3] = (zraw >> 12) * L.zmul + L.zadd;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) mma_m16n8k16(grp[tt * MT + mt], za[2 * tt], za[2 * tt + 1], 0u, 0u, bz[mt], 0u);
    }
    // grp = 2^-24 * sum_k a_k (w_k - z).  accumulators 0,1 belong to column 2*tt (rows m = 2r, 2r+1), 2,3 to column 2*tt+1.
    const float2 s01 = __half22float2(u2h2(sraw.x));
    const float2 s23 = __half22float2(u2h2(sraw.y));
    const float sf[4] = {s01.x, s01.y, s23.x, s23.y};
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int tt = 0; tt < 2; ++tt)
#pragma unroll
        for (int i = 0; i < 4; ++i) tot[tt * MT + mt][i] = fmaf(sf[2 * tt + (i >> 1)], grp[tt * MT + mt][i], tot[tt * MT + mt][i]);
  }
}

// Every cross-GPU spin in this file gives up after about 30 s of SM clocks and reports it (timeout flag /
// local_state[3]) instead of hanging the GPU.  Long on purpose: the peer may simply be late (another process,
// lazy module loading, a descheduled host thread); 2 s was observed to fire spuriously once in a few runs.
constexpr long long kSpinGuardClocks = 60000000000ll;

// Fused gather wait: ONE lane per CTA polls this rank's flag array until every rank's slot has reached
// this rank's own published call count (relaxed system-scope polls with a back-off, one fence at the
// end).  Thousands of threads polling one L2 sector delay the very NVLink write they wait for: with
// all 8 warps of every CTA polling, the wait cost 14 us per call (profiles/r01_v6_bench_n2_*).
__device__ __forceinline__ bool peers_poll(const unsigned int* flags, int world, int rank, int lane) {
  bool ok = true;
  if (lane == 0) {
    unsigned int expected, f;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(expected) : "l"(flags + rank) : "memory");
    const long long c0 = clock64();
    for (int p = 0; p < world; ++p) {
      for (unsigned int n = 1;; ++n) {
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(flags + p) : "memory");
        if ((int)(f - expected) >= 0) break;
        if ((n & 255u) == 0 && clock64() - c0 > kSpinGuardClocks) { ok = false; break; }
        __nanosleep(200);
      }
    }
    __threadfence_system();
  }
  __syncwarp();
  return ok;
}

// Flag-in-data ("LL") exchange helpers.  A slot is 8 bytes {two fp16 results, call number}, written with
// ONE 8-byte store (atomic on NVLink), so the flag validates its own data and nobody needs a fence.
__device__ __forceinline__ void ll_store(unsigned long long* slot, uint32_t data, uint32_t epoch) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(slot), "r"(data), "r"(epoch) : "memory");
}
// 8 consecutive halves = 4 slots = 32 bytes; spins until all four carry `epoch` (guard: reports in local_state[3])
__device__ __forceinline__ uint4 ll_load8(const unsigned long long* slots, uint32_t epoch, unsigned int* timeout_flag) {
  uint4 q0, q1;
  const long long c0 = clock64();
  for (unsigned int n = 1;; ++n) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w) : "l"(slots) : "memory");
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w) : "l"(slots + 2) : "memory");
    if (q0.y == epoch && q0.w == epoch && q1.y == epoch && q1.w == epoch) break;
    if ((n & 1023u) == 0 && clock64() - c0 > kSpinGuardClocks) {
      *timeout_flag = 1u;
      break;
    }
  }
  return make_uint4(q0.x, q0.z, q1.x, q1.z);
}

// tools/trace.py: wall-clock stamps of one CTA's phases (debug only; a.trace is null in production)
__device__ __forceinline__ void trace_stamp(const GemvArgs& a, int slot) {
  if (a.trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + slot] = t;
  }
}

// MT = 0: SIMT kernel, M == 1.  MT = 1, 2: mma.sync kernel, M <= 8 * MT.
// UPG = 32-k units per scale group inside a 128-k block: 4 (groupsize 128), 2 (64), 1 (32).
//
// K is cut into 128-k blocks (16 packed word-rows); blocks [b0, b1) belong to this CTA (cluster
// rank = blockIdx.y).  One elected producer thread streams them through a kStages-deep shared-
// memory ring, WK blocks per stage: WC weight boxes + one scale box + one zero box per stage, all
// completing on the stage's "full" mbarrier.  (Per-row 512-byte cp.async.bulk copies were measured
// first: the TMA unit then limits a CTA to ~8 B/clk -- profiles/r01_bw_probe_access_patterns.log.)
// Consumer warp (wc, wk) takes block wk of every stage, columns [32*wc, 32*wc+32): lane
// (r = lane%4, c8 = lane/4) reads, per 32-k unit u, word-row 8*(u/2) + 2r + (u%2) of its block
// (4u + r when a unit must stay inside one 32-k group), columns 4*c8..4*c8+3, with one LDS.128 --
// the row choice makes the swizzled access conflict-free, and because the order of K inside an MMA
// is free as long as the activation fragment follows it, those four words ARE the A-fragment
// sources of two m16n8k16 tiles.  When a warp has consumed a stage it arrives on the stage's
// "empty" mbarrier and the producer refills it.  No global address arithmetic, bounds checks or
// register landing buffers in the consumer loop; bytes in flight are bounded by shared memory
// (4 stages x 17 KiB per CTA), not by registers.
__device__ __forceinline__ void trace_value(const GemvArgs& a, int slot, unsigned long long v) {
  if (a.trace) a.trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + slot] = v;
}

template <int MT, int UPG, int WC, int HYB>
// Two CTAs per SM.  A third slot (72 registers, 3-stage ring) was measured twice: the next launch's CTAs do become
// resident early and prefetch, but the hardware then packs three CTAs of ONE launch on some SMs and one on others,
// and the imbalance costs more than the prefetch wins (4096x4096: 5.5 vs 4.3 us; profiles/r01_v6_trace_*).
__global__ void __launch_bounds__(kW4Threads, 2)
gemv_w4_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap smap,
               const __grid_constant__ CUtensorMap zmap, const GemvArgs a) {
  using Cfg = W4Cfg<UPG, WC>;
  constexpr bool kMma = MT > 0;
  constexpr bool kV2 = kMma && HYB == 0;            // v2 block math (see w4_consume_block_v2)
  constexpr int MROWS = kMma ? 8 * MT : 1;
  constexpr int WK = Cfg::WK, NT = Cfg::NT, GPB = Cfg::GPB;
  // lane-independent part of the word-row index of unit u
  auto unit_row = [](int u) constexpr { return (UPG == 1) ? 4 * u : 8 * (u >> 1) + (u & 1); };
  extern __shared__ __align__(1024) unsigned char smem_raw[];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.y;
  const int n_cta = blockIdx.x * NT;
  const int cw = min(NT, a.N - n_cta);              // valid columns of this tile (multiple of 32)
  const int nblocks = a.K >> 7;
  const int b0 = min(split * a.units_per_split, nblocks);
  const int b1 = min(b0 + a.units_per_split, nblocks);
  const int ntiles = (b1 - b0 + WK - 1) / WK;
  const int pitch = a.units_per_split * 128 + 8;    // halves per staged activation row (+16 B: batch rows land in different banks)
  const int ngroups = a.units_per_split * GPB;      // scale groups in this CTA's K range
  const int kStages = a.ring;

  // SWIZZLE_128B boxes need 1024-byte aligned destinations; the dynamic segment only promises 16
  unsigned char* stage_base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_base + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  __half* act_sm = reinterpret_cast<__half*>(stage_base + kStages * Cfg::kStageBytes + 128);        // [M][pitch]
  float* asum_sm = reinterpret_cast<float*>(act_sm + (size_t)a.M * pitch);                         // v1 mma: [ngroups][MROWS] floats; v2: [ngroups][M][4] words
  float* red_sm = asum_sm + (kV2 ? ngroups * a.M * 4 : (kMma ? ngroups * MROWS : 0));              // [WK][M][NT]
  float* clus_sm = red_sm + WK * a.M * NT;          // [splits][M][NT], only the cluster leader's is used

  const bool clustered = a.splits > 1;
  if (clustered) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");   // "I have started"
  if (tid == 0) {
    trace_stamp(a, 0);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // Let the next kernel in the stream become resident now: its producer starts streaming ITS weights
  // while this kernel is still running (its consumers block in griddepcontrol.wait until this grid
  // has completed and flushed).  One kernel uses at most about half of an SM's shared memory.
  // Exception: the FIRST call of an LL chain.  The calls behind it do not wait for the grid before them, and they
  // read the chain base that the previous chain's xbit_ll_unpack_f16 advances; so this call releases its dependents
  // only after its own griddepcontrol.wait has returned, i.e. after that kernel has completed (otherwise the next
  // calls can start on the stale base, match the previous chain's slots and publish call numbers nobody waits for).
  const bool defer_dependents = a.ll_out && !a.a_is_ll;
  if (!defer_dependents) griddep_launch_dependents();

  float tot[kMma ? 2 * MT : 1][4];
  float tot_s[4] = {0.f, 0.f, 0.f, 0.f};            // FHFMA partial sums (HYB only)
#pragma unroll
  for (int v = 0; v < (kMma ? 2 * MT : 1); ++v)
#pragma unroll
    for (int i = 0; i < 4; ++i) tot[v][i] = 0.f;
  const int r = lane & 3, c8 = lane >> 2;
  const int wc = warp & (WC - 1), wk = (warp / WC) & (WK - 1);

  if (warp == 8) {
    // =========================== producer ===========================
    // Default: nothing is read before the previous kernel in the stream has completed.  With
    // XBIT_GEMV_FLAG_STATIC_WEIGHTS the caller promises the weights were not produced by that
    // kernel: the weight stream starts at once and only the activation staging waits.
    if (!a.static_weights) griddep_wait();
    if (lane == 0) {
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&smap) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&zmap) : "memory");
      trace_stamp(a, 1);
      int s = 0, ph = 0;                            // slot and the parity of its fill count, kept without divisions
      for (int t = 0; t < ntiles; ++t) {
        if (t >= kStages) mbar_wait(&empty_bar[s], ph ^ 1);
        const int blk = b0 + t * WK;
        unsigned char* st = stage_base + s * Cfg::kStageBytes;
        mbar_arrive_expect_tx(&full_bar[s], Cfg::kTxBytes);
#pragma unroll
        for (int c = 0; c < WC; ++c) tma_load_2d(st + c * Cfg::kBoxBytes, &wmap, n_cta + 32 * c, blk * 16, &full_bar[s], policy);
        tma_load_2d(st + Cfg::kWeights, &smap, n_cta, blk * GPB, &full_bar[s], policy);
        tma_load_2d(st + Cfg::kWeights + Cfg::kScales, &zmap, n_cta >> 3, blk * GPB, &full_bar[s], policy);
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // =========================== consumers ===========================
    // stage the activations of this CTA's K range (the only data that depends on the previous
    // kernel); the mma path also needs sum_k a_k per scale group for the folded zero point
    // A call fed from an LL buffer carries its dependency in the data (every slot is validated by its own call
    // number), so it does not wait for the previous grid to complete and flush: its CTAs start staging as soon as
    // they are resident.  (The previous launch's CTAs are all running by then -- that is when a programmatic
    // dependent launch happens -- so whoever we spin on is making progress.)
    if (!a.a_is_ll) griddep_wait();
    if (defer_dependents) griddep_launch_dependents();
    if (a.sig_wait) {                               // the previous N-split call has landed here (warp 0 polls, the others wait for it)
      if (warp == 0) peers_poll(a.sig_flags[a.sig_rank], a.world, a.sig_rank, lane);
      asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
    }
    if (tid == 0) trace_stamp(a, 2);
    {
      // no divisions in here: the loop is instruction bound, not latency bound (tools/trace.py)
      const int vecs_per_row = (b1 - b0) * 16;      // 8-half vectors (a multiple of 16); a scale group = 4*UPG consecutive vectors
      for (int m = 0; m < a.M; ++m) {
        const uint4* arow = reinterpret_cast<const uint4*>(a.a + (size_t)m * a.K + b0 * 128);
        // LL form: row m of the previous call's result, 2 halves per 8-byte slot; its call number = calls completed so far
        const unsigned long long* ll_row = reinterpret_cast<const unsigned long long*>(a.a) + (((size_t)m * a.K + b0 * 128) >> 1);
        const uint32_t ll_epoch = a.a_is_ll ? a.sig_state[2] + (uint32_t)a.ll_chain_index : 0u;   // the previous call's number
        __half* srow = act_sm + (size_t)m * pitch;
        for (int v = tid; v - lane < vecs_per_row; v += kConsumerThreads) {      // warp-uniform trip count (shuffles below)
          const bool ok = v < vecs_per_row;
          uint4 val = make_uint4(0, 0, 0, 0);
          if (ok) {
            if (a.a_is_ll) val = ll_load8(ll_row + 4 * (size_t)v, ll_epoch, a.sig_state + 3);   // arrives slot by slot from every rank
            else           val = __ldcg(arow + v);                            // L2 only: may just have been written by peer GPUs
            *reinterpret_cast<uint4*>(srow + v * 8) = kV2 ? permute_act8_v2(val) : permute_act8<kMma>(val);
          }
          if constexpr (kMma) {
            const float2 f0 = __half22float2(u2h2(val.x)), f1 = __half22float2(u2h2(val.y));
            const float2 f2 = __half22float2(u2h2(val.z)), f3 = __half22float2(u2h2(val.w));
            float sum = ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
            // a group's 4*UPG vectors sit in 4*UPG consecutive lanes
#pragma unroll
            for (int o = 1; o < 4 * UPG; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (ok && (lane & (4 * UPG - 1)) == 0) {
              if constexpr (kV2) {
                // sum_k a_k / 64 as an fp16 (hi, lo) pair in the r == 0 slot, zeros in the other three
                const float q64 = sum * 0.015625f;
                const __half hi = __float2half_rn(q64);
                const __half lo = __float2half_rn(q64 - __half2float(hi));
                *reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(asum_sm) + ((size_t)(v / (4 * UPG)) * a.M + m) * 4) =
                    make_uint4((uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16), 0u, 0u, 0u);
              } else {
                asum_sm[(v / (4 * UPG)) * MROWS + m] = sum;
              }
            }
          }
        }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
    if (tid == 0) trace_stamp(a, 3);

    const W4Lane<MT> L = make_w4_lane<MT, UPG, WC>(lane, wc, wk, a.M, pitch, a.zero_bias);
    const W4Lane2<kMma ? MT : 1> L2 = make_w4_lane2<kMma ? MT : 1, UPG, WC>(lane, wc, wk, a.M, pitch, a.zero_bias);
    const int zt_group_bytes = a.M * 16;
    (void)L; (void)L2; (void)zt_group_bytes;
    const __half* aptr = act_sm + wk * 128 + w4_lane_row<UPG>(lane) * 8;   // this lane's word-row; + t * WK * 128 per stage

    int s = 0, ph = 0;
    const long long loop0 = a.trace ? clock64() : 0;
    for (int t = 0; t < ntiles; ++t) {
      mbar_wait(&full_bar[s], ph);
      if (a.trace && tid == 0 && t == 0) trace_stamp(a, 4);
      const unsigned char* st = stage_base + s * Cfg::kStageBytes;
      const int blk_local = t * WK + wk;                          // block index inside this CTA's range
      if (b0 + blk_local < b1) {                                  // warp-uniform (the last stage may be partly empty)
        if (!a.debug_skip) {
          if constexpr (kV2)
            w4_consume_block_v2<MT, UPG, WC>(st, aptr + t * (WK * 128), reinterpret_cast<const unsigned char*>(asum_sm) + blk_local * GPB * zt_group_bytes,
                                             zt_group_bytes, L2, tot);
          else
            w4_consume_block<MT, UPG, WC, HYB>(st, aptr + t * (WK * 128), asum_sm + blk_local * GPB * MROWS, L, tot, tot_s);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
      if (++s == kStages) { s = 0; ph ^= 1; }
    }
    if (tid == 0 && a.trace) {
      trace_value(a, 8, (unsigned long long)(clock64() - loop0));
      trace_value(a, 11, (unsigned long long)ntiles);
    }
  }

  if (tid == 0) trace_stamp(a, 5);
  // ---- split-K reduction: r-lanes (shuffle, SIMT only) -> K-slices (smem) -> cluster (DSMEM)
  if (warp < 8) {
    if constexpr (kMma && HYB != 0 && MT == 1) {
      // FHFMA partials: sum over the 4 r-lanes, then into the m = 0 accumulators held by the r == 0 lanes
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = tot_s[j];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (r == 0) tot[j >> 1][(j & 1) * 2] += v;
      }
    }
    if constexpr (!kMma) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = tot[0][j];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (r == 0) red_sm[wk * NT + 32 * wc + 4 * c8 + j] = v;
      }
    } else {
#pragma unroll
      for (int tt = 0; tt < 2; ++tt)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int m = 8 * mt + 2 * r + (i & 1);
            const int col = 32 * wc + 4 * c8 + 2 * tt + (i >> 1);
            if (m < a.M) red_sm[(wk * a.M + m) * NT + col] = kV2 ? tot[tt * MT + mt][i] * 16777216.f : tot[tt * MT + mt][i];
          }
    }
  }
  __syncthreads();

  const int nout = a.M * NT;
  if (clustered) {
    cg::cluster_group cluster = cg::this_cluster();
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // every CTA of the cluster has started
    float* leader = cluster.map_shared_rank(clus_sm, 0);
    for (int o = tid; o < nout; o += kW4Threads) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < WK; ++w) v += red_sm[w * nout + o];
      leader[split * nout + o] = v;
    }
    cluster.sync();
    if (tid == 0) trace_stamp(a, 6);
    if (split != 0) return;
  }
  auto tile_sum = [&](int o) {
    float v = 0.f;
    if (clustered) {
      for (int s = 0; s < a.splits; ++s) v += clus_sm[s * nout + o];
    } else {
#pragma unroll
      for (int w = 0; w < WK; ++w) v += red_sm[w * nout + o];
    }
    return v;
  };
  if (a.ll_out) {
    // flag-in-data all-gather: each pair of results goes to every rank as one {half2, call number} store
    // call number = chain base (device memory, bumped by xbit_ll_unpack_f16 at the end of every chain, so
    // graphs replay) + position in the chain + 1: known at launch, no counter shared with overlapping launches
    const uint32_t epoch = a.sig_state[2] + (uint32_t)a.ll_chain_index + 1u;
    for (int o2 = tid; o2 < (nout >> 1); o2 += kW4Threads) {
      const int m = o2 / (NT / 2), col = 2 * (o2 - m * (NT / 2));
      if (col < cw) {
        const __half2 h2 = __floats2half2_rn(tile_sum(m * NT + col), tile_sum(m * NT + col + 1));
        const size_t slot = ((size_t)m * a.ldo + a.col_offset + n_cta + col) >> 1;
        for (int p = 0; p < a.world; ++p) ll_store(reinterpret_cast<unsigned long long*>(a.out[p]) + slot, h22u(h2), epoch);
      }
    }
    if (tid == 0) trace_stamp(a, 7);
    return;
  }
  for (int o = tid; o < nout; o += kW4Threads) {
    const int m = o / NT, col = o - m * NT;
    const float v = tile_sum(o);
    if (col < cw) {
      const __half h = __float2half_rn(v);
      const size_t off = (size_t)m * a.ldo + a.col_offset + n_cta + col;
      a.out[0][off] = h;
      for (int p = 1; p < a.world; ++p) a.out[p][off] = h;   // fused all-gather: NVLink peer stores
    }
  }
  if (a.sig_state != nullptr) {
    // Fused completion signal: the column tile that finishes LAST on this GPU tells every rank
    // "rank sig_rank's slice of call #epoch has landed in your buffer".  Stores of every tile are made
    // visible system-wide (one fence per CTA, after the CTA barrier) before its count; the last one fences
    // again before the flags.
    __syncthreads();
    if (tid == 0) {
      // one thread, cumulative over the CTA's stores ordered by the barrier.  GPU scope here (a system
      // fence per CTA was measured at +8 us per call); the CTA that arrives last issues the one
      // system-scope fence, which is cumulative over everything it has observed through the counter.
      __threadfence();
      const unsigned int done = atomicAdd(a.sig_state, 1u);
      if (done == gridDim.x - 1) {
        a.sig_state[0] = 0u;
        const unsigned int epoch = a.sig_state[1] + 1u;
        a.sig_state[1] = epoch;
        __threadfence_system();
        for (int p = 0; p < a.world; ++p)
          asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a.sig_flags[p] + a.sig_rank), "r"(epoch) : "memory");
      }
    }
  }
  if (tid == 0) trace_stamp(a, 7);
}

// ------------------------------------------------------------------------------------------------
// Persistent stream-K schedule: one CTA per SM that owns HALF an SM, perfectly balanced.
//
// Why (tools/pipe_probe2.cu, tools/trace.py; profiles/r01_v5_*): on B200 an SM sub-partition retires
// about one LOP3 / IMAD / FFMA per 2 cycles whatever the mix (the ALU and FMA pipes do not overlap)
// and an HMMA per ~4, so the consumer instruction stream -- not HBM, not occupancy -- bounds an SM at
// 2 KB per ~2 x (instructions per block) cycles: 8 warps already reach it, 16 add nothing.  What is
// left to win is (a) balance: any (tiles x splits) grid leaves part of the SMs with half the work of
// the others, and (b) between back-to-back calls only what already sits in shared memory when the
// previous kernel retires is free.  So:
//   * one CTA per SM (8 consumer warps + 1 TMA producer warp) using at most half of the SM's shared
//     memory and registers: the NEXT launch's CTA is resident on the same SM and
//     (XBIT_GEMV_FLAG_STATIC_WEIGHTS) fills its whole ring while this one computes;
//   * work = (128-column tile, stage of 2 x 128 k) units in tile-major order, W = tiles * S_t; CTA c
//     of G owns the contiguous range [W*c/G, W*(c+1)/G): the tail of one tile, whole tiles, the head
//     of another.  Whole tiles are written directly.  For a shared tile the CTA holding its LAST
//     stage is the finisher: the others publish fp32 partial tiles in the workspace (release flag);
//     the finisher adds them in CTA order (deterministic, no atomics) and clears the flags.  Waits
//     only ever target lower-numbered CTAs, which never wait themselves;
//   * the activation rows are staged per CTA for exactly the k ranges its units touch.
constexpr int kSkMaxRing = 8;

template <int MT, int UPG>
__global__ void __launch_bounds__(kW4Threads, MT == 2 ? 1 : 2)
gemv_w4_streamk_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap smap,
                       const __grid_constant__ CUtensorMap zmap, const GemvArgs a) {
  static_assert(MT == 1 || MT == 2, "tensor-core path only");
  constexpr int WC = 4;
  using Cfg = W4Cfg<UPG, WC>;
  constexpr int WK = Cfg::WK, NT = Cfg::NT, GPB = Cfg::GPB;
  extern __shared__ __align__(1024) unsigned char smem_raw[];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, c = blockIdx.x;
  const int S_t = a.sk_stages_per_tile;
  const long long W = a.sk_total_stages;
  const int s_lo = (int)(W * c / G), s_hi = (int)(W * (c + 1) / G);
  const int nst = s_hi - s_lo;
  const int ring = a.sk_ring;
  const int act_units = a.sk_act_units;             // 256-k activation chunks staged per CTA
  const bool act_abs = a.sk_act_abs != 0;           // chunk index = stage inside the tile (whole K staged) / unit index
  const int pitch = act_units * (WK * 128) + 8;     // halves per staged activation row (+16 B: batch rows spread over banks)
  const int tile0 = s_lo / S_t, stg0 = s_lo - tile0 * S_t;
  // Processing order.  If the range starts with the tail of a tile somebody else began (this CTA is that
  // tile's finisher) and goes on, that tail is processed LAST: the head of the last tile (a contribution to
  // a higher-numbered CTA) is then published early, and by the time this CTA needs the contributions
  // to its own first tile they are usually there.  Units [rot, nst) first, then [0, rot).
  const int rot = (stg0 > 0 && S_t - stg0 < nst) ? S_t - stg0 : 0;

  unsigned char* stage_base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_base + ring * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kSkMaxRing;
  __half* act_sm = reinterpret_cast<__half*>(stage_base + ring * Cfg::kStageBytes + 128);          // [M][pitch]
  uint32_t* zt_sm = reinterpret_cast<uint32_t*>(act_sm + (size_t)a.M * pitch);                     // [act_units*WK*GPB][M][4]: (hi, lo) of sum_k a_k / 64, then 3 zero words
  float* red_sm = reinterpret_cast<float*>(zt_sm + (size_t)act_units * WK * GPB * a.M * 4);         // [WK][M][NT]
  float* first_sm = red_sm + WK * a.M * NT;                                                        // [M][NT] deferred first tile

  if (tid == 0) {
    trace_stamp(a, 0);
    for (int s = 0; s < ring; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  griddep_launch_dependents();     // see gemv_w4_kernel

  if (warp == 8) {
    // =========================== producer ===========================
    if (!a.static_weights) griddep_wait();
    if (lane == 0 && nst > 0) {
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&smap) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&zmap) : "memory");
      trace_stamp(a, 1);
      int s = 0, ph = 0, issued = 0;
      for (int pass = 0; pass < 2; ++pass) {
        int tile = (pass == 0 && rot > 0) ? tile0 + 1 : tile0;
        int stg = (pass == 0 && rot > 0) ? 0 : stg0;
        const int count = pass == 0 ? nst - rot : rot;
        for (int j = 0; j < count; ++j, ++issued) {
          if (issued >= ring) mbar_wait(&empty_bar[s], ph ^ 1);
          unsigned char* st = stage_base + s * Cfg::kStageBytes;
          const int n0 = tile * NT, blk = stg * WK;
          mbar_arrive_expect_tx(&full_bar[s], Cfg::kTxBytes);
#pragma unroll
          for (int cc = 0; cc < WC; ++cc) tma_load_2d(st + cc * Cfg::kBoxBytes, &wmap, n0 + 32 * cc, blk * 16, &full_bar[s], policy);
          tma_load_2d(st + Cfg::kWeights, &smap, n0, blk * GPB, &full_bar[s], policy);
          tma_load_2d(st + Cfg::kWeights + Cfg::kScales, &zmap, n0 >> 3, blk * GPB, &full_bar[s], policy);
          if (++stg == S_t) { stg = 0; ++tile; }
          if (++s == ring) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
    return;
  }

  // =========================== consumers ===========================
  const int r = lane & 3, c8 = lane >> 2;
  const int wc = warp & (WC - 1), wk = (warp / WC) & (WK - 1);
  const W4Lane2<MT> L = make_w4_lane2<MT, UPG, WC>(lane, wc, wk, a.M, pitch, a.zero_bias);
  const __half* aptr = act_sm + wk * 128 + w4_lane_row<UPG>(lane) * 8;     // + chunk * 256 per stage
  const int zt_group_bytes = a.M * 16;
  const unsigned char* zt_w = reinterpret_cast<const unsigned char*>(zt_sm) + wk * GPB * zt_group_bytes;   // + chunk * WK * GPB groups
  const int nout = a.M * NT;
  float tot[2 * MT][4];
#pragma unroll
  for (int v = 0; v < 2 * MT; ++v)
#pragma unroll
    for (int i = 0; i < 4; ++i) tot[v][i] = 0.f;

  griddep_wait();                                   // the activations are the only data produced by the previous kernel
  if (tid == 0) trace_stamp(a, 2);
  {
    // activation chunk ci covers k in [256 * stage(ci), +256), zero padded past K; next to it,
    // per scale group and batch row, sum_k a_k / 64 as an fp16 (hi, lo) pair for the zero-point MMA.
    // Four independent loads per thread are in flight at a time (the L2 round trip is what costs).
    // No divisions in here: the loop is instruction bound, not latency bound (tools/trace.py).
    const int nchunks = act_abs ? S_t : nst;
    const int row_vecs = nchunks * 32;              // 8-half vectors per staged row (a multiple of the warp size)
    const int vecs_valid = a.K >> 3;
    const int stage_add = act_abs ? 0 : stg0;
    constexpr int kBatch = 4;
    for (int m = 0; m < a.M; ++m) {
      const uint4* arow = reinterpret_cast<const uint4*>(a.a + (size_t)m * a.K);
      __half* srow = act_sm + (size_t)m * pitch;
      for (int v0 = 0; v0 < row_vecs; v0 += kBatch * kConsumerThreads) {
        uint4 val[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const int v = v0 + b * kConsumerThreads + tid;        // vector inside the staged row
          int stage = stage_add + (v >> 5);
          if (stage >= S_t) stage -= S_t;                       // relative chunks wrap into the next tile at most once
          const int gv = stage * 32 + (v & 31);                 // vector inside the activation row
          val[b] = make_uint4(0, 0, 0, 0);
          if (v < row_vecs && gv < vecs_valid && a.debug_skip != 2) val[b] = __ldg(arow + gv);
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const int v = v0 + b * kConsumerThreads + tid;
          const bool ok = v < row_vecs;                         // warp-uniform
          if (!ok) continue;
          *reinterpret_cast<uint4*>(srow + v * 8) = permute_act8_v2(val[b]);
          const float2 f0 = __half22float2(u2h2(val[b].x)), f1 = __half22float2(u2h2(val[b].y));
          const float2 f2 = __half22float2(u2h2(val[b].z)), f3 = __half22float2(u2h2(val[b].w));
          float sum = ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
#pragma unroll
          for (int o = 1; o < 4 * UPG; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          if (ok && (lane & (4 * UPG - 1)) == 0) {
            const float q64 = sum * 0.015625f;
            const __half hi = __float2half_rn(q64);
            const __half lo = __float2half_rn(q64 - __half2float(hi));
            *reinterpret_cast<uint4*>(zt_sm + ((size_t)(v / (4 * UPG)) * a.M + m) * 4) =
                make_uint4((uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16), 0u, 0u, 0u);
          }
        }
      }
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
  if (tid == 0) trace_stamp(a, 3);

  auto store_tile = [&](int tl, int o, float v) {
    const int m = o / NT, col = o - m * NT;
    const int n = tl * NT + col;
    if (n < a.N) {
      const __half h = __float2half_rn(v);
      const size_t off = (size_t)m * a.ldo + a.col_offset + n;
      a.out[0][off] = h;
      for (int p = 1; p < a.world; ++p) a.out[p][off] = h;   // fused all-gather: NVLink peer stores
    }
  };

  int s = 0, ph = 0;                                // ring slot of the next unit and the parity of the slot's fill count
  int first_tile = -1;                              // tile whose fix-up is deferred to the end
  const long long loop0 = a.trace ? clock64() : 0;

  for (int pass = 0; pass < 2; ++pass) {
    int it = pass == 0 ? rot : 0;                   // next unit (relative to s_lo)
    const int pass_end = pass == 0 ? nst : rot;
    int tile = (pass == 0 && rot > 0) ? tile0 + 1 : tile0;
    while (it < pass_end) {
      const int seg_begin = it;
      const int stg_first = seg_begin == 0 ? stg0 : 0;
      const int seg_len = min(S_t - stg_first, pass_end - seg_begin);
      const int seg_end = seg_begin + seg_len;
      const int ci_off = act_abs ? stg_first - seg_begin : 0;     // activation chunk of unit it = it + ci_off
      for (; it < seg_end; ++it) {
        mbar_wait(&full_bar[s], ph);
        if (a.trace && tid == 0 && it == rot) trace_stamp(a, 4);
        const unsigned char* st = stage_base + s * Cfg::kStageBytes;
        const int ci = it + ci_off;
        // blocks past K inside the last stage are all-zero weights and scales (TMA zero fill)
        if (a.debug_skip != 1) w4_consume_block_v2<MT, UPG, WC>(st, aptr + ci * (WK * 128), zt_w + ci * (WK * GPB) * zt_group_bytes, zt_group_bytes, L, tot);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
        if (++s == ring) { s = 0; ph ^= 1; }
      }
      // ---- end of the segment [stg_first, stg_first + seg_len) of `tile`: K-slices -> one fp32 partial tile
#pragma unroll
      for (int tt = 0; tt < 2; ++tt)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int m = 8 * mt + 2 * r + (i & 1);
            const int col = 32 * wc + 4 * c8 + 2 * tt + (i >> 1);
            if (m < a.M) red_sm[(wk * a.M + m) * NT + col] = tot[tt * MT + mt][i] * 16777216.f;
            tot[tt * MT + mt][i] = 0.f;
          }
      asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
      const bool is_end = (stg_first + seg_len == S_t);
      const bool whole = (stg_first == 0 && is_end);
      for (int o = tid; o < nout; o += kConsumerThreads) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < WK; ++w) v += red_sm[w * nout + o];
        if (whole) store_tile(tile, o, v);
        else if (is_end) first_sm[o] = v;                          // finisher: add the other CTAs' parts at the end
        else a.sk_partials[(size_t)c * nout + o] = v;              // contributor: publish
      }
      if (!whole && is_end) first_tile = tile;
      if (!whole && !is_end) {
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
        if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.sk_flags + c), "r"(1u) : "memory");
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");   // red_sm is reused by the next segment
      ++tile;
    }
  }
  if (tid == 0 && a.trace) {
    trace_value(a, 8, (unsigned long long)(clock64() - loop0));
    trace_value(a, 11, (unsigned long long)nst);
    trace_stamp(a, 5);
  }

  if (first_tile >= 0) {
    // contributors = the CTAs before this one whose ranges touch the tile, in order
    const long long x = (long long)first_tile * S_t;                // first stage of the tile
    const int c_first = (int)(((x + 1) * G + W - 1) / W) - 1;
    if (c_first + tid < c) {
      unsigned int f;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(a.sk_flags + c_first + tid) : "memory");
      } while (f == 0u);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
    if (tid == 0) trace_stamp(a, 6);
    for (int o = tid; o < nout; o += kConsumerThreads) {
      float v = first_sm[o];
      // four contributors at a time: independent loads, summed in CTA order
      for (int cc = c_first; cc < c; cc += 4) {
        float p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = (cc + j < c) ? __ldcg(a.sk_partials + (size_t)(cc + j) * nout + o) : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) v += p[j];
      }
      store_tile(first_tile, o, v);
    }
    if (c_first + tid < c) a.sk_flags[c_first + tid] = 0u;          // leave the workspace clean
  }
  if (tid == 0) trace_stamp(a, 7);
}

// (A tcgen05 / TMEM family -- converter warps -> TMEM A tiles -> tcgen05.mma -> tcgen05.ld epilogue -- existed in round 1
// and was removed in round 2: the converter warps pay the same unpack as the mma.sync kernels, and it measured behind them
// at every M, profiles/r01_v4_*tcgen05*, profiles/r01_v6_skinny_m_crossover.log.)

// ------------------------------------------------------------------------------------------------
// generic: any bits / groupsize / M / N.  One column per thread, 32 columns x 8 K-slices per CTA.
// A 32-k unit of a column is exactly BITS packed words: they are loaded together (and the next unit's before this one
// is unpacked -- the loop was bound by one dependent global load per word), every field position is a compile-time
// constant (shift + mask, a funnel shift where a field straddles two words), the activations come eight k at a time as
// one 16-byte load per batch row (every lane the same address), and the group boundary is tracked as a k value: no
// division and no run-time indexed word array in the loop.  8-bit 4096 x 11008: 238 -> 128 (vector activations) -> see
// profiles/r02_pw8_generic_kernel.log.
// (scale and effective zero point of (group g, column n); kept out of line: it is reached once per group from 32
// unrolled call sites)
template <int BITS>
__device__ __noinline__ float2 generic_group_params(const __half* scales, const uint32_t* qzeros, int g, int n, int N, int zwords, int zero_bias) {
  constexpr uint32_t mask = (1u << BITS) - 1u;
  const float s = __half2float(scales[(size_t)g * N + n]);
  const int zpos = n * BITS, zi = zpos >> 5, zsh = zpos & 31;
  const uint32_t zlo = __ldg(qzeros + (size_t)g * zwords + zi);
  const uint32_t zhi = (zsh + BITS > 32 && zi + 1 < zwords) ? __ldg(qzeros + (size_t)g * zwords + zi + 1) : 0u;
  return make_float2(s, (float)((__funnelshift_r(zlo, zhi, zsh) & mask) + (uint32_t)zero_bias));
}

// NWARPS K-slices (warps) per CTA: the loop is a chain of dependent latencies, so what counts is warps per SM: 16 per CTA
// when N / 32 CTAs do not give every SM at least two CTAs, 8 otherwise (three CTAs of 8 warps fit an SM, one of 16)
template <int BITS, int MC, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32)
gemv_generic_kernel(const GemvArgs a, int m_base) {
  __shared__ float red[NWARPS][4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const bool valid = n < a.N;
  constexpr uint32_t mask = (1u << BITS) - 1u;
  const int total_units = (a.K + 31) >> 5;
  const int wq = (total_units + NWARPS - 1) / NWARPS;
  const int my0 = min(warp * wq, total_units), my1 = min(my0 + wq, total_units);
  const int mcount = min(MC, a.M - m_base);
  const int groups = (a.K + a.groupsize - 1) / a.groupsize;

  float tot[MC], dsum[MC], asum[MC];      // dsum / asum: sum a_k * w_k and sum a_k inside the current group
#pragma unroll
  for (int m = 0; m < MC; ++m) tot[m] = dsum[m] = asum[m] = 0.f;
  float s = 0.f, z = 0.f;
  const bool vec_ok = (a.K % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.a) & 15u) == 0);
  griddep_wait();
  if (valid && my0 < my1) {
    auto load_unit = [&](int u, uint32_t (&w)[BITS]) {
#pragma unroll
      for (int j = 0; j < BITS; ++j) {
        const int row = u * BITS + j;
        w[j] = (u < my1 && row < a.qrows) ? __ldg(a.qweight + (size_t)row * a.N + n) : 0u;
      }
    };
    int g = (my0 * 32) / a.groupsize;
    int next_gk = (g + 1) * a.groupsize;
    {
      const float2 p = generic_group_params<BITS>(a.scales, a.qzeros, g, n, a.N, a.zwords, a.zero_bias);
      s = p.x; z = p.y;
    }
    uint32_t wnext[BITS];
    load_unit(my0, wnext);
    for (int u = my0; u < my1; ++u) {
      uint32_t w[BITS];
#pragma unroll
      for (int j = 0; j < BITS; ++j) w[j] = wnext[j];
      load_unit(u + 1, wnext);
#pragma unroll
      for (int i8 = 0; i8 < 32; i8 += 8) {
        const int kk = u * 32 + i8;
        if (kk < a.K) {
          // (activations beyond K read as 0: the fields there contribute nothing, whatever the packer left in them)
          float av[MC][8];
#pragma unroll
          for (int m = 0; m < MC; ++m) {
            const __half* arow = a.a + (size_t)(m_base + (m < mcount ? m : 0)) * a.K + kk;
            if (vec_ok) {
              const uint4 v = __ldg(reinterpret_cast<const uint4*>(arow));
              const float2 f0 = __half22float2(u2h2(v.x)), f1 = __half22float2(u2h2(v.y));
              const float2 f2 = __half22float2(u2h2(v.z)), f3 = __half22float2(u2h2(v.w));
              av[m][0] = f0.x; av[m][1] = f0.y; av[m][2] = f1.x; av[m][3] = f1.y;
              av[m][4] = f2.x; av[m][5] = f2.y; av[m][6] = f3.x; av[m][7] = f3.y;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) av[m][j] = kk + j < a.K ? __half2float(arow[j]) : 0.f;
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = kk + j;
            const int pos = (i8 + j) * BITS, wi = pos >> 5, sh = pos & 31;      // compile-time after unrolling
            uint32_t f;
            if (sh + BITS <= 32) f = (w[wi] >> sh) & mask;
            else f = __funnelshift_r(w[wi], w[wi + 1 < BITS ? wi + 1 : wi], sh) & mask;
            const float wv = (float)f;
            if (k == next_gk) {
#pragma unroll
              for (int m = 0; m < MC; ++m) { tot[m] = fmaf(s, dsum[m] - z * asum[m], tot[m]); dsum[m] = 0.f; asum[m] = 0.f; }
              g = min(g + 1, groups - 1);
              next_gk += a.groupsize;
              const float2 p = generic_group_params<BITS>(a.scales, a.qzeros, g, n, a.N, a.zwords, a.zero_bias);
              s = p.x; z = p.y;
            }
#pragma unroll
            for (int m = 0; m < MC; ++m) {
              dsum[m] = fmaf(av[m][j], wv, dsum[m]);
              asum[m] += av[m][j];
            }
          }
        }
      }
    }
#pragma unroll
    for (int m = 0; m < MC; ++m) tot[m] = fmaf(s, dsum[m] - z * asum[m], tot[m]);
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) red[warp][m][lane] = m < MC ? tot[m < MC ? m : 0] : 0.f;
  __syncthreads();
  if (warp < mcount && valid) {
    const int m = warp;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) v += red[w][m][lane];
    const __half h = __float2half_rn(v);
    const size_t off = (size_t)(m_base + m) * a.ldo + a.col_offset + n;
    for (int p = 0; p < a.world; ++p) a.out[p][off] = h;
  }
}

// ------------------------------------------------------------------------------------------------
// host side

int device_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool gemv_w4_supported(const GemvArgs& a) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(a.a) | reinterpret_cast<uintptr_t>(a.qweight) |
                       reinterpret_cast<uintptr_t>(a.scales) | reinterpret_cast<uintptr_t>(a.qzeros);
  const bool group_ok = (a.groupsize == 32) || (a.groupsize == 64) || (a.groupsize == 128);
  // N % 32 == 0: TMA global strides (N*4, N*2, N/2 bytes) must be multiples of 16
  return a.bits == 4 && group_ok && a.K % 128 == 0 && a.N % 32 == 0 && (al & 15u) == 0 && a.M >= 1;
}

// Policy switches (XBIT_GEMV_* / XBIT_W4P_* / XBIT_DQ_*): a fixed table of atomics, initialised ONCE from the
// environment and changed at run time only through xbit_set_option (tests, tools/*.py).  No getenv on the launch path,
// no unsynchronised lazy statics.
namespace {
struct Option {
  const char* name;
  std::atomic<int> value;
  std::atomic<bool> set;
};
Option g_options[] = {
    {"XBIT_GEMV_FAMILY", {0}, {false}},  {"XBIT_GEMV_STREAMK", {0}, {false}}, {"XBIT_GEMV_WC", {0}, {false}},
    {"XBIT_GEMV_SPLITS", {0}, {false}},  {"XBIT_GEMV_RING", {0}, {false}},    {"XBIT_W4P_WARPS", {0}, {false}},
    {"XBIT_W4P_I8", {0}, {false}},       {"XBIT_W4P_FINE", {0}, {false}},     {"XBIT_W4P_GRID", {0}, {false}},
    {"XBIT_W4P_RING", {0}, {false}},     {"XBIT_W4P_ALLWAIT", {0}, {false}},
    {"XBIT_W4P_DELAY", {0}, {false}},    {"XBIT_DQ_SMEM_KB", {0}, {false}},
    {"XBIT_GEMV_DEBUG_SKIP", {0}, {false}}, {"XBIT_LL_PERSIST", {0}, {false}},
};
std::once_flag g_options_once;
void load_options_from_env() {
  for (Option& o : g_options) {
    const char* v = getenv(o.name);
    if (v && *v) {
      o.value.store(atoi(v));
      o.set.store(true);
    }
  }
}
Option* find_option(const char* name) {
  std::call_once(g_options_once, load_options_from_env);
  for (Option& o : g_options)
    if (strcmp(o.name, name) == 0) return &o;
  return nullptr;
}
}  // namespace

int env_int(const char* name, int dflt) {
  const Option* o = find_option(name);
  return (o && o->set.load(std::memory_order_relaxed)) ? o->value.load(std::memory_order_relaxed) : dflt;
}

// value == INT_MIN: back to "unset" (the built-in policy).  false: unknown option name.
bool set_option(const char* name, int value) {
  Option* o = find_option(name);
  if (!o) return false;
  if (value == INT_MIN) o->set.store(false);
  else {
    o->value.store(value);
    o->set.store(true);
  }
  return true;
}


struct W4Plan {
  int wc, splits, blocks_per_split;
  size_t smem;
  dim3 grid;
};

static size_t w4_smem_bytes(int upg, int wc, int mt, int m, int blocks_per_split, int splits, int ring) {
  const int wk = 8 / wc, nt = 32 * wc, gpb = 4 / upg;
  const size_t stage = ((size_t)(wc * wk * 16 * 128) + (size_t)wk * gpb * nt * 2 + (size_t)wk * gpb * (nt / 8) * 4 + 1023) / 1024 * 1024;
  return 1024 /* alignment slack */ + (size_t)ring * stage + 128                     // ring + mbarriers
         + (size_t)m * (blocks_per_split * 128 + 8) * sizeof(__half)                  // act_sm
         + (size_t)(mt > 0 ? blocks_per_split * gpb * (m * 4 > 8 * mt ? m * 4 : 8 * mt) : 0) * sizeof(float)   // asum_sm (v1) / group-sum table (v2)
         + (size_t)wk * m * nt * sizeof(float)                                        // red_sm
         + (size_t)(splits > 1 ? splits : 0) * m * nt * sizeof(float);                // clus_sm
}

// developer knobs of one launch: tools/sweep.py (skip the math) and tools/trace.py (phase stamps).  Only in the
// -DXBIT_DEVTOOLS build (libxbitops_b200_dev.so, used by tools/); the shipped library has neither.
void apply_debug_knobs(GemvArgs& a) {
  a.debug_skip = 0;
  a.trace = nullptr;
#ifdef XBIT_DEVTOOLS
  a.debug_skip = env_int("XBIT_GEMV_DEBUG_SKIP", 0);
  if (const char* tp = getenv("XBIT_GEMV_TRACE")) {   // device buffer of [launch % 64][1024 CTAs][16] stamps
    static std::atomic<int> launches{0};
    a.trace = reinterpret_cast<unsigned long long*>(strtoull(tp, nullptr, 0)) + (size_t)(launches.fetch_add(1) % 64) * 16384;
  }
#endif
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Host-side cost matters for the per-call (e2e) figure: cuTensorMapEncodeTiled is a pure function of
// its arguments, so the encoded maps are kept in a small cache keyed by all of them.
cudaError_t encode_2d(CUtensorMap* map, CUtensorMapDataType dt, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t row_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle sw, CUtensorMapL2promotion promo) {
  using Key = std::tuple<int, const void*, uint64_t, uint64_t, uint64_t, uint32_t, uint32_t, int, int>;
  static std::mutex mu;
  static auto* cache = new std::map<Key, CUtensorMap>();
  const Key key((int)dt, base, inner, outer, row_bytes, box_inner, box_outer, (int)sw, (int)promo);
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache->find(key);
    if (it != cache->end()) {
      memcpy(map, &it->second, sizeof(CUtensorMap));
      return cudaSuccess;
    }
  }
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return cudaErrorNotSupported;
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {row_bytes};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  std::lock_guard<std::mutex> lock(mu);
  if (cache->size() >= 8192) cache->clear();
  (*cache)[key] = *map;
  return cudaSuccess;
}

// opt in to large dynamic shared memory once per (device, kernel); not a stream operation
cudaError_t ensure_max_dyn_smem(const void* kern) {
  static std::mutex mu;
  static auto* done = new std::set<std::pair<int, const void*>>();
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  if (done->count({dev, kern})) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
  if (e == cudaSuccess) done->insert({dev, kern});
  return e;
}

using W4Kernel = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemvArgs);

static int upg_of(int groupsize) { return groupsize == 32 ? 1 : (groupsize == 64 ? 2 : 4); }

static cudaError_t launch_w4(W4Kernel kern, const GemvArgs& a, const W4Plan& p, int upg, cudaStream_t stream) {
  const int wk = 8 / p.wc, nt = 32 * p.wc, gpb = 4 / upg;
  alignas(64) CUtensorMap wmap, smap, zmap;
  // qweight [qrows, N] u32: box = wk*16 rows x 32 columns (128 B), 128-byte swizzle
  cudaError_t e = encode_2d(&wmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, a.qweight, (uint64_t)a.N, (uint64_t)a.qrows,
                            (uint64_t)a.N * 4, 32, (uint32_t)(wk * 16), CU_TENSOR_MAP_SWIZZLE_128B);
  if (e != cudaSuccess) return e;
  // scales [groups, N] f16 (moved as u16): box = wk*gpb rows x nt columns, dense
  e = encode_2d(&smap, CU_TENSOR_MAP_DATA_TYPE_UINT16, a.scales, (uint64_t)a.N, (uint64_t)a.groups, (uint64_t)a.N * 2,
                (uint32_t)nt, (uint32_t)(wk * gpb), CU_TENSOR_MAP_SWIZZLE_NONE);
  if (e != cudaSuccess) return e;
  // qzeros [groups, N/8] u32: box = wk*gpb rows x nt/8 words, dense
  e = encode_2d(&zmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, a.qzeros, (uint64_t)a.zwords, (uint64_t)a.groups,
                (uint64_t)a.zwords * 4, (uint32_t)(nt / 8), (uint32_t)(wk * gpb), CU_TENSOR_MAP_SWIZZLE_NONE);
  if (e != cudaSuccess) return e;
  e = ensure_max_dyn_smem(reinterpret_cast<const void*>(kern));
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = p.grid;
  cfg.blockDim = dim3(kW4Threads, 1, 1);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  if (a.splits > 1) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = 1;
    attrs[na].val.clusterDim.y = (unsigned)a.splits;
    attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, wmap, smap, zmap, a);
}


template <int MT, int HYB>
static W4Kernel pick_w4_kernel(int upg, int wc) {
#define XBIT_W4_CASE(UPG_, WC_) if (upg == UPG_ && wc == WC_) return gemv_w4_kernel<MT, UPG_, WC_, HYB>;
  XBIT_W4_CASE(1, 1) XBIT_W4_CASE(1, 2) XBIT_W4_CASE(1, 4) XBIT_W4_CASE(1, 8)
  XBIT_W4_CASE(2, 1) XBIT_W4_CASE(2, 2) XBIT_W4_CASE(2, 4) XBIT_W4_CASE(2, 8)
  XBIT_W4_CASE(4, 1) XBIT_W4_CASE(4, 2) XBIT_W4_CASE(4, 4) XBIT_W4_CASE(4, 8)
#undef XBIT_W4_CASE
  return nullptr;
}

// CTA slots the device really offers to clusters of `splits` CTAs of this kernel (cached per kernel, cluster size and
// shared-memory size): clusters are placed inside one GPC, so the answer is below 2 * #SMs for awkward sizes.
static long long max_active_cluster_ctas(W4Kernel kern, int splits, size_t smem) {
  using Key = std::tuple<int, const void*, int, size_t>;
  static std::mutex mu;
  static auto* cache = new std::map<Key, long long>();
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  const Key key(dev, reinterpret_cast<const void*>(kern), splits, smem);
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache->find(key);
    if (it != cache->end()) return it->second;
  }
  long long result = -1;
  if (ensure_max_dyn_smem(reinterpret_cast<const void*>(kern)) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1, (unsigned)splits, 1);
    cfg.blockDim = dim3(kW4Threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = (unsigned)splits;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&clusters, reinterpret_cast<const void*>(kern), &cfg) == cudaSuccess) result = (long long)clusters * splits;
    else cudaGetLastError();
  }
  std::lock_guard<std::mutex> lock(mu);
  (*cache)[key] = result;
  return result;
}

// Decomposition (measured: profiles/r01_v6_planner_sweep.log).  The consumer instruction stream
// bounds an SM and a lone 8-warp CTA does not saturate it, so what counts is how close the CTA count
// comes to two per SM in ONE wave; at equal counts fewer K splits win (no DSMEM reduction), tiles of
// 128 columns win for long K slices (512-byte DRAM runs), and a CTA wants at least 1024 k (4 stages).
// Ring depth: 3 stages when a CTA streams only a few (a third CTA slot per SM stays free, so more of
// the NEXT launch's CTAs are resident early and prefetch), up to 5 for long streams.
struct W4PlanEntry {
  W4Plan plan;
  int ring;
  double score;
  bool ok;
};

static bool plan_w4(GemvArgs& a, int mt, int upg, W4Plan& p, double* score_out = nullptr) {
  const int sms = device_sm_count();
  const int nblocks = a.K / 128;
  const int cap = 2 * sms;
  const int env_wc = env_int("XBIT_GEMV_WC", 0);     // tuning knobs for tools/sweep.py
  const int env_splits = env_int("XBIT_GEMV_SPLITS", 0);
  const int env_ring = env_int("XBIT_GEMV_RING", 0);
  // the decomposition is a pure function of these: computed once per shape
  using Key = std::tuple<int, int, int, int, int, int, int, int, int, int>;
  static std::mutex mu;
  static auto* cache = new std::map<Key, W4PlanEntry>();
  int dev = 0;
  cudaGetDevice(&dev);
  const Key key(dev, a.M, a.K, a.N, upg, mt, env_wc, env_splits, env_ring, a.ll_out);
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache->find(key);
    if (it != cache->end()) {
      const W4PlanEntry& e = it->second;
      if (!e.ok) return false;
      p = e.plan;
      a.ring = e.ring;
      a.splits = p.splits;
      a.units_per_split = p.blocks_per_split;
      if (score_out) *score_out = e.score;
      return true;
    }
  }
  double best = -1.0;
  bool found = false;
  for (int wc = 1; wc <= 8; wc *= 2) {
    if (env_wc ? wc != env_wc : (wc == 1 && a.N >= 64)) continue;
    if (a.N < 32 * wc && wc > 1) continue;
    const int tiles = (a.N + 32 * wc - 1) / (32 * wc);
    const int wk = 8 / wc;
    for (int splits = 1; splits <= 8; ++splits) {
      // cluster sizes 1, 2, 3, 4, 8: 3 fills shapes like N = 11008 (86 tiles -> 258 CTAs) that powers of two leave
      // at 58 %; 5, 6 and 7 were measured slower than their CTA count suggests (profiles/r01_v6_planner_sweep_odd_splits.log)
      if (env_splits ? splits != env_splits : (splits >= 5 && splits <= 7)) continue;
      const int bps = (nblocks + splits - 1) / splits;
      if (splits > 1 && (splits - 1) * bps >= nblocks) continue;       // an empty split
      const int stages = (bps + wk - 1) / wk;
      const long long ctas = (long long)tiles * splits;
      const int ring = env_ring >= 2 && env_ring <= kMaxStages ? env_ring : (stages >= 24 ? 5 : (stages >= 12 ? 4 : 3));
      const size_t smem = w4_smem_bytes(upg, wc, mt, a.M, bps, splits, ring);
      if (smem > kMaxDynSmem) continue;                                 // K slice too long for the staged activations at this M
      // staged activations grow with M: above half an SM only one CTA is resident and the SM runs at
      // about half its rate (a lone 8-warp CTA does not saturate the issue slots).  Clusters must be
      // co-scheduled inside a GPC: the driver's occupancy query gives the real number of CTA slots
      // (odd cluster sizes fragment the GPCs: e.g. 5-CTA clusters of 7168x7168 ran in two waves).
      const bool two_per_sm = smem <= 113 * 1024;
      long long slots = two_per_sm ? cap : cap / 2;
      if (splits > 1) {
        const W4Kernel kern = mt == 0 ? pick_w4_kernel<0, 0>(upg, wc) : (mt == 1 ? pick_w4_kernel<1, 0>(upg, wc) : pick_w4_kernel<2, 0>(upg, wc));
        const long long real = max_active_cluster_ctas(kern, splits, smem);
        if (real > 0 && real < slots) slots = real;
      }
      const long long waves = (ctas + slots - 1) / slots;
      double score = (double)ctas / (double)(waves * slots) * ((double)slots / (double)cap);   // fill of the whole machine
      if (!two_per_sm) score *= 1.2;                                                            // (cap/2 slots already halve it)
      if (waves > 1) score *= 0.75;                                      // every extra wave serialises a prologue and an epilogue
      // chains of LL-fed calls overlap: the next call's clusters must find 8 free slots in one GPC while this call still
      // runs, so 8-CTA clusters that fill the machine start late (4 GPUs, 8192x8192: 8.0 vs 6.0 us per call with 4-CTA clusters)
      if (a.ll_out && splits == 8 && tiles * 4 >= sms - 20) score *= 0.5;
      if (bps * 128 < 1024) score *= 0.8;
      score *= 1.0 - 0.02 * log2((double)splits);                        // at equal fill fewer splits win (smaller DSMEM reduction)
      score *= wc == 4 ? 1.0 : (wc == 2 ? (bps * 128 >= 4096 ? 0.94 : 0.99) : 0.98);
      if (score > best) {
        best = score;
        found = true;
        p.wc = wc;
        p.splits = splits;
        p.blocks_per_split = bps;
        p.smem = smem;
        p.grid = dim3((unsigned)tiles, (unsigned)splits, 1);
        a.ring = ring;
      }
    }
  }
  {
    std::lock_guard<std::mutex> lock(mu);
    if (cache->size() >= 4096) cache->clear();
    (*cache)[key] = W4PlanEntry{p, a.ring, best, found};
  }
  if (!found) return false;
  if (score_out) *score_out = best;
  a.splits = p.splits;
  a.units_per_split = p.blocks_per_split;
  return true;
}


cudaError_t launch_gemv_w4_simt(GemvArgs a, cudaStream_t stream) {
  if (a.M != 1) return cudaErrorInvalidValue;
  const int upg = upg_of(a.groupsize);
  W4Plan p;
  if (!plan_w4(a, 0, upg, p)) return cudaErrorInvalidValue;
  apply_debug_knobs(a);
  W4Kernel k = pick_w4_kernel<0, 0>(upg, p.wc);
  return k ? launch_w4(k, a, p, upg, stream) : cudaErrorInvalidValue;
}

// true when the cluster split-K kernel has a decomposition for this problem (the staged activations of a large
// M * K may not fit in shared memory for any K split)
bool gemv_w4_mma_has_plan(GemvArgs a) {
  if (!gemv_w4_supported(a) || a.M > 16) return false;
  W4Plan p;
  return plan_w4(a, a.M <= 8 ? 1 : 2, upg_of(a.groupsize), p);
}

cudaError_t launch_gemv_w4_mma(GemvArgs a, cudaStream_t stream) {
  const int mt = a.M <= 8 ? 1 : 2;
  const int upg = upg_of(a.groupsize);
  W4Plan p;
  if (!plan_w4(a, mt, upg, p)) return cudaErrorInvalidValue;
  apply_debug_knobs(a);
  // (the FHFMA / HMMA hybrid of w4_consume_block -- template parameter HYB -- was measured on B200 and is no
  // longer instantiated: it does not reduce issue slots, which is what bounds the loop)
  W4Kernel k = mt == 1 ? pick_w4_kernel<1, 0>(upg, p.wc) : pick_w4_kernel<2, 0>(upg, p.wc);
  return k ? launch_w4(k, a, p, upg, stream) : cudaErrorInvalidValue;
}

// ---- persistent stream-K launch
size_t gemv_w4_streamk_workspace_bytes(int M) {
  const int sms = device_sm_count();
  return (size_t)sms * ((size_t)M * 128 * sizeof(float)) + (((size_t)sms * sizeof(unsigned int) + 255) / 256) * 256 + 256;
}

struct SkPlan {
  int grid, ring, act_units, act_abs, stages_per_tile;
  long long total;
  size_t smem;
};

static size_t streamk_smem(int upg, int mt, int m, int act_units, int ring) {
  const int gpb = 4 / upg, wk = 2, nt = 128;
  (void)mt;
  const size_t stage = ((size_t)(4 * wk * 16 * 128) + (size_t)wk * gpb * nt * 2 + (size_t)wk * gpb * (nt / 8) * 4 + 1023) / 1024 * 1024;
  return 1024 + (size_t)ring * stage + 128 + (size_t)m * (act_units * 256 + 8) * sizeof(__half)
         + (size_t)act_units * wk * gpb * m * 16 + (size_t)(wk + 1) * m * nt * sizeof(float);
}

static bool plan_streamk(const GemvArgs& a, int mt, SkPlan& p) {
  const int sms = device_sm_count();
  const int upg = upg_of(a.groupsize);
  const int tiles = (a.N + 127) / 128;
  p.stages_per_tile = (a.K / 128 + 1) / 2;
  p.total = (long long)tiles * p.stages_per_tile;
  p.grid = p.total < sms ? (int)p.total : sms;
  const int per_cta = (int)((p.total + p.grid - 1) / p.grid);
  p.act_abs = per_cta >= p.stages_per_tile;
  p.act_units = p.act_abs ? p.stages_per_tile : per_cta;
  // Half an SM (so that the next launch is resident while this one computes) when that still
  // leaves a ring of 4 stages; otherwise whatever fits, up to 6
  const size_t half = 113 * 1024;
  int ring = env_int("XBIT_GEMV_RING", 0);
  if (ring < 2 || ring > kSkMaxRing) {
    ring = 6;
    while (ring > 4 && streamk_smem(upg, mt, a.M, p.act_units, ring) > half) --ring;
    while (ring > 2 && streamk_smem(upg, mt, a.M, p.act_units, ring) > kMaxDynSmem) --ring;
  }
  p.ring = ring;
  p.smem = streamk_smem(upg, mt, a.M, p.act_units, ring);
  return p.smem <= kMaxDynSmem;
}

bool gemv_w4_streamk_applicable(const GemvArgs& a, int family) {
  if (!gemv_w4_supported(a)) return false;
  if (family != XBIT_GEMV_MMA || a.M > 16) return false;      // tensor-core block math only
  const int mt = a.M <= 8 ? 1 : 2;
  SkPlan p;
  return plan_streamk(a, mt, p);
}

// AUTO policy (profiles/r01_v6_cluster_vs_streamk.log): the cluster split-K kernel wins wherever its grid
// fills the CTA slots of one wave reasonably; when it cannot (e.g. N = 5120: 160 of 296 slots) and the
// matrix is large enough to amortise the stream-K fix-up, the balanced persistent schedule is faster.
bool gemv_w4_prefers_streamk(const GemvArgs& a, int family) {
  if (!gemv_w4_streamk_applicable(a, family)) return false;
  GemvArgs probe = a;
  W4Plan p;
  double score = 1.0;
  if (!plan_w4(probe, a.M <= 8 ? 1 : 2, upg_of(a.groupsize), p, &score)) return true;   // the cluster kernel cannot stage this K at this M
  const double weight_bytes = (double)a.K * a.N * 0.5;
  return score < 0.56 && weight_bytes >= 32e6;
}

template <int MT>
static W4Kernel pick_sk_kernel(int upg) {
  if (upg == 1) return gemv_w4_streamk_kernel<MT, 1>;
  if (upg == 2) return gemv_w4_streamk_kernel<MT, 2>;
  return gemv_w4_streamk_kernel<MT, 4>;
}

cudaError_t launch_gemv_w4_streamk(GemvArgs a, int family, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const int sms = device_sm_count();
  const int upg = upg_of(a.groupsize), gpb = 4 / upg;
  if (family != XBIT_GEMV_MMA) return cudaErrorInvalidValue;
  const int mt = a.M <= 8 ? 1 : 2;
  if (workspace_bytes < gemv_w4_streamk_workspace_bytes(a.M) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return cudaErrorInvalidValue;
  SkPlan p;
  if (!plan_streamk(a, mt, p)) return cudaErrorInvalidValue;
  a.sk_stages_per_tile = p.stages_per_tile;
  a.sk_total_stages = (int)p.total;
  a.sk_ring = p.ring;
  a.sk_act_units = p.act_units;
  a.sk_act_abs = p.act_abs;
  a.sk_flags = reinterpret_cast<unsigned int*>(workspace);
  a.sk_partials = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + (((size_t)sms * sizeof(unsigned int) + 255) / 256) * 256);
  a.splits = 1;
  apply_debug_knobs(a);
  W4Kernel kern = mt == 1 ? pick_sk_kernel<1>(upg) : pick_sk_kernel<2>(upg);

  const int wk = 2, nt = 128;
  alignas(64) CUtensorMap wmap, smap, zmap;
  cudaError_t e = encode_2d(&wmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, a.qweight, (uint64_t)a.N, (uint64_t)a.qrows,
                            (uint64_t)a.N * 4, 32, (uint32_t)(wk * 16), CU_TENSOR_MAP_SWIZZLE_128B);
  if (e != cudaSuccess) return e;
  e = encode_2d(&smap, CU_TENSOR_MAP_DATA_TYPE_UINT16, a.scales, (uint64_t)a.N, (uint64_t)a.groups, (uint64_t)a.N * 2,
                (uint32_t)nt, (uint32_t)(wk * gpb), CU_TENSOR_MAP_SWIZZLE_NONE);
  if (e != cudaSuccess) return e;
  e = encode_2d(&zmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, a.qzeros, (uint64_t)a.zwords, (uint64_t)a.groups,
                (uint64_t)a.zwords * 4, (uint32_t)(nt / 8), (uint32_t)(wk * gpb), CU_TENSOR_MAP_SWIZZLE_NONE);
  if (e != cudaSuccess) return e;
  e = ensure_max_dyn_smem(reinterpret_cast<const void*>(kern));
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)p.grid, 1, 1);
  cfg.blockDim = dim3(kW4Threads, 1, 1);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, wmap, smap, zmap, a);
}

cudaError_t launch_gemv_generic(GemvArgs a, cudaStream_t stream) {
  const unsigned grid = (unsigned)((a.N + 31) / 32);
  for (int m0 = 0; m0 < a.M; m0 += 4) {
    const bool one = a.M - m0 == 1;
    const bool wide = grid < 2u * (unsigned)device_sm_count();
    switch (a.bits) {
#define XBIT_GENERIC_CASE(B_) case B_:                                                              \
      if (one && wide) gemv_generic_kernel<B_, 1, 16><<<grid, 512, 0, stream>>>(a, m0);              \
      else if (one) gemv_generic_kernel<B_, 1, 8><<<grid, 256, 0, stream>>>(a, m0);                  \
      else if (wide) gemv_generic_kernel<B_, 4, 16><<<grid, 512, 0, stream>>>(a, m0);                \
      else gemv_generic_kernel<B_, 4, 8><<<grid, 256, 0, stream>>>(a, m0);                           \
      break;
      XBIT_GENERIC_CASE(2) XBIT_GENERIC_CASE(3) XBIT_GENERIC_CASE(4) XBIT_GENERIC_CASE(5) XBIT_GENERIC_CASE(6) XBIT_GENERIC_CASE(7) XBIT_GENERIC_CASE(8)
#undef XBIT_GENERIC_CASE
      default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}


// ---- explicit consumer side of the fused completion signal (end of a chain): one warp
__global__ void peers_wait_kernel(const unsigned int* flags, int world, int rank, unsigned int* timeout_flag) {
  griddep_launch_dependents();
  griddep_wait();                                   // this rank's own call has completed: its flag holds the call count
  if (!peers_poll(flags, world, rank, threadIdx.x) && timeout_flag) *timeout_flag = 1u;
}

cudaError_t launch_peers_wait(const unsigned int* flags, int world, int rank, unsigned int* timeout_flag, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1, 1, 1);
  cfg.blockDim = dim3(32, 1, 1);
  cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // resident before the GEMV retires: no launch gap
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, peers_wait_kernel, flags, world, rank, timeout_flag);
}

// ---- LL buffer -> plain fp16: the consumer at the END of a chain.  One CTA; it also advances the chain base.
__global__ void __launch_bounds__(1024)
ll_unpack_kernel(const unsigned long long* __restrict__ ll, uint32_t* __restrict__ out, long long n_pairs,
                 unsigned int* __restrict__ state, int chain_len, unsigned int* timeout_flag) {
  griddep_launch_dependents();
  griddep_wait();                                   // orders this kernel behind the previous chain end (state[2])
  const uint32_t epoch = state[2] + (uint32_t)chain_len;   // the last call's number
  for (long long i = threadIdx.x; i < n_pairs; i += blockDim.x) {
    uint32_t d, f;
    const long long c0 = clock64();
    for (unsigned int n = 1;; ++n) {
      asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(d), "=r"(f) : "l"(ll + i) : "memory");
      if (f == epoch) break;
      if ((n & 1023u) == 0 && clock64() - c0 > kSpinGuardClocks) {
        if (timeout_flag) *timeout_flag = 1u;
        break;
      }
    }
    out[i] = d;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    state[2] = epoch;                               // the next chain numbers its calls from here
    state[1] = epoch;                               // calls completed (informational)
  }
}

cudaError_t launch_ll_unpack(const void* ll_in, void* out_f16, long long n_pairs, unsigned int* state, int chain_len,
                             unsigned int* timeout_flag, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1, 1, 1);
  cfg.blockDim = dim3(1024, 1, 1);
  cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, ll_unpack_kernel, reinterpret_cast<const unsigned long long*>(ll_in),
                            reinterpret_cast<uint32_t*>(out_f16), n_pairs, state, chain_len, timeout_flag);
}

// ---- host-buffer entry: pull activation rows out of page-locked host memory with a kernel instead of a
// copy node, so that the chain  pull -> gemv  stays on programmatic dependent launches (a copy node in
// front of the GEMV costs ~10 us of copy-engine latency per call; measured in bench.py's e2e leg)
__global__ void pull_rows_kernel(const uint4* __restrict__ src_host, uint4* __restrict__ dst, size_t nvec) {
  griddep_launch_dependents();                      // the GEMV behind may start streaming its weights
  // The host rows do not depend on the kernel in front (the caller wrote them before enqueueing the call): the PCIe reads
  // -- about 1.5 us of latency -- are issued BEFORE the wait and only the stores into dst, which the previous GEMV may still
  // be reading, come after it.  Four vectors per thread cover 1 MiB of activations; a longer tail is read after the wait.
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  uint4 pre[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (i0 + j * stride < nvec) pre[j] = src_host[i0 + j * stride];
  griddep_wait();                                   // the previous GEMV has finished reading dst
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (i0 + j * stride < nvec) dst[i0 + j * stride] = pre[j];
  for (size_t i = i0 + 4 * stride; i < nvec; i += stride) dst[i] = src_host[i];
}

cudaError_t launch_pull_rows(const void* src_host_devptr, void* dst, size_t bytes, cudaStream_t stream) {
  const size_t nvec = bytes / 16;
  cudaLaunchConfig_t cfg = {};
  const unsigned blocks = (unsigned)((nvec + 255) / 256);
  cfg.gridDim = dim3(blocks < 1 ? 1 : (blocks > 64 ? 64 : blocks), 1, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, pull_rows_kernel, reinterpret_cast<const uint4*>(src_host_devptr), reinterpret_cast<uint4*>(dst), nvec);
}

}  // namespace xbit
