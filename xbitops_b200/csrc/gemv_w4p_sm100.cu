// gemv_w4p_sm100.cu -- the default A16W4 GEMV / skinny GEMM (M <= 8) for sm_100a: one persistent CTA per SM,
// per-warp TMA rings, block-granular stream-K.
//
// Replaces /root/reference/src/cu/gemv_w4a16_pt.cu: gemv<T> (:35-145) and lauch_Gemv_kernel (:149-173); same
// semantics as gemv_sm100.cu (y[m, n] = RN16( sum_k a[m, k] * DQ[k, n] ), fp32 accumulation).
//
// Why another schedule (round-1 traces, profiles/r01_v6_trace_cluster_kernel.log): a 4096x4096 call spent 2.8 of
// its 4.3 us in serial phases -- CTAs that only became resident when the previous launch's CTAs left, the first
// stage's DRAM round trip behind griddepcontrol.wait, the skew of a 4-CTA cluster reduction -- while the
// persistent stream-K kernel of round 1 (one 8-warp CTA per SM) ran its loop at half rate: 8 warps do not fill
// the issue slots of an SM.  This kernel keeps what worked (TMA feed, 128-byte swizzle, the exact-product
// tensor-core block math, programmatic dependent launch) and changes the decomposition:
//
//   * grid = one CTA per SM, at most half of the SM's shared memory and registers, so the NEXT call's CTA is
//     co-resident on every SM and has its whole ring filled before this call retires;
//   * the work list is every (32-column tile, 128-k block) in tile-major order; CTA c owns the contiguous range
//     [total*c/G, total*(c+1)/G) and its 8 consumer warps own contiguous eighths of that.  A warp takes TWO blocks
//     per step (two independent accumulator sets): 8 warps issue like 16, with one copy of the loop overhead;
//   * every consumer warp has its OWN ring of 2 KiB block slots (weights box 16 rows x 128 B, one scale row box,
//     one zero row box) with full / empty mbarriers; lanes 0..7 of one producer warp each drive one consumer's
//     ring, so there is no CTA-wide lock step, no partly filled stage and no K granularity beyond one block;
//   * split-K is resolved where it happens: warps that share a tile inside a CTA meet through shared memory (the
//     last one to arrive finishes the tile); CTAs that share a tile meet through 8-byte {fp32 partial, flag}
//     slots in the caller's workspace -- the CTA holding the tile's FIRST block is the finisher and processes that
//     tile last, the others hold its later blocks and process them first, so a finisher normally finds the
//     partials waiting.  Sums are taken in warp / CTA order: deterministic, no atomics on data.
//     Without a workspace CTA boundaries are tile-aligned instead and nothing crosses CTAs.
//
// Round-2 additions, all compile-time forms of the one kernel template (gemv_w4p_kernel):
//   I8    integer block math (groupsize 128, M <= 2): the packed words, masked, ARE the u8 operands of
//         mma.sync.m16n8k32; the activations become 24-bit fixed point per scale group in three byte planes;
//   BITS  4, 8 (a packed word is an A register as it is: no unpack instruction) or 2 (four byte masks per word, the
//         activation planes scaled per k mod 4): A16W8 / A16W2 on the same schedule;
//   BF    bf16-native: activations, scales and output bf16, one rounding of the fp32 result;
//   GEN   0 = one matrix per launch (short parameter block, plain loads), 1 = several matrices sharing the activations
//         (xbit_gemv_f16_multi), 2 = also the flag-in-data multi-GPU forms;
//   MINB  2 = half an SM (the next launch co-resident, its share prefetched whole), 1 = a full SM (deeper rings).
// No integer division in the kernel (host-supplied multipliers); single-tile CTAs finish through a named barrier.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#include "gemv_prims.cuh"
#include "unpack.cuh"
#include "xbit_internal.h"
#include "../../include/xbitops_b200.h"

namespace xbit {


constexpr int kPMaxRing = 8;
constexpr long long kPSpinGuardClocks = 60000000000ll;   // ~30 s: see kSpinGuardClocks in gemv_sm100.cu

constexpr int kPMaxProblems = 4;                    // weight matrices per launch (xbit_gemv_f16_multi)

// one weight matrix of a launch: all of them share the activations, M, K, bits and group size
struct W4PProblem {
  const __half* scales;     // [G, N]: copied into the rings with cp.async (64 bytes per block and group: too small for TMA requests)
  const uint32_t* qzeros;   // [G, zwords]
  __half* out[kMaxPeers];   // world output buffers ([M, ldo] each)
  long long ldo, col_offset;
  int N, zwords;
  int total;                // tiles * nb
  int unit;                 // CTA boundaries are multiples of `unit` blocks: 1 (stream-K, needs ws) or nb (tile aligned)
  int uq, ur;               // (total / unit) = uq * grid + ur
  unsigned long long* ws;   // [grid][M][32] {fp32 partial, flag} slots, zero outside a launch
};

// NP = matrices the parameter block has room for: the single-matrix instantiation takes the short form (kernel
// parameters of 0.5 instead of 1.7 KB)
template <int NP>
struct W4PArgsN {
  const __half* a;          // [M, K]
  int world;
  int M, K, zero_bias;
  int nb;                   // 128-k blocks per tile = K / 128
  int ring;                 // slots per ring
  int static_weights;
  int all_wait;             // every consumer warp executes griddepcontrol.wait (comparison knob)
  int prefetch_delay;       // SM clocks the producer waits before its first request (only when it starts ahead of the wait)
  // flag-in-data ("LL") form of the N-split exchange (xbit_gemv_f16_peers_ll, see gemv_sm100.cu): results leave as 8-byte
  // {two fp16 results, call number} stores into every rank's buffer; activations may arrive the same way
  int ll_out;               // prob[].out[] are LL buffers ([M][ldo/2] 8-byte slots); call number = ll_state[2] + ll_chain_index + 1
  int a_is_ll;              // `a` is the LL buffer the previous call filled (K/2 slots per row): no wait for the previous grid
  int ll_chain_index;       // position of this call in its chain (0 = first)
  unsigned int* ll_state;   // local: [2] chain base, [3] timeout flag
  int count;                // weight matrices: the CTA works through its range of each, one after the other
  // division by multiplication (exact for the ranges involved, see range_lo / tile_of): the integer divisions of the
  // CTA prologue sat between CTA start and the first TMA request
  unsigned int g_magic;     // ceil(2^32 / grid):  n / grid = umulhi(n, g_magic)         for n * grid < 2^32
  unsigned int nb_magic;    // ceil(2^nb_shift / nb): j / nb = (j * nb_magic) >> nb_shift  for j * nb < 2^nb_shift
  int nb_shift;
  int debug_skip;
  unsigned long long* trace;
  W4PProblem prob[NP];      // last: the short form is a prefix of the long one
};
using W4PArgs = W4PArgsN<kPMaxProblems>;

// tensor maps of the launch: [problem][0] = box of two blocks, [problem][1] = box of one block
template <int NP>
struct W4PMapsN {
  CUtensorMap m[2 * NP];
};
using W4PMaps = W4PMapsN<kPMaxProblems>;

template <int UPG, int BPS, int BITS = 4>
struct W4PCfg {
  static constexpr int GPB = 4 / UPG;                                   // scale groups per 128-k block
  static constexpr int kRowsPerBlock = 4 * BITS;                        // word-rows of a 128-k block
  static constexpr int kBlockBytes = kRowsPerBlock * 128;               // x 32 columns (4 bits: 2 KiB, 8 bits: 4 KiB)
  static constexpr int kZRow = BITS == 8 ? 32 : 16;                     // bytes of a tile's zero points per group row
  static constexpr int kWSlot = BPS * kBlockBytes;                      // a ring slot holds BPS (1 or 2) consecutive blocks
  static constexpr int kSSlot = BPS * GPB * 64;                         // their scale rows (32 columns x fp16)
  static constexpr int kZSlot = BPS * GPB * kZRow;                      // their zero rows
};

__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
// the mbarrier receives one (pre-counted) arrival when all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct W4PLane {
  uint32_t w_x0;            // byte offset of this lane's 16-byte chunk in a weight slot for even units (odd: ^ kOddXor)
  uint32_t s_off, z_off;    // byte offsets of this lane's 4 scales / 4 zero nibbles in a scale / zero slot
  uint32_t zmul, zadd;      // z * zmul + zadd = half2 bits of -(z + bias) * 64 * 2^-24 (twice) in lanes r == 0, 0 elsewhere
  int brow_off;             // activation row offset (halves) of this lane's batch column
  int zt_off;               // byte offset of this lane's (hi, lo) group-sum word inside a group's table row
};

#ifdef XBIT_DEVTOOLS
// Phase stamps of tools/trace.py: SM clock values parked in shared memory and written out once, by thread 0 as it leaves
// (a %globaltimer read and a global store per stamp cost 0.4 us per call and sat in front of every fence).  Slots 0..7:
// clock64 at the phase; 13 / 14: %globaltimer at CTA start / at the dump; 15: clock64 at the dump (the host converts).
__shared__ unsigned long long p_trace_sm[16];
template <typename Args>
__device__ __forceinline__ void p_trace_stamp(const Args& a, int slot) {
  if (a.trace) {
    p_trace_sm[slot] = (unsigned long long)clock64();
    if (slot == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p_trace_sm[13] = t;
    }
  }
}
template <typename Args>
__device__ __forceinline__ void p_trace_value(const Args& a, int slot, unsigned long long v) {
  if (a.trace) p_trace_sm[slot] = v;
}
template <typename Args>
__device__ __forceinline__ void p_trace_dump(const Args& a) {
  if (a.trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p_trace_sm[14] = t;
    p_trace_sm[15] = (unsigned long long)clock64();
    for (int i = 0; i < 16; ++i) a.trace[(size_t)blockIdx.x * 16 + i] = p_trace_sm[i];
  }
}
#define P_TRACE(slot) p_trace_stamp(a, slot)
#define P_TRACE_VALUE(slot, v) p_trace_value(a, slot, v)
#else
#define P_TRACE(slot) ((void)0)
#define P_TRACE_VALUE(slot, v) ((void)0)
#endif

// non-volatile MMA wrappers: pure functions of their operands, so ptxas may interleave the two blocks of a step
__device__ __forceinline__ void pmma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void pmma_zero(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}

// NB (1 or 2) 128-k blocks of one tile.  Same exact-product math as w4_consume_block_v2 (gemv_sm100.cu): the
// masked nibble / byte bits are fp16 subnormals, activations are staged as (a0 - a1/16, a1/16) pairs, the zero
// point rides on one extra MMA per column pair and scale group; tot accumulates 2^-24 * y.
template <int UPG, int NB>
__device__ __forceinline__ void w4p_consume(const unsigned char* const (&wp)[NB], const unsigned char* const (&sp)[NB],
                                            const unsigned char* const (&zp)[NB], const __half* const (&ap)[NB],
                                            const unsigned char* const (&ztp)[NB], int zt_group_bytes, const W4PLane& L,
                                            float (&tot)[2][4]) {
  constexpr int GPB = 4 / UPG;
  constexpr uint32_t kOddXor = (UPG == 1) ? 64u : 16u;
  auto unit_row = [](int u) constexpr { return (UPG == 1) ? 4 * u : 8 * (u >> 1) + (u & 1); };
#pragma unroll
  for (int q = 0; q < GPB; ++q) {
    uint2 sraw[NB];
    uint32_t zraw[NB], bz[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      sraw[b] = *reinterpret_cast<const uint2*>(sp[b] + L.s_off + q * 64);
      zraw[b] = *reinterpret_cast<const unsigned short*>(zp[b] + L.z_off + q * 16);
      bz[b] = *reinterpret_cast<const uint32_t*>(ztp[b] + q * zt_group_bytes + L.zt_off);
    }
    float grp[NB][2][4];
#pragma unroll
    for (int uu = 0; uu < UPG; ++uu) {
      const int u = q * UPG + uu;
      uint4 wv[NB], bf[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        wv[b] = *reinterpret_cast<const uint4*>(wp[b] + ((u & 1) ? (L.w_x0 ^ kOddXor) : L.w_x0) + unit_row(u) * 128);
        bf[b] = *reinterpret_cast<const uint4*>(ap[b] + L.brow_off + unit_row(u) * 8);
      }
#pragma unroll
      for (int tt = 0; tt < 2; ++tt)
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          uint32_t ea[4], eb[4];
          unpack_w4_bytes(tt == 0 ? wv[b].x : wv[b].z, ea);
          unpack_w4_bytes(tt == 0 ? wv[b].y : wv[b].w, eb);
          if (uu == 0) pmma_zero(grp[b][tt], ea[0], eb[0], ea[1], eb[1], bf[b].x, bf[b].y);
          else         pmma(grp[b][tt], ea[0], eb[0], ea[1], eb[1], bf[b].x, bf[b].y);
          pmma(grp[b][tt], ea[2], eb[2], ea[3], eb[3], bf[b].z, bf[b].w);
        }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      // zero point: -(z + bias) * sum_k a_k on the tensor core (A = z in k slots 0, 1 of lanes r == 0, B = (hi, lo) of sum/64)
      uint32_t za[4];
      za[0] = (zraw[b] & 0xFu) * L.zmul + L.zadd;
      za[1] = ((zraw[b] >> 4) & 0xFu) * L.zmul + L.zadd;
      za[2] = ((zraw[b] >> 8) & 0xFu) * L.zmul + L.zadd;
      za[3] = (zraw[b] >> 12) * L.zmul + L.zadd;
#pragma unroll
      for (int tt = 0; tt < 2; ++tt) pmma(grp[b][tt], za[2 * tt], za[2 * tt + 1], 0u, 0u, bz[b], 0u);
      // grp = 2^-24 * sum_k a_k (w_k - z); accumulators 0,1 belong to column 2*tt (rows m = 2r, 2r+1), 2,3 to column 2*tt+1
      const float2 s01 = __half22float2(u2h2(sraw[b].x));
      const float2 s23 = __half22float2(u2h2(sraw[b].y));
      const float sf[4] = {s01.x, s01.y, s23.x, s23.y};
#pragma unroll
      for (int tt = 0; tt < 2; ++tt)
#pragma unroll
        for (int i = 0; i < 4; ++i) tot[tt][i] = fmaf(sf[2 * tt + (i >> 1)], grp[b][tt][i], tot[tt][i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Integer block math (groupsize 128, M <= 2): 2 LOP3 per packed word instead of 2 PRMT + 2 LOP3, 8 IMMA per block
// instead of 18 HMMA, dependent MMA chains of 4 instead of 9.
//   * A operand = the packed word itself, masked: (w & 0x0F0F0F0F) are the even nibbles of a word as four u8, and
//     (w & 0xF0F0F0F0) the odd nibbles times 16 -- still u8, so no shift is needed;
//   * B operand = the activations as 24-bit fixed point relative to the largest |a| of their scale group, split into
//     three unsigned byte planes: u = a * 2^(22-E) + 2^22 is read straight out of the mantissa of
//     fmaf(a, 2^(22-E), 1.5 * 2^23) (odd k: a * 2^(18-E), to undo the 16 of the A operand), its three low bytes are the
//     planes, and a fourth plane of ones gives S = sum_k A_k for the 2^22 offset:
//         sum_k A_k q_k = D0 + 256 D1 + 65536 D2 - 2^22 S,   all four sums exact in int32;
//     rounding the activations to 2^(E-23) (odd k: 2^(E-19)) of their group maximum is 2 to 5 orders of magnitude below
//     the fp16 rounding of the result;
//   * mma.sync.m16n8k32.u8.u8.s32: weight columns on M, (row 0: d0 d1 d2 ones, row 1: d0 d1 d2 ones) on N, so lane
//     (t, g) holds, for its four weight columns, the pair (D0, D1) or (D2, S) of activation row t / 2 and folds it with
//     ONE integer multiply-add (D0 + 256 D1, resp. D2 - 64 S, to be scaled by 65536) before the single I2F;
//   * the zero point is applied per lane and column: -(s z) * sum_k a_k, with the group sums staged next to the group's
//     2^(E-22).
struct W4PLaneI {
  uint32_t w_x0;            // byte offset of this lane's 16-byte chunk in a weight block for even units (odd: ^ 16)
  uint32_t s_off, z_off;    // byte offsets of this lane's 4 scales / 4 zero nibbles in a block's scale / zero row
  const unsigned char* bptr;   // digit plane of this lane's B column (lane / 4), or the ones / zero constant
  int bstride;              // bytes per word-row in that plane: 8, or 0 for the constants
  int lane_row;             // 2 * (lane % 4): this lane's word-row inside a unit pair
  int crow4;                // 4 * (activation row of this lane's accumulators) = 4 * ((lane % 4) / 2)
  int cmul;                 // 256 for (D0, D1) lanes, -64 for (D2, S) lanes
  uint32_t ssel;            // PRMT selector: scale of column (lane % 4) of the lane's four into the low half
  int zsh;                  // 4 * (lane % 4)
  float zbias;
};

__device__ __forceinline__ void pimma(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void pimma_zero(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(0));
}

// NB (1 or 2) blocks of 128 k = one scale group each.  wrow[b] = first word-row of the block inside the activation
// row (16 * k-block), gt[b] = the group's table entry {2^(E-22) row 0, row 1, sum_k a_k row 0, row 1}.
// tot[tt][h] accumulates sf * 2^(E-22) * (D0 + 256 D1 | D2 - 64 S) for weight column 4g + 2tt + h; zc[m] the zero-point
// term of column 4g + t for activation row m.
// BF: the scales are bf16 (bf16-native form of the kernel)
template <int NB, bool BF = false>
__device__ __forceinline__ void w4p_consume_i8(const unsigned char* const (&wp)[NB], const unsigned char* const (&sp)[NB],
                                               const unsigned char* const (&zp)[NB], const int (&wrow)[NB],
                                               const float* const (&gt)[NB], const W4PLaneI& L, float (&tot)[2][2], float (&zc)[2]) {
  uint2 sraw[NB];
  uint32_t zraw[NB];
  float gsv[NB];
  float2 aq[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    sraw[b] = *reinterpret_cast<const uint2*>(sp[b] + L.s_off);
    zraw[b] = *reinterpret_cast<const unsigned short*>(zp[b] + L.z_off);
    gsv[b] = *reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(gt[b]) + L.crow4);
    aq[b] = *reinterpret_cast<const float2*>(gt[b] + 2);
  }
  int acc[NB][2][4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int urow = 8 * (u >> 1) + (u & 1);
    uint4 wv[NB];
    uint2 bf[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      wv[b] = *reinterpret_cast<const uint4*>(wp[b] + ((u & 1) ? (L.w_x0 ^ 16u) : L.w_x0) + urow * 128);
      bf[b] = *reinterpret_cast<const uint2*>(L.bptr + (wrow[b] + urow + L.lane_row) * L.bstride);
    }
#pragma unroll
    for (int tt = 0; tt < 2; ++tt)
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const uint32_t w0 = tt == 0 ? wv[b].x : wv[b].z, w1 = tt == 0 ? wv[b].y : wv[b].w;
        const uint32_t a0 = w0 & 0x0F0F0F0Fu, a1 = w1 & 0x0F0F0F0Fu, a2 = w0 & 0xF0F0F0F0u, a3 = w1 & 0xF0F0F0F0u;
        if (u == 0) pimma_zero(acc[b][tt], a0, a1, a2, a3, bf[b].x, bf[b].y);
        else        pimma(acc[b][tt], a0, a1, a2, a3, bf[b].x, bf[b].y);
      }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const float2 s01 = BF ? make_float2(__uint_as_float(sraw[b].x << 16), __uint_as_float(sraw[b].x & 0xffff0000u)) : __half22float2(u2h2(sraw[b].x));
    const float2 s23 = BF ? make_float2(__uint_as_float(sraw[b].y << 16), __uint_as_float(sraw[b].y & 0xffff0000u)) : __half22float2(u2h2(sraw[b].y));
    const float sfg[4] = {s01.x * gsv[b], s01.y * gsv[b], s23.x * gsv[b], s23.y * gsv[b]};
#pragma unroll
    for (int tt = 0; tt < 2; ++tt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int v = acc[b][tt][2 * h] + L.cmul * acc[b][tt][2 * h + 1];
        tot[tt][h] = fmaf(sfg[2 * tt + h], (float)v, tot[tt][h]);
      }
    // zero point of column 4g + t: -(s * (z + bias)) * sum_k a_k per activation row
    const uint32_t s16 = prmt(sraw[b].x, sraw[b].y, L.ssel) & 0xffffu;
    const float sz = (BF ? __uint_as_float(s16 << 16) : __half2float(__ushort_as_half((unsigned short)s16))) *
                     ((float)((zraw[b] >> L.zsh) & 0xFu) + L.zbias);
    zc[0] = fmaf(sz, aq[b].x, zc[0]);
    zc[1] = fmaf(sz, aq[b].y, zc[1]);
  }
}

// The same for 8-bit weights (A16W8, groupsize 128, M <= 2): a packed word IS an A register -- four consecutive k of one
// column as four u8 -- so there is no unpack instruction at all; the activations' digit planes are in natural k order and
// all carry the scale 2^(22-E).  A block is 32 word-rows: per step u a lane takes rows 2(t + 4u), 2(t + 4u) + 1 (eight
// consecutive k) of its four columns with two 16-byte loads, and the eight digit bytes of those k with one 8-byte load.
// Largest sum: 255 * 128 * 65535 < 2^31.
template <int NB, bool BF = false>
__device__ __forceinline__ void w8p_consume_i8(const unsigned char* const (&wp)[NB], const unsigned char* const (&sp)[NB],
                                               const unsigned char* const (&zp)[NB], const int (&wrow)[NB],
                                               const float* const (&gt)[NB], const W4PLaneI& L, float (&tot)[2][2], float (&zc)[2]) {
  uint2 sraw[NB];
  uint32_t zraw[NB];
  float gsv[NB];
  float2 aq[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    sraw[b] = *reinterpret_cast<const uint2*>(sp[b] + L.s_off);
    zraw[b] = *reinterpret_cast<const uint32_t*>(zp[b] + 2 * L.z_off);        // four 8-bit zero points of this lane's columns
    gsv[b] = *reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(gt[b]) + L.crow4);
    aq[b] = *reinterpret_cast<const float2*>(gt[b] + 2);
  }
  int acc[NB][2][4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    uint4 w1[NB], w2[NB];
    uint2 bf[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      w1[b] = *reinterpret_cast<const uint4*>(wp[b] + L.w_x0 + u * 1024);
      w2[b] = *reinterpret_cast<const uint4*>(wp[b] + (L.w_x0 ^ 16u) + 128 + u * 1024);
      bf[b] = *reinterpret_cast<const uint2*>(L.bptr + (wrow[b] + (L.lane_row >> 1) + 4 * u) * L.bstride);
    }
#pragma unroll
    for (int tt = 0; tt < 2; ++tt)
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const uint32_t a0 = tt == 0 ? w1[b].x : w1[b].z, a1 = tt == 0 ? w1[b].y : w1[b].w;
        const uint32_t a2 = tt == 0 ? w2[b].x : w2[b].z, a3 = tt == 0 ? w2[b].y : w2[b].w;
        if (u == 0) pimma_zero(acc[b][tt], a0, a1, a2, a3, bf[b].x, bf[b].y);
        else        pimma(acc[b][tt], a0, a1, a2, a3, bf[b].x, bf[b].y);
      }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const float2 s01 = BF ? make_float2(__uint_as_float(sraw[b].x << 16), __uint_as_float(sraw[b].x & 0xffff0000u)) : __half22float2(u2h2(sraw[b].x));
    const float2 s23 = BF ? make_float2(__uint_as_float(sraw[b].y << 16), __uint_as_float(sraw[b].y & 0xffff0000u)) : __half22float2(u2h2(sraw[b].y));
    const float sfg[4] = {s01.x * gsv[b], s01.y * gsv[b], s23.x * gsv[b], s23.y * gsv[b]};
#pragma unroll
    for (int tt = 0; tt < 2; ++tt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int v = acc[b][tt][2 * h] + L.cmul * acc[b][tt][2 * h + 1];
        tot[tt][h] = fmaf(sfg[2 * tt + h], (float)v, tot[tt][h]);
      }
    const uint32_t s16 = prmt(sraw[b].x, sraw[b].y, L.ssel) & 0xffffu;
    const float sz = (BF ? __uint_as_float(s16 << 16) : __half2float(__ushort_as_half((unsigned short)s16))) *
                     ((float)((zraw[b] >> (2 * L.zsh)) & 0xFFu) + L.zbias);
    zc[0] = fmaf(sz, aq[b].x, zc[0]);
    zc[1] = fmaf(sz, aq[b].y, zc[1]);
  }
}

// ... and for 2-bit weights (A16W2): a packed word is 16 consecutive k; the masks 0x03 << 2c of its bytes are the k with
// k mod 4 = c as four u8 times 4^c, so one word yields four A registers with no shift (the activations' digit planes carry
// 4^-c of the scale for those k, see the staging).  A block is 8 word-rows: a lane takes rows t and t + 4 of its four columns.
template <int NB, bool BF = false>
__device__ __forceinline__ void w2p_consume_i8(const unsigned char* const (&wp)[NB], const unsigned char* const (&sp)[NB],
                                               const unsigned char* const (&zp)[NB], const int (&wrow)[NB],
                                               const float* const (&gt)[NB], const W4PLaneI& L, float (&tot)[2][2], float (&zc)[2]) {
  uint2 sraw[NB];
  uint32_t zraw[NB];
  float gsv[NB];
  float2 aq[NB];
  const int t = L.lane_row >> 1;
  const uint32_t x0 = (uint32_t)(t * 128) + (((L.s_off >> 3) ^ (uint32_t)t) << 4);      // row t, chunk (lane / 4) ^ t
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    sraw[b] = *reinterpret_cast<const uint2*>(sp[b] + L.s_off);
    zraw[b] = *reinterpret_cast<const unsigned char*>(zp[b] + (L.z_off >> 1));           // four 2-bit zero points of this lane's columns
    gsv[b] = *reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(gt[b]) + L.crow4);
    aq[b] = *reinterpret_cast<const float2*>(gt[b] + 2);
  }
  int acc[NB][2][4];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    uint4 wv[NB];
    uint2 bf[NB][2];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      wv[b] = *reinterpret_cast<const uint4*>(wp[b] + (u ? (x0 ^ 64u) + 512u : x0));
#pragma unroll
      for (int cc = 0; cc < 2; ++cc)
        bf[b][cc] = *reinterpret_cast<const uint2*>(L.bptr + (wrow[b] + 2 * (t + 4 * u) + cc) * L.bstride);
    }
#pragma unroll
    for (int cc = 0; cc < 2; ++cc)
#pragma unroll
      for (int tt = 0; tt < 2; ++tt)
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const uint32_t ca = tt == 0 ? wv[b].x : wv[b].z, cb = tt == 0 ? wv[b].y : wv[b].w;
          const uint32_t m0 = 0x03030303u << (4 * cc), m1 = 0x0C0C0C0Cu << (4 * cc);
          if (u == 0 && cc == 0) pimma_zero(acc[b][tt], ca & m0, cb & m0, ca & m1, cb & m1, bf[b][cc].x, bf[b][cc].y);
          else                   pimma(acc[b][tt], ca & m0, cb & m0, ca & m1, cb & m1, bf[b][cc].x, bf[b][cc].y);
        }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const float2 s01 = BF ? make_float2(__uint_as_float(sraw[b].x << 16), __uint_as_float(sraw[b].x & 0xffff0000u)) : __half22float2(u2h2(sraw[b].x));
    const float2 s23 = BF ? make_float2(__uint_as_float(sraw[b].y << 16), __uint_as_float(sraw[b].y & 0xffff0000u)) : __half22float2(u2h2(sraw[b].y));
    const float sfg[4] = {s01.x * gsv[b], s01.y * gsv[b], s23.x * gsv[b], s23.y * gsv[b]};
#pragma unroll
    for (int tt = 0; tt < 2; ++tt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int v = acc[b][tt][2 * h] + L.cmul * acc[b][tt][2 * h + 1];
        tot[tt][h] = fmaf(sfg[2 * tt + h], (float)v, tot[tt][h]);
      }
    const uint32_t s16 = prmt(sraw[b].x, sraw[b].y, L.ssel) & 0xffffu;
    const float sz = (BF ? __uint_as_float(s16 << 16) : __half2float(__ushort_as_half((unsigned short)s16))) *
                     ((float)((zraw[b] >> (L.zsh >> 1)) & 0x3u) + L.zbias);
    zc[0] = fmaf(sz, aq[b].x, zc[0]);
    zc[1] = fmaf(sz, aq[b].y, zc[1]);
  }
}

template <int BITS, int NB, bool BF>
__device__ __forceinline__ void wxp_consume_i8(const unsigned char* const (&wp)[NB], const unsigned char* const (&sp)[NB],
                                               const unsigned char* const (&zp)[NB], const int (&wrow)[NB],
                                               const float* const (&gt)[NB], const W4PLaneI& L, float (&tot)[2][2], float (&zc)[2]) {
  if constexpr (BITS == 8) w8p_consume_i8<NB, BF>(wp, sp, zp, wrow, gt, L, tot, zc);
  else if constexpr (BITS == 2) w2p_consume_i8<NB, BF>(wp, sp, zp, wrow, gt, L, tot, zc);
  else w4p_consume_i8<NB, BF>(wp, sp, zp, wrow, gt, L, tot, zc);
}

__device__ __forceinline__ void p_ll_store(unsigned long long* slot, uint32_t data, uint32_t epoch) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(slot), "r"(data), "r"(epoch) : "memory");
}
// 8 consecutive halves = 4 slots; spins until all four carry `epoch` (guard: reports in ll_state[3])
__device__ __forceinline__ uint4 p_ll_load8(const unsigned long long* slots, uint32_t epoch, unsigned int* timeout_flag) {
  uint4 q0, q1;
  const long long c0 = clock64();
  for (unsigned int n = 1;; ++n) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w) : "l"(slots) : "memory");
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w) : "l"(slots + 2) : "memory");
    if (q0.y == epoch && q0.w == epoch && q1.y == epoch && q1.w == epoch) break;
    if ((n & 1023u) == 0 && clock64() - c0 > kPSpinGuardClocks) {
      *timeout_flag = 1u;
      break;
    }
  }
  return make_uint4(q0.x, q0.z, q1.x, q1.z);
}

// slice (eighth of the CTA's range) of consumer warp w: warps w and w + 4 share an SM sub-partition and adjacent
// slices differ by at most one block, so every sub-partition gets two ADJACENT slices
template <int NW>
__device__ __forceinline__ int p_slice_of(int w) { return (w & 3) * (NW / 4) + (w >> 2); }

// NW consumer warps + one producer warp, a ring per consumer warp.  DUAL: the two blocks of a step are taken together
// (two accumulator sets; the form every instantiation uses -- one block at a time, 12 warps at a time, 16 warps sharing
// rings in pairs and 16 warps on half an SM were all measured slower).
// BPS: blocks per ring slot (2 = one TMA request per step of two blocks).
// MINB: CTAs per SM the register budget is cut for: 2 = two launches co-resident (the whole share of a CTA fits its
// rings and is prefetched while the previous launch computes), 1 = one CTA per SM with deeper rings or 16 consumer warps
// (larger matrices, which keep streaming while they compute: what counts there is the compute rate and bytes in flight).
// GEN: what the instantiation covers.  0 = one weight matrix, plain loads and stores (the form every single call takes:
// the loops over matrices fold away and the activation loads of the staging phase are issued back to back; measured
// 3.37 against 3.96 us on 4096 x 4096 and 6.97 against 8.25 us on 11008 x 4096 for the general form, profiles/r02_ab.log);
// 1 = several matrices (xbit_gemv_f16_multi); 2 = also the flag-in-data (LL) input / output forms.
// BF: bf16-native form (activations, scales and output bf16; integer block math only): the activations' 24-bit fixed
// point comes from the bf16 exponent, the scales widen by a shift, the result is rounded once, to bf16.
// BITS: 4, or 8 (A16W8: integer block math only; the packed words are the MMA operands as they are, w8p_consume_i8).
template <int UPG, int NW, int MODE, bool I8, int BPS, int MINB, int GEN, bool BF = false, int BITS = 4>
__global__ void __launch_bounds__((NW + 1) * 32, MINB)
gemv_w4p_kernel(const __grid_constant__ W4PMapsN<(GEN ? kPMaxProblems : 1)> maps, const __grid_constant__ W4PArgsN<(GEN ? kPMaxProblems : 1)> a) {
  using Cfg = W4PCfg<UPG, BPS, BITS>;
  static_assert(BITS == 4 || ((BITS == 8 || BITS == 2) && I8 && GEN == 0), "2- / 8-bit weights: integer block math, one matrix per launch");
  const int count = GEN >= 1 ? a.count : 1;
  const bool ll_out = GEN == 2 && a.ll_out, a_is_ll = GEN == 2 && a.a_is_ll;
  constexpr int GPB = Cfg::GPB;
  constexpr bool DUAL = MODE == 1;
  static_assert(MODE == 0 || MODE == 1, "MODE: 0 = one block at a time, 1 = two blocks together");
  static_assert(!I8 || UPG == 4, "the integer block math covers groupsize 128");
  static_assert(!BF || (I8 && GEN == 0), "the bf16-native form exists for the integer block math, one matrix per launch");
  constexpr int kPWarps = NW;                       // rings = slices of the CTA's range
  constexpr int kPConsumerThreads = NW * 32;
  constexpr int LPR = kPWarps <= 8 ? 4 : 2;         // producer lanes per ring
  auto slice_of_ring = [](int rg) { return p_slice_of<kPWarps>(rg); };
  extern __shared__ __align__(1024) unsigned char smem_raw[];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, c = blockIdx.x;
  const int R = a.ring;
  const int nb = a.nb;
  const int pitch = a.K + 8;                        // halves per staged activation row (+16 B: batch rows land in different banks)
  const int zt_group_bytes = a.M * 16;

  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B boxes need 1024-byte alignment
  unsigned char* wring = base;                                                        // [8][R] weight slots
  unsigned char* sring = wring + kPWarps * R * Cfg::kWSlot;                           // [8][R] scale slots
  unsigned char* zring = sring + kPWarps * R * Cfg::kSSlot;                           // [8][R] zero slots
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(zring + kPWarps * R * Cfg::kZSlot);     // [8][kPMaxRing]
  uint64_t* empty_bar = full_bar + kPWarps * kPMaxRing;
  int* cnt_sm = reinterpret_cast<int*>(empty_bar + kPWarps * kPMaxRing);              // [problems][16] arrival counters of shared tiles
  int* bnd_sm = cnt_sm + kPMaxProblems * 16;                                        // [problems][32] first block of every slice
  float* part_sm = reinterpret_cast<float*>(bnd_sm + kPMaxProblems * 32);           // [problems][slices]([2 warps])[2]([2])[M][32] partial tiles
  const int part_stride = kPWarps * 2 * a.M * 32;                        // floats per problem
  uint32_t* zt_sm = reinterpret_cast<uint32_t*>(part_sm + count * part_stride);    // [groups][M][4]: (hi, lo) of sum_k a_k / 64, 3 zero words
  __half* act_sm = reinterpret_cast<__half*>(zt_sm + (size_t)nb * GPB * a.M * 4);     // [M][pitch]
  // integer block math instead: group table {2^(E-22) row 0, row 1, sum_k a_k row 0, row 1}, the ones / zero constants,
  // and three digit planes per activation row ([K/8 word-rows][even k x4, odd k x4] bytes each; plane p starts
  // p * (K + 128) + 8 * {0, 1, 8, 9}[p % 4] bytes in, which keeps the 8 lanes (t, g < 3) of a half-warp on distinct banks)
  float* gt_sm = reinterpret_cast<float*>(zt_sm);                                     // [groups][4]
  unsigned char* const_sm = reinterpret_cast<unsigned char*>(gt_sm + (size_t)nb * 4);    // 8 x 0x01, 8 x 0x00
  unsigned char* dig_sm = const_sm + 16;                                              // [3 * M planes]
  auto plane_off = [&](int p) { return (size_t)p * (a.K + 128) + 8 * ((p & 1) + 8 * ((p >> 1) & 1)); };

  // this CTA's range of a matrix' tile-major block list: units [U*c/G, U*(c+1)/G) with U = uq*G + ur (no 64-bit division
  // here: the time from CTA start to the first TMA request is on the critical path whenever the CTA could not start early)
  auto range_lo = [&](const W4PProblem& P, int cc) { return (P.uq * cc + (int)__umulhi((unsigned)(P.ur * cc), a.g_magic)) * P.unit; };
  auto tile_of = [&](int jj) { return (int)(((unsigned long long)(unsigned)jj * a.nb_magic) >> a.nb_shift); };
  // (the first matrix' range: worked out by every thread while the barriers are being initialised)
  const int lo0 = range_lo(a.prob[0], c), hi0 = range_lo(a.prob[0], c + 1);

  if (tid < kPWarps * kPMaxRing) {
    if (tid == 0) {
#ifdef XBIT_DEVTOOLS
      for (int i = 0; i < 16; ++i) p_trace_sm[i] = 0;
#endif
      P_TRACE(0);
#ifdef XBIT_DEVTOOLS
      unsigned int smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      P_TRACE_VALUE(12, (unsigned long long)smid + 1);
#endif
    }
    if ((tid & (kPMaxRing - 1)) < R) {
      mbar_init(&full_bar[tid], 1 + LPR);           // the TMA issuer's expect_tx arrival + LPR cp.async arrivals
      mbar_init(&empty_bar[tid], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  if constexpr (GEN == 0) {
    if (tid < 32) {
      if (tid < 16) cnt_sm[tid] = 0;
      if (tid <= kPWarps) bnd_sm[tid] = lo0 + (int)(((long long)(hi0 - lo0) * tid) / kPWarps);
    }
  } else if (tid < kPMaxProblems * 32) {
    const int pi = tid >> 5, i = tid & 31;
    if (i < 16) cnt_sm[pi * 16 + i] = 0;
    if (pi < count && i <= kPWarps) {
      const int plo = pi == 0 ? lo0 : range_lo(a.prob[pi], c), plen = (pi == 0 ? hi0 : range_lo(a.prob[pi], c + 1)) - plo;
      bnd_sm[pi * 32 + i] = plo + (int)(((long long)plen * i) / kPWarps);
    }
  }
  // the producer lanes' slices of the first matrix, worked out before anything else, while the barriers are initialised:
  // the time from CTA start to the first TMA request is on the critical path whenever the CTA could not start early
  // (7.00 against 7.19 us on 11008 x 4096 with these divisions behind the barrier, profiles/r02_ab.log)
  int j0 = 0, jend0 = 0, tile0 = 0, kb0 = 0;
  if (warp == NW && lane < kPWarps * LPR) {
    const int rho = slice_of_ring(lane / LPR);
    j0 = lo0 + (int)((long long)(hi0 - lo0) * rho / kPWarps);
    jend0 = lo0 + (int)((long long)(hi0 - lo0) * (rho + 1) / kPWarps);
    tile0 = tile_of(j0);
    kb0 = j0 - tile0 * nb;
  }
  __syncthreads();
  // the next kernel in the stream may become resident now: its producer streams ITS weights while this one runs.
  // Exception: the FIRST call of an LL chain releases its dependents only after its own griddepcontrol.wait has returned
  // (the calls behind it do not wait for the grid before them and read the chain base that the previous chain's unpack
  // advances; see gemv_w4_kernel)
  const bool defer_dependents = ll_out && !a_is_ll;
  if (!defer_dependents) griddep_launch_dependents();

  if (warp == NW) {
    // =========================== producer: LPR lanes drive each consumer warp's ring ===========================
    // One step = up to two consecutive blocks of one tile: ONE TMA request for the weights (the TMA unit serves about
    // one request per 64 clk whatever its size: three 2 KiB / 64 B / 16 B boxes per block ran at 8 B/clk per SM), and the
    // scale / zero rows (64 + 16 bytes per block and group) as 16-byte cp.async that arrive on the same mbarrier.
    if (!a.static_weights) griddep_wait();
    else if (a.prefetch_delay > 0) {
      const long long t0 = clock64();
      while (clock64() - t0 < a.prefetch_delay) __nanosleep(100);
    }
    if (lane < kPWarps * LPR) {
      const int w = lane / LPR, sub = lane % LPR, rho = slice_of_ring(w);
      uint64_t policy = 0;
      if (sub == 0) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      if (lane < 2 * count) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.m[lane]) : "memory");
      if (lane == 0) P_TRACE(1);
      int s = 0, ph = 0, n = 0;
      for (int pi = 0; pi < count; ++pi) {
        const W4PProblem& P = a.prob[pi];
        int j = j0, jend = jend0, tile = tile0, kb = kb0;
        if (pi > 0) {
          const int lo = range_lo(P, c), len = range_lo(P, c + 1) - lo;
          j = lo + (int)((long long)len * rho / kPWarps);
          jend = lo + (int)((long long)len * (rho + 1) / kPWarps);
          tile = tile_of(j);
          kb = j - tile * nb;
        }
        const unsigned char* sbase = reinterpret_cast<const unsigned char*>(P.scales) + sub * 16;
        const unsigned char* zbase = reinterpret_cast<const unsigned char*>(P.qzeros);
        for (; j < jend; ++n) {
          const int nblk = min(BPS, min(jend - j, nb - kb));
          uint64_t* fb = &full_bar[w * kPMaxRing + s];
          if (n >= R) mbar_wait(&empty_bar[w * kPMaxRing + s], ph ^ 1);
          if (sub == 0) {
            mbar_arrive_expect_tx(fb, (uint32_t)(nblk * Cfg::kBlockBytes));
            tma_load_2d(wring + (w * R + s) * Cfg::kWSlot, &maps.m[2 * pi + (nblk == 2 ? 0 : 1)], tile * 32, kb * Cfg::kRowsPerBlock, fb, policy);
          }
          unsigned char* sdst = sring + (w * R + s) * Cfg::kSSlot + sub * 16;
          unsigned char* zdst = zring + (w * R + s) * Cfg::kZSlot;
          const int rows = nblk * GPB, row0 = kb * GPB;
#pragma unroll
          for (int rr = 0; rr < BPS * GPB; ++rr)
            if (rr < rows) {
#pragma unroll
              for (int ch = 0; ch < 4; ch += LPR)
                cp_async_16(sdst + rr * 64 + ch * 16, sbase + ((size_t)(row0 + rr) * P.N + tile * 32) * 2 + ch * 16);
            }
#pragma unroll
          for (int rr = 0; rr < BPS * GPB; ++rr)
            if (rr < rows && (rr % LPR) == sub) {
              // (a tile's zero points of one group: 32 * BITS / 8 bytes)
              const unsigned char* zsrc = zbase + ((size_t)(row0 + rr) * P.zwords + tile * BITS) * 4;
              if (BITS == 2) cp_async_8(zdst + rr * Cfg::kZRow, zsrc);
              else cp_async_16(zdst + rr * Cfg::kZRow, zsrc);
              if (BITS == 8) cp_async_16(zdst + rr * Cfg::kZRow + 16, zsrc + 16);
            }
          cp_async_mbar_arrive_noinc(fb);
          j += nblk;
          kb += nblk;
          if (kb == nb) { kb = 0; ++tile; }
          if (++s == R) { s = 0; ph ^= 1; }
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
    }
    return;
  }

  // =========================== consumers ===========================
  const int r = lane & 3, c8 = lane >> 2;
  // everything that does not depend on the activations happens BEFORE the wait: these warps idle until then anyway
  W4PLane L;
  {
    const int lane_row = ((UPG == 1) ? 1 : 2) * r;
    L.w_x0 = (uint32_t)(lane_row * 128 + ((c8 ^ lane_row) * 16));
    L.s_off = (uint32_t)(c8 * 8);
    L.z_off = (uint32_t)(c8 * 2);
    L.zmul = r == 0 ? 0x00400040u : 0u;
    L.zadd = r == 0 ? (uint32_t)a.zero_bias * 0x00400040u + 0x80008000u : 0u;
    const int m = min(c8, a.M - 1);                 // batch rows >= M read a clamped (valid) row: their accumulators are never stored
    L.brow_off = m * pitch + lane_row * 8;
    L.zt_off = (m * 4 + r) * 4;
  }
  W4PLaneI LI;
  if constexpr (I8) {
    LI.w_x0 = L.w_x0;
    LI.s_off = L.s_off;
    LI.z_off = L.z_off;
    // B column c8: planes (row 0: d0 d1 d2, ones; row 1: d0 d1 d2, ones); rows >= M read the zero constant
    const int brow = c8 >> 2, bdig = c8 & 3;
    if (bdig == 3) { LI.bptr = const_sm; LI.bstride = 0; }
    else if (brow >= a.M) { LI.bptr = const_sm + 8; LI.bstride = 0; }
    else { LI.bptr = dig_sm + plane_off(brow * 3 + bdig); LI.bstride = 8; }
    LI.lane_row = 2 * r;
    LI.crow4 = 4 * (r >> 1);
    LI.cmul = (r & 1) ? -64 : 256;
    LI.ssel = (r & 1) ? (0x4400u | ((r & 2) ? 0x76u : 0x32u)) : (0x4400u | ((r & 2) ? 0x54u : 0x10u));
    LI.zsh = 4 * r;
    LI.zbias = (float)a.zero_bias;
  }

  const int rg = warp;                              // ring
  const int rho = slice_of_ring(rg);
  int s = 0, ph = 0;
  // this warp's slice of the CTA's range of a matrix: the first matrix' is worked out here, BEFORE the wait (integer
  // divisions: 0.2 us when they sat between the activation staging and the first block)
  int lo, hi, j, jend, tile, kb;
  auto slice_range = [&](const W4PProblem& P, bool first) {
    lo = first ? lo0 : range_lo(P, c);
    hi = first ? hi0 : range_lo(P, c + 1);
    const int len = hi - lo;
    j = lo + (int)((long long)len * rho / kPWarps);
    jend = lo + (int)((long long)len * (rho + 1) / kPWarps);
    tile = tile_of(j);
    kb = j - tile * nb;
  };
  slice_range(a.prob[0], true);
  uint64_t* const my_full = full_bar + rg * kPMaxRing;
  uint64_t* const my_empty = empty_bar + rg * kPMaxRing;
  const unsigned char* const my_w = wring + rg * R * Cfg::kWSlot;
  const unsigned char* const my_s = sring + rg * R * Cfg::kSSlot;
  const unsigned char* const my_z = zring + rg * R * Cfg::kZSlot;
  const unsigned char* const zt_bytes = reinterpret_cast<const unsigned char*>(zt_sm);
  // (Measured: forcing all of the above to be computed HERE, ahead of the wait -- the compiler sinks most of it behind the
  // wait and the staging barrier, 60 + 70 instructions -- gains 0.02 us where the CTA starts early (half SM) and LOSES
  // 0.2 us where it starts late (full SM): there the consumer warps' set-up competes with the producer warp for issue
  // slots on its way to the first TMA request.  Left to the compiler; profiles/r02_ab.log.)
  // The activations are the only data produced by the previous kernel.  ONE warp waits for it; the others block on a
  // hardware barrier behind that warp: warps parked in griddepcontrol.wait were measured to slow the co-resident CTA of
  // the previous launch down (XBIT_W4P_ALLWAIT=1 restores the plain form for the comparison).
  // (a call fed from an LL buffer carries its dependency in the data: every slot is validated by its own call number)
  if (!a_is_ll) {
    if (a.all_wait || warp == 0) griddep_wait();
    if (!a.all_wait) asm volatile("bar.sync 1, %0;" ::"n"(kPConsumerThreads) : "memory");
  }
  if (defer_dependents) griddep_launch_dependents();
  const uint32_t ll_in_epoch = a_is_ll ? a.ll_state[2] + (uint32_t)a.ll_chain_index : 0u;     // the previous call's number
  if (tid == 0) P_TRACE(2);
  if constexpr (I8) {
    // stage the activations once per CTA as three unsigned byte planes of 24-bit fixed point relative to the largest |a|
    // of their scale group (see w4p_consume_i8), plus per group 2^(E-22) and sum_k a_k.  A group = 16 consecutive
    // vectors = half a warp.
    if (tid < 4) reinterpret_cast<uint32_t*>(const_sm)[tid] = tid < 2 ? 0x01010101u : 0u;
    const int vecs = a.K >> 3;
    // all loads of a thread are issued before the first conversion (K = 11008: 5.4 vectors per thread -- in batches of two
    // the three L2 round trips, under the weight stream's load, made the staging 2.2 us of that call's 7)
    constexpr int kBatch = NW == 16 ? 3 : 6;
    for (int m = 0; m < a.M; ++m) {
      const uint4* arow = reinterpret_cast<const uint4*>(a.a + (size_t)m * a.K);
      unsigned char* const pl0 = dig_sm + plane_off(m * 3 + 0);
      unsigned char* const pl1 = dig_sm + plane_off(m * 3 + 1);
      unsigned char* const pl2 = dig_sm + plane_off(m * 3 + 2);
      for (int v0 = warp * 32; v0 < vecs; v0 += kBatch * kPConsumerThreads) {
        uint4 val[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const int v = v0 + b * kPConsumerThreads + lane;
          val[b] = make_uint4(0, 0, 0, 0);
          if (v < vecs) val[b] = a_is_ll ? p_ll_load8(reinterpret_cast<const unsigned long long*>(a.a) + (((size_t)m * a.K) >> 1) + 4 * (size_t)v, ll_in_epoch, a.ll_state + 3)
                                           : __ldcg(arow + v);
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const int vw = v0 + b * kPConsumerThreads;                // warp-uniform
          if (vw >= vecs) break;
          const int v = vw + lane;
          const bool ok = v < vecs;
          auto widen = [](uint32_t x) {
            return BF ? make_float2(__uint_as_float(x << 16), __uint_as_float(x & 0xffff0000u)) : __half22float2(u2h2(x));
          };
          const float2 f0 = widen(val[b].x), f1 = widen(val[b].y), f2 = widen(val[b].z), f3 = widen(val[b].w);
          // largest |a| of the group as fp16 (bf16) bits (the bit patterns of non-negative halves order like integers): one
          // REDUX over the half-warp instead of a shuffle tree -- this phase sits on the critical path of every call
          uint32_t am2;
          if constexpr (BF) {
            auto b2 = [](uint32_t x) { return *reinterpret_cast<const __nv_bfloat162*>(&x); };
            const __nv_bfloat162 ax = __hmax2(__habs2(b2(val[b].x)), __habs2(b2(val[b].y)));
            const __nv_bfloat162 az = __hmax2(__habs2(b2(val[b].z)), __habs2(b2(val[b].w)));
            const __nv_bfloat162 am = __hmax2(ax, az);
            am2 = *reinterpret_cast<const uint32_t*>(&am);
          } else {
            const __half2 ax = __hmax2(__habs2(u2h2(val[b].x)), __habs2(u2h2(val[b].y)));
            const __half2 az = __hmax2(__habs2(u2h2(val[b].z)), __habs2(u2h2(val[b].w)));
            am2 = h22u(__hmax2(ax, az));
          }
          // (full-warp REDUX twice, one per half: a half-warp mask makes the compiler serialise the halves)
          const bool upper = (lane & 16) != 0;
          const uint32_t amax1 = max(am2 & 0xFFFFu, am2 >> 16);
          const uint32_t mlo = __reduce_max_sync(0xffffffffu, upper ? 0u : amax1);
          const uint32_t mhi = __reduce_max_sync(0xffffffffu, upper ? amax1 : 0u);
          const uint32_t amax = upper ? mhi : mlo;
          // |a| < 2^E with E = exponent field - 14; q = a * 2^(22-E) (odd k: 2^(18-E)); inf / nan activations poison the group
          // (bf16: |a| < 2^E with E = exponent field - 126; fields below 22 -- |a| < 2^-104 -- share the scale of 22)
          const uint32_t eb = BF ? max(amax >> 7, 22u) : amax >> 10;
          const float se = __uint_as_float(((BF ? 275u : 163u) - eb) << 23), so = __uint_as_float(((BF ? 271u : 159u) - eb) << 23);
          const float gs = eb >= (BF ? 255u : 31u) ? __uint_as_float(0x7fc00000u) : __uint_as_float(BF ? (eb - 21u) << 23 : (91u + eb) << 23);
          const float kMagic = 12582912.f;         // 1.5 * 2^23: the bits of (q + kMagic) are 0x4B400000 + q
          // 4-bit weights: (u0, u2, u4, u6) = even k, (u1, u3, u5, u7) = odd k at 1/16 of the scale (the odd nibbles reach the
          // MMA times 16).  8-bit weights: natural order, one scale: (u0, u2, u4, u6) = k 0..3, (u1, u3, u5, u7) = k 4..7.
          // 2-bit weights: k mod 4 = c reaches the MMA times 4^c (the masks 0x03 << 2c of a byte): u_k carries 4^-c of the scale
          const float so_ = BITS == 8 ? se : so;
          const float x1 = BITS == 8 ? f2.x : f0.y, x2 = BITS == 8 ? f0.y : f1.x, x3 = BITS == 8 ? f2.y : f1.y;
          const float x4 = BITS == 8 ? f1.x : f2.x, x5 = BITS == 8 ? f3.x : f2.y, x6 = BITS == 8 ? f1.y : f3.x;
          uint32_t u0, u1, u2, u3, u4, u5, u6, u7;
          int q1;
          if constexpr (BITS == 2) {
            const float s1 = se * 0.25f, s2 = so, s3 = so * 0.25f;       // 2^(22-E) / 4, / 16, / 64
            u0 = __float_as_uint(fmaf(f0.x, se, kMagic)); u1 = __float_as_uint(fmaf(f0.y, s1, kMagic));
            u2 = __float_as_uint(fmaf(f1.x, s2, kMagic)); u3 = __float_as_uint(fmaf(f1.y, s3, kMagic));
            u4 = __float_as_uint(fmaf(f2.x, se, kMagic)); u5 = __float_as_uint(fmaf(f2.y, s1, kMagic));
            u6 = __float_as_uint(fmaf(f3.x, s2, kMagic)); u7 = __float_as_uint(fmaf(f3.y, s3, kMagic));
            q1 = (int)((u0 + u4) + 4u * (u1 + u5) + 16u * (u2 + u6) + 64u * (u3 + u7) - 170u * 0x4B400000u);
          } else {
            u0 = __float_as_uint(fmaf(f0.x, se, kMagic)); u1 = __float_as_uint(fmaf(x1, so_, kMagic));
            u2 = __float_as_uint(fmaf(x2, se, kMagic)); u3 = __float_as_uint(fmaf(x3, so_, kMagic));
            u4 = __float_as_uint(fmaf(x4, se, kMagic)); u5 = __float_as_uint(fmaf(x5, so_, kMagic));
            u6 = __float_as_uint(fmaf(x6, se, kMagic)); u7 = __float_as_uint(fmaf(f3.y, so_, kMagic));
            // sum_k q_k in units of 2^(E-22), exactly, in integers (|.| <= 2^29): even k + 16 * odd k - 68 * 0x4B400000
            q1 = BITS == 8 ? (int)(((u0 + u2) + (u4 + u6)) + ((u1 + u3) + (u5 + u7)) - 8u * 0x4B400000u)
                           : (int)(((u0 + u2) + (u4 + u6)) + 16u * ((u1 + u3) + (u5 + u7)) - 68u * 0x4B400000u);
          }
          const int slo = __reduce_add_sync(0xffffffffu, upper ? 0 : q1);
          const int shi = __reduce_add_sync(0xffffffffu, upper ? q1 : 0);
          const int qsum = upper ? shi : slo;
          const float sum = (float)qsum * gs;
          uint2 st0, st1, st2;
          if constexpr (BITS == 2) {
            // a word-row of 2-bit weights is 16 k = two vectors = a lane pair.  Per plane its record is 16 bytes,
            // [k mod 4 = 0: k 0 4 8 12][= 1][= 2][= 3]: the even lane (k 0..7) stores the first eight, the odd lane (k 8..15)
            // the last eight, after one exchange of the two classes the other lane stores
            const bool odd = (lane & 1) != 0;
            auto mix = [&](uint32_t own01, uint32_t own23) {
              const uint32_t recv = __shfl_xor_sync(0xffffffffu, odd ? own01 : own23, 1);
              const uint32_t x = odd ? recv : own01, y = odd ? own23 : recv;
              return make_uint2(prmt(x, y, 0x5410), prmt(x, y, 0x7632));
            };
            const uint32_t t04 = prmt(u0, u4, 0x5140), t15 = prmt(u1, u5, 0x5140), t26 = prmt(u2, u6, 0x5140), t37 = prmt(u3, u7, 0x5140);
            const uint32_t h04 = prmt(u0, u4, 0x0062), h15 = prmt(u1, u5, 0x0062), h26 = prmt(u2, u6, 0x0062), h37 = prmt(u3, u7, 0x0062);
            st0 = mix(prmt(t04, t15, 0x5410), prmt(t26, t37, 0x5410));
            st1 = mix(prmt(t04, t15, 0x7632), prmt(t26, t37, 0x7632));
            st2 = mix(prmt(h04, h15, 0x5410), prmt(h26, h37, 0x5410));
          } else {
            // byte j of (u0, u2, u4, u6) -> even word of plane j, of (u1, u3, u5, u7) -> odd word
            const uint32_t e02 = prmt(u0, u2, 0x5140), e46 = prmt(u4, u6, 0x5140);
            const uint32_t o13 = prmt(u1, u3, 0x5140), o57 = prmt(u5, u7, 0x5140);
            const uint32_t e02h = prmt(u0, u2, 0x0062), e46h = prmt(u4, u6, 0x0062);
            const uint32_t o13h = prmt(u1, u3, 0x0062), o57h = prmt(u5, u7, 0x0062);
            st0 = make_uint2(prmt(e02, e46, 0x5410), prmt(o13, o57, 0x5410));
            st1 = make_uint2(prmt(e02, e46, 0x7632), prmt(o13, o57, 0x7632));
            st2 = make_uint2(prmt(e02h, e46h, 0x5410), prmt(o13h, o57h, 0x5410));
          }
          if (ok) {
            *reinterpret_cast<uint2*>(pl0 + (size_t)v * 8) = st0;
            *reinterpret_cast<uint2*>(pl1 + (size_t)v * 8) = st1;
            *reinterpret_cast<uint2*>(pl2 + (size_t)v * 8) = st2;
            if ((lane & 15) == 0) {
              gt_sm[(v >> 4) * 4 + m] = gs;
              gt_sm[(v >> 4) * 4 + 2 + m] = sum;
            }
          }
        }
      }
    }
  } else
  {
    // stage the activations once per CTA as (a0 - a1/16, a1/16) pairs in fragment order, and per scale group and
    // batch row sum_k a_k / 64 as an fp16 (hi, lo) pair for the zero-point MMA.  All loads of a thread are issued
    // before the first use: what costs here is the L2 round trip, not the arithmetic.
    const int vecs = a.K >> 3;                      // 8-half vectors per row, a multiple of 16
    constexpr int kBatch = 4;
    for (int m = 0; m < a.M; ++m) {
      const uint4* arow = reinterpret_cast<const uint4*>(a.a + (size_t)m * a.K);
      __half* srow = act_sm + (size_t)m * pitch;
      for (int v0 = warp * 32; v0 < vecs; v0 += kBatch * kPConsumerThreads) {
        uint4 val[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const int v = v0 + b * kPConsumerThreads + lane;
          val[b] = make_uint4(0, 0, 0, 0);
          if (v < vecs) val[b] = a_is_ll ? p_ll_load8(reinterpret_cast<const unsigned long long*>(a.a) + (((size_t)m * a.K) >> 1) + 4 * (size_t)v, ll_in_epoch, a.ll_state + 3)
                                           : __ldcg(arow + v);  // L2 only: may just have been written by the previous kernel or a peer GPU
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const int vw = v0 + b * kPConsumerThreads;                // warp-uniform
          if (vw >= vecs) break;
          const int v = vw + lane;
          const bool ok = v < vecs;
          if (ok) *reinterpret_cast<uint4*>(srow + v * 8) = permute_act8_v2(val[b]);
          const float2 f0 = __half22float2(u2h2(val[b].x)), f1 = __half22float2(u2h2(val[b].y));
          const float2 f2 = __half22float2(u2h2(val[b].z)), f3 = __half22float2(u2h2(val[b].w));
          float sum = ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
          // a scale group = 4 * UPG consecutive vectors = consecutive lanes
#pragma unroll
          for (int o = 1; o < 4 * UPG; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          if (ok && (lane & (4 * UPG - 1)) == 0) {
            const float q64 = sum * 0.015625f;
            const __half hi16 = __float2half_rn(q64);
            const __half lo16 = __float2half_rn(q64 - __half2float(hi16));
            *reinterpret_cast<uint4*>(zt_sm + ((size_t)(v / (4 * UPG)) * a.M + m) * 4) =
                make_uint4((uint32_t)__half_as_ushort(hi16) | ((uint32_t)__half_as_ushort(lo16) << 16), 0u, 0u, 0u);
          }
        }
      }
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kPConsumerThreads) : "memory");
  if (tid == 0) P_TRACE(3);

#ifdef XBIT_DEVTOOLS
  const long long loop0 = a.trace ? clock64() : 0;
  long long wait_clk = 0;
  bool first_wait = true;
#endif

  for (int pi = 0; pi < count; ++pi) {
  const W4PProblem& P = a.prob[pi];
  if (pi > 0) slice_range(P, false);
  // (every slice non-empty and all of them in one tile: see the tile epilogue)
  const bool one_tile = hi - lo >= kPWarps && tile_of(lo) == tile_of(hi - 1);
  int* const bnd = bnd_sm + pi * 32;
  while (j < jend) {
    const int cnt = min(jend - j, nb - kb);         // this warp's blocks of `tile`: [kb, kb + cnt)
    float tot[2][4];
#pragma unroll
    for (int tt = 0; tt < 2; ++tt)
#pragma unroll
      for (int i = 0; i < 4; ++i) tot[tt][i] = 0.f;
    float toti[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, zci[2] = {0.f, 0.f};   // integer block math (I8)

    for (int i = 0; i < cnt; i += 2) {
      const int s0 = s, ph0 = ph;
      if (++s == R) { s = 0; ph ^= 1; }
      const bool two_slots = BPS == 1 && i + 2 <= cnt;              // the step's second block sits in the next slot
      const int s1 = s, ph1 = ph;
      if (two_slots) { if (++s == R) { s = 0; ph ^= 1; } }
#ifdef XBIT_DEVTOOLS
      const long long w0c = a.trace ? clock64() : 0;
      if (first_wait && tid == 0) P_TRACE(10);
#endif
      mbar_wait(&my_full[s0], ph0);
      if (two_slots) mbar_wait(&my_full[s1], ph1);
#ifdef XBIT_DEVTOOLS
      if (a.trace) wait_clk += clock64() - w0c;
      if (first_wait && tid == 0) P_TRACE(4);
      first_wait = false;
      if (a.debug_skip != 1)
#endif
      {
        const unsigned char* const w0 = my_w + s0 * Cfg::kWSlot;
        const unsigned char* const sc0 = my_s + s0 * Cfg::kSSlot;
        const unsigned char* const z0 = my_z + s0 * Cfg::kZSlot;
        const unsigned char* const w1 = BPS == 2 ? w0 + Cfg::kBlockBytes : my_w + s1 * Cfg::kWSlot;      // the step's second block
        const unsigned char* const sc1 = BPS == 2 ? sc0 + GPB * 64 : my_s + s1 * Cfg::kSSlot;
        const unsigned char* const z1 = BPS == 2 ? z0 + GPB * Cfg::kZRow : my_z + s1 * Cfg::kZSlot;
        const __half* const a0 = act_sm + (kb + i) * 128;
        const unsigned char* const zt0 = zt_bytes + (size_t)(kb + i) * GPB * zt_group_bytes;
        if constexpr (I8) {
          const float* const g0 = gt_sm + (size_t)(kb + i) * 4;
          if (DUAL && i + 2 <= cnt) {
            const unsigned char* const wp[2] = {w0, w1};
            const unsigned char* const sp[2] = {sc0, sc1};
            const unsigned char* const zp[2] = {z0, z1};
            const int wr[2] = {(kb + i) * 16, (kb + i + 1) * 16};
            const float* const gp[2] = {g0, g0 + 4};
            wxp_consume_i8<BITS, 2, BF>(wp, sp, zp, wr, gp, LI, toti, zci);
          } else {
            const unsigned char* const wp[1] = {w0};
            const unsigned char* const sp[1] = {sc0};
            const unsigned char* const zp[1] = {z0};
            const int wr[1] = {(kb + i) * 16};
            const float* const gp[1] = {g0};
            wxp_consume_i8<BITS, 1, BF>(wp, sp, zp, wr, gp, LI, toti, zci);
            if (!DUAL && i + 2 <= cnt) {
              const unsigned char* const wp1[1] = {w1};
              const unsigned char* const sp1[1] = {sc1};
              const unsigned char* const zp1[1] = {z1};
              const int wr1[1] = {(kb + i + 1) * 16};
              const float* const gp1[1] = {g0 + 4};
              wxp_consume_i8<BITS, 1, BF>(wp1, sp1, zp1, wr1, gp1, LI, toti, zci);
            }
          }
        } else if (DUAL && i + 2 <= cnt) {
          const unsigned char* const wp[2] = {w0, w1};
          const unsigned char* const sp[2] = {sc0, sc1};
          const unsigned char* const zp[2] = {z0, z1};
          const __half* const ap[2] = {a0, a0 + 128};
          const unsigned char* const ztp[2] = {zt0, zt0 + GPB * zt_group_bytes};
          w4p_consume<UPG, 2>(wp, sp, zp, ap, ztp, zt_group_bytes, L, tot);
        } else {
          const unsigned char* const wp[1] = {w0};
          const unsigned char* const sp[1] = {sc0};
          const unsigned char* const zp[1] = {z0};
          const __half* const ap[1] = {a0};
          const unsigned char* const ztp[1] = {zt0};
          w4p_consume<UPG, 1>(wp, sp, zp, ap, ztp, zt_group_bytes, L, tot);
          if (!DUAL && i + 2 <= cnt) {
            const unsigned char* const wp1[1] = {w1};
            const unsigned char* const sp1[1] = {sc1};
            const unsigned char* const zp1[1] = {z1};
            const __half* const ap1[1] = {a0 + 128};
            const unsigned char* const ztp1[1] = {zt0 + GPB * zt_group_bytes};
            w4p_consume<UPG, 1>(wp1, sp1, zp1, ap1, ztp1, zt_group_bytes, L, tot);
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&my_empty[s0]);
        if (two_slots) mbar_arrive(&my_empty[s1]);
      }
    }

    if (tid == 0) P_TRACE(5);
    // ---- this warp's piece [j, j + cnt) of `tile` is done: partial tile -> shared memory; whoever completes the
    // CTA's portion of the tile sums the warps' partials in slice order and stores / publishes / finishes it
    const int t0 = tile * nb;
    const int p0 = max(lo, t0), p1 = min(hi, t0 + nb);              // the CTA's portion of the tile
    const bool is_first = (j == p0);                                // this piece starts the portion
    auto part_of = [&](int sl, int first) { return part_sm + (size_t)pi * part_stride + (size_t)((sl * 2 + first) * a.M) * 32; };
    {
      float* mine = part_of(rho, is_first ? 1 : 0);
      if constexpr (I8) {
        // lanes (t, t ^ 1) hold (D0 + 256 D1) and 65536 (D2 - 64 S) of activation row t / 2: the even one stores the sum,
        // less the zero-point term of that column and row, which lane (g, 2 tt + h) holds (no read-modify-write of the
        // partial tile in shared memory: this runs once per piece, on the critical path of every call)
        const int row = r >> 1;
#pragma unroll
        for (int tt = 0; tt < 2; ++tt)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float v = (r & 1) ? toti[tt][h] * 65536.f : toti[tt][h];
            const float y = v + __shfl_xor_sync(0xffffffffu, v, 1);
            const int src = (lane & ~3) | (2 * tt + h);
            const float z0 = __shfl_sync(0xffffffffu, zci[0], src);
            const float z1 = a.M > 1 ? __shfl_sync(0xffffffffu, zci[1], src) : 0.f;
            if ((r & 1) == 0 && row < a.M) mine[row * 32 + 4 * c8 + 2 * tt + h] = y - (row ? z1 : z0);
          }
      } else {
#pragma unroll
        for (int tt = 0; tt < 2; ++tt)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int m = 2 * r + (q & 1);
            const int col = 4 * c8 + 2 * tt + (q >> 1);
            if (m < a.M) mine[m * 32 + col] = tot[tt][q] * 16777216.f;
          }
      }
    }
    __syncwarp();
#ifdef W4P_TRACE_EXTRA
    if (tid == 0) P_TRACE(8);
#endif
    unsigned int held;
    int sl_first;
    bool finalize = true;
    if (one_tile) {
      // The CTA's whole range lies in this tile and every warp holds a piece of it (the tile-aligned schedule with one tile
      // per CTA: five of the seven matrices of a Llama-2-7B layer): the warp of slice 0 waits on a named barrier for the
      // others' arrivals and sums -- no slice search, no counter (0.2 us of every call before)
      held = (1u << kPWarps) - 1u;
      sl_first = 0;
      if (rho != 0) {
        asm volatile("bar.arrive %0, %1;" ::"r"(2 + (pi & 3)), "n"(kPConsumerThreads) : "memory");
        finalize = false;
      } else {
        asm volatile("bar.sync %0, %1;" ::"r"(2 + (pi & 3)), "n"(kPConsumerThreads) : "memory");
      }
    } else {
      // slices of the CTA that hold blocks of the portion (a slice can be empty when the CTA has fewer blocks than
      // slices): lane l looks at slice l -- one round of shared-memory loads and a ballot instead of loops over the slices
      const int b_lo = bnd[min(lane, kPWarps)], b_hi = bnd[min(lane + 1, kPWarps)];
      held = __ballot_sync(0xffffffffu, lane < kPWarps && min(b_hi, p1) > max(b_lo, p0));
      sl_first = __ffs((int)held) - 1;
#ifdef W4P_TRACE_EXTRA
      if (tid == 0) P_TRACE(9);
#endif
      if (held & (held - 1u)) {                                     // more than one slice: the last to arrive finalizes
        // (one acquire-release atomic at CTA scope instead of fence.sc + atomic + fence.sc: the __syncwarp before it orders
        // this warp's partial tile ahead of lane 0's release, the one after it orders the other lanes' loads behind its acquire)
        uint32_t old = 0;
        if (lane == 0)
          asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&cnt_sm[pi * 16 + sl_first])) : "memory");
        old = __shfl_sync(0xffffffffu, old, 0);
        __syncwarp();
        finalize = ((int)old == __popc(held) - 1);
      }
    }
#ifdef W4P_TRACE_EXTRA
    if (tid == 0) P_TRACE(11);
#endif
    if (finalize) {
      const bool starts_tile = (p0 == t0), ends_tile = (p1 == t0 + nb);
      const int n = tile * 32 + lane;                               // this lane's output column
      for (int m = 0; m < a.M; ++m) {
        // the slices' partial tiles in slice order; all loads issued together (slices outside the portion add 0)
        float v = 0.f;
#pragma unroll
        for (int sl = 0; sl < kPWarps; ++sl) {
          float x = 0.f;
          if ((held >> sl) & 1u) x = part_of(sl, sl == sl_first ? 1 : 0)[m * 32 + lane];
          v += x;
        }
        if (starts_tile && !ends_tile) {
          // finisher: the CTAs after this one that hold the tile's later blocks published their parts (normally long ago)
          // (the CTAs whose ranges begin inside this tile.  On matrices of 120 MB and more this loop and the 64-bit division it
          // replaced put the back-to-back launches into different steady states -- late CTAs start the next launch late,
          // with empty rings, and finish it late again: 25.9 against 22.3 us on 8192 x 28672,
          // profiles/r02_trace_persist_8192x28672_two_regimes.log; the 7B shapes gain 0.1..0.2 us from this form)
          for (int cc = c + 1; cc < G && range_lo(P, cc) < t0 + nb; ++cc) {
            unsigned long long* slot = P.ws + ((size_t)cc * a.M + m) * 32 + lane;
            uint32_t bits, flag;
            const long long c0 = clock64();
            for (unsigned int spin = 1;; ++spin) {
              asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(bits), "=r"(flag) : "l"(slot) : "memory");
              if (flag == 1u) break;
              if ((spin & 1023u) == 0 && clock64() - c0 > kPSpinGuardClocks) { bits = 0x7fc00000u; break; }   // NaN: visible, not silent
            }
            v += __uint_as_float(bits);
            asm volatile("st.volatile.global.v2.u32 [%0], {%1, %1};" ::"l"(slot), "r"(0u) : "memory");      // left zeroed for the next call
          }
        }
        if (starts_tile) {
          const __half h = BF ? __ushort_as_half(__bfloat16_as_ushort(__float2bfloat16_rn(v))) : __float2half_rn(v);   // (16 result bits)
          const size_t off = (size_t)m * P.ldo + P.col_offset + n;
          if (ll_out) {
            // flag-in-data all-gather: each pair of results goes to every rank as one {half2, call number} store
            const uint32_t lo16 = (uint32_t)__half_as_ushort(h);
            const uint32_t hi16 = __shfl_down_sync(0xffffffffu, lo16, 1);
            if ((lane & 1) == 0) {
              const uint32_t epoch = a.ll_state[2] + (uint32_t)a.ll_chain_index + 1u;
              for (int p = 0; p < a.world; ++p) p_ll_store(reinterpret_cast<unsigned long long*>(P.out[p]) + (off >> 1), lo16 | (hi16 << 16), epoch);
            }
          } else {
            P.out[0][off] = h;
            for (int p = 1; p < a.world; ++p) P.out[p][off] = h;    // fused all-gather: NVLink peer stores
          }
        } else {
          // contributor: this CTA holds later blocks of a tile that starts in an earlier CTA
          asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(P.ws + ((size_t)c * a.M + m) * 32 + lane), "r"(__float_as_uint(v)), "r"(1u) : "memory");
        }
      }
    }
    if (tid == 0) P_TRACE(6);
    j += cnt;
    kb = 0;
    ++tile;
  }
  }
#ifdef XBIT_DEVTOOLS
  if (tid == 0 && a.trace) {
#ifndef W4P_TRACE_EXTRA
    P_TRACE_VALUE(8, (unsigned long long)(clock64() - loop0));
    P_TRACE_VALUE(9, (unsigned long long)wait_clk);
    P_TRACE_VALUE(11, (unsigned long long)((range_lo(a.prob[0], c + 1) - range_lo(a.prob[0], c)) / kPWarps));
#endif
    P_TRACE(7);
    p_trace_dump(a);
  }
#endif
}

// ------------------------------------------------------------------------------------------------ host side

static int p_upg_of(int groupsize) { return groupsize == 32 ? 1 : (groupsize == 64 ? 2 : 4); }

struct W4PPlan {
  int grid, unit, ring, nw, i8, minb;
  bool preferred;
  size_t smem;
};

static size_t w4p_smem_bytes(int upg, int nr, int m, int k, int ring, bool i8, int bps, int count = 1, int bits = 4) {
  const int gpb = 4 / upg;
  const size_t acts = i8 ? (size_t)(k / 128) * 16 + 16 + (size_t)3 * m * (k + 128) + 128      // group table, constants, digit planes
                         : (size_t)(k / 128) * gpb * m * 16 + (size_t)m * (k + 8) * sizeof(__half);   // zt_sm, act_sm
  return 1024                                                        // alignment slack
         + (size_t)nr * ring * bps * (512 * bits + (64 + (bits == 8 ? 32 : 16)) * gpb)   // rings of bps-block slots
         + (size_t)2 * nr * kPMaxRing * 8 + kPMaxProblems * 48 * 4   // mbarriers, counters, slice boundaries
         + (size_t)count * nr * 2 * m * 32 * sizeof(float)            // part_sm
         + acts;
}

size_t gemv_w4p_workspace_bytes(int M) {
  return (size_t)device_sm_count() * (size_t)(M < 1 ? 1 : (M > 8 ? 8 : M)) * 32 * sizeof(unsigned long long);
}

// 2- and 8-bit weights on this kernel (A16W2 / A16W8): groupsize 128, M <= 2 (integer block math only), one matrix per launch
static bool w8p_supported(const GemvArgs& a) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(a.a) | reinterpret_cast<uintptr_t>(a.qweight) |
                       reinterpret_cast<uintptr_t>(a.scales) | reinterpret_cast<uintptr_t>(a.qzeros);
  return (a.bits == 8 || a.bits == 2) && a.groupsize == 128 && a.K % 128 == 0 && a.N % 32 == 0 && (al & 15u) == 0 && a.M >= 1 && a.M <= 2 &&
         !a.ll_out && !a.a_is_ll && a.world <= 1;
}

static bool plan_w4p_nw(const GemvArgs& a, bool have_ws, W4PPlan& p, bool allow16, int count, long long share_all) {
  const bool w8 = w8p_supported(a) && count == 1 && env_int("XBIT_W4P_I8", 1) != 0;
  if ((!gemv_w4_supported(a) && !w8) || a.M > 8) return false;
  const int sms = device_sm_count();
  const int upg = p_upg_of(a.groupsize);
  const long long tiles = a.N / 32, nb = a.K / 128;
  if (tiles * nb > 0x3fffffffLL) return false;
  // Two shapes of CTA, consumer warps with a ring each that take two blocks per step (measured on the Llama shapes,
  // profiles/r02_ptime_*; one block at a time, 12 warps, 16 warps sharing rings in pairs and 16 warps on half an SM
  // were slower everywhere they were tried):
  //   half SM:  8 warps, rings of 2..3 two-block slots, two launches co-resident: a matrix whose per-CTA share fits the
  //             rings is prefetched whole while the previous call computes (4096 x 4096: 3.4 us against 4.3 us for the
  //             cluster kernel);
  //   full SM:  rings of 4..6 slots (16 warps: 2), one CTA per SM and the register budget of one: larger matrices keep
  //             streaming while they compute, and what counts is the compute rate and the bytes in flight
  //             (8192 x 28672: 22.3 against 26.0 us).
  const long long total_blocks = tiles * nb;
  const long long share = (total_blocks + sms - 1) / sms;           // blocks per CTA
  // (a launch of several matrices, xbit_gemv_f16_multi: the CTA streams the shares of all of them, share_all)
  // half SM: the share fits 8 rings of 3 two-block slots AND no CTA gets more than one tile (4096 x 5504, 172 tiles: 4.55 us
  // on a full SM with rings of 6 against 5.1 us on half an SM, profiles/r02_ptime_half_vs_full_sm_mid_shapes.log)
  const bool small = (share_all > 0 ? share_all : share) <= 8 * 2 * 3 && (share_all > 0 || tiles <= sms);
  // (16 consumer warps -- rings of 2 slots -- paid off from about 35 MB while 8-warp rings stopped at 4 slots; with rings of
  // 5..6 slots 8 warps are ahead everywhere: 8192 x 8192 8.0 against 8.3 us, 8192 x 28672 23.6 against 25.9,
  // profiles/r02_ptime_8_vs_16_warps_deep_rings.log.  XBIT_W4P_WARPS=16 keeps the form reachable for tools/ptime.py.)
  const int env_nw = env_int("XBIT_W4P_WARPS", 0);
  p.nw = (!w8 && env_nw == 16 && allow16) ? 16 : 8;
  const int nr = p.nw;                               // rings
  // integer block math: groupsize 128, M <= 2 (XBIT_W4P_I8=0: the fp16 exact-product math everywhere)
  const bool i8 = w8 || (a.groupsize == 128 && a.M <= 2 && env_int("XBIT_W4P_I8", 1) != 0);
  p.i8 = i8 ? 1 : 0;
  if (a.bf16 && (!i8 || count != 1 || a.ll_out || a.a_is_ll)) return false;   // the bf16-native form: integer block math, one matrix
  // CTA boundaries: block granular (perfect balance, tiles shared between CTAs meet in the workspace) or tile aligned
  // (nothing crosses CTAs).  Cost model in blocks per CTA; the cross-CTA fix-up is worth about 4 blocks of time.
  // Tile-aligned: always one CTA per SM, also when there are fewer tiles (CTAs without work hold their slot until the
  // previous launch has finished, so that the hardware never stacks two working CTAs of one launch on an SM: measured
  // 2.7 us against 1.7 us for the stacked ones, profiles/r02_*trace*)
  const long long total = tiles * nb;
  const long long g_fine = total < sms ? total : sms;
  // Cost model in blocks per CTA (8 warps x 4 blocks ~ 1 us), fitted to profiles/r02_ptime_shard_shapes.log:
  //   * a tile shared between CTAs costs its finisher a round trip through the workspace: ~1.3 us (42 blocks) when the
  //     CTA ranges are shorter than a tile -- the contributor's piece is then its WHOLE range and lands when the finisher
  //     is already waiting -- and ~0.5 us (16 blocks) when they are longer (the contributor does that piece first);
  //   * tile-aligned with fewer tiles than SMs leaves the stream to few SMs: whatever exceeds the rings of a half-SM CTA
  //     (48 blocks, prefetched during the previous call) arrives at one SM's share of the bandwidth.
  const long long per_cta_fine = (total + g_fine - 1) / g_fine;
  const bool fine_is_aligned = total % g_fine == 0 && (total / g_fine) % nb == 0;
  //   * with more than one activation row every shared tile also moves M rows of partial sums: 8 blocks per row
  //     (4096 x 11008: M = 2 7.1 us tile-aligned against 7.3 us, M = 4 8.9 against 10.3, M = 8 11.8 against 14.8)
  const long long cost_fine = per_cta_fine + (fine_is_aligned ? 0 : (per_cta_fine < nb ? 42 : 16) + (a.M > 1 ? 8 * a.M : 0));
  const long long cost_tile = (tiles + sms - 1) / sms * nb + ((tiles * 4 < sms * 3 && nb > 48) ? nb - 48 : 0);
  bool fine = have_ws && cost_fine < cost_tile;
  const int env_unit = env_int("XBIT_W4P_FINE", -1);
  if (env_unit == 0) fine = false;
  if (env_unit == 1 && have_ws) fine = true;
  p.unit = fine ? 1 : (int)nb;
  p.grid = (int)(fine ? g_fine : sms);
  const int env_grid = env_int("XBIT_W4P_GRID", 0);
  if (env_grid > 0 && (env_grid <= total || !fine)) p.grid = env_grid;
  // Rings: a share that fits 8 rings of 2..3 slots gets them on half an SM (the next call's CTA is co-resident);
  // everything else one CTA per SM with the deepest rings that fit (XBIT_W4P_RING overrides, tools/ptime.py)
  const size_t half = 113 * 1024;
  int ring = 0;
  if (small && nr == 8 && a.bits != 8)
    for (int r = 3; r >= 2 && !ring; --r)
      if (w4p_smem_bytes(upg, nr, a.M, a.K, r, i8, 2, count, a.bits) <= half) ring = r;
  // (8 rings: as deep as fits, up to 6 slots -- 5 against 4: 8.05 / 8.7 us on 8192 x 8192, 6.18 / 6.27 on 4096 x 11008,
  // profiles/r02_ptime_ring_5_6.log; 16 rings: 3 slots measured no better than 2)
  for (int r = (nr == 16 ? 2 : 6); r >= 2 && !ring; --r)
    if (w4p_smem_bytes(upg, nr, a.M, a.K, r, i8, 2, count, a.bits) <= kMaxDynSmem) ring = r;
  if (!ring) return false;
  const int env_ring = env_int("XBIT_W4P_RING", 0);
  if (env_ring >= 2 && env_ring <= kPMaxRing) ring = env_ring;
  if (w4p_smem_bytes(upg, nr, a.M, a.K, ring, i8, 2, count, a.bits) > kMaxDynSmem) return false;
  p.ring = ring;
  p.smem = w4p_smem_bytes(upg, nr, a.M, a.K, ring, i8, 2, count, a.bits);
  p.minb = (p.smem <= half && p.nw != 16) ? 2 : 1;
  // AUTO prefers this kernel where it was measured ahead of the cluster split-K kernel: shares that fit the rings (any
  // block math), and with the integer block math every matrix that gets the 4-slot rings
  p.preferred = (small && p.minb == 2) || (i8 && (ring >= 4 || (nr == 16 && ring >= 2)));
  // ... and NOT where the cluster kernel (split-K across a cluster, every SM streaming) stayed ahead: few tiles of a long
  // K (the column shards of a tensor-parallel layer: 8192 x 1024 3.8 against 5.6 us, 11008 x 1024 4.4 against 6.3), and
  // K beyond 16384, where staging the whole activation row per CTA costs more than it saves (28672 x 8192: 23.9 / 27.7)
  if (fine ? tiles < 64 : (nb > 32 && tiles < 100)) p.preferred = false;
  // batches of 3..8 rows (fp16 block math): ahead only where a CTA works through several whole tiles of a short K
  // (4096 x 11008: M = 4 8.9 against 10.0 us, M = 8 11.8 against 18.0 us; 4096 x 4096 and 8192 x 8192: behind or level)
  if (a.M > 2) p.preferred = !fine && nb <= 32 && tiles >= 2 * sms;
  if (a.K > 16384) p.preferred = false;
  if (w8) p.preferred = true;                        // (the alternative for 8-bit weights is the generic kernel)
  return true;
}

static bool plan_w4p(const GemvArgs& a, bool have_ws, W4PPlan& p, int count = 1, long long share_all = 0) {
  // sixteen consumer warps where they pay off and their rings fit next to the staged activations, eight otherwise
  return plan_w4p_nw(a, have_ws, p, true, count, share_all) || plan_w4p_nw(a, have_ws, p, false, count, share_all);
}

bool gemv_w4p_applicable(const GemvArgs& a) {
  W4PPlan p;
  return plan_w4p(a, false, p);
}

bool gemv_w4p_preferred(const GemvArgs& a) {
  W4PPlan p;
  return plan_w4p(a, true, p) && p.preferred;
}


// One launch for `count` weight matrices that share the activations (and M, K, bits, group size, zero bias): every CTA
// works through its range of each matrix in turn -- exactly the range, the warp slices and therefore the fp32 summation
// order of a separate call on that matrix, so the results are bit-identical to `count` separate calls -- with one kernel
// boundary, one activation staging and one prologue for all of them, and the rings running ahead across matrices.
// cudaErrorNotSupported: the matrices want different CTA shapes (the caller launches them one by one).
cudaError_t launch_gemv_w4p_multi(const GemvArgs* gs, int count, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (count < 1 || count > kPMaxProblems) return cudaErrorInvalidValue;
  GemvArgs g0 = gs[0];
  apply_debug_knobs(g0);
  const size_t region = gemv_w4p_workspace_bytes(g0.M);
  const bool have_ws = workspace && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0 && workspace_bytes >= region * count;
  const int sms = device_sm_count();
  long long share_all = 0;
  for (int i = 0; i < count; ++i) {
    if (gs[i].M != g0.M || gs[i].K != g0.K || gs[i].groupsize != g0.groupsize || gs[i].zero_bias != g0.zero_bias || gs[i].a != g0.a ||
        gs[i].world != g0.world)
      return cudaErrorInvalidValue;
    share_all += ((long long)(gs[i].N / 32) * (g0.K / 128) + sms - 1) / sms;
  }
  W4PPlan p;                                         // CTA shape of the launch: from the first matrix and the combined share
  if (!plan_w4p(g0, have_ws, p, count, count > 1 ? share_all : 0)) return cudaErrorInvalidValue;
  const int upg = p_upg_of(g0.groupsize);
  W4PArgs a;
  memset(&a, 0, sizeof(a));
  alignas(64) W4PMaps maps;
  memset(&maps, 0, sizeof(maps));
  a.a = g0.a;
  a.world = g0.world;
  a.M = g0.M; a.K = g0.K; a.zero_bias = g0.zero_bias;
  a.nb = g0.K / 128;
  a.ring = p.ring;
  a.static_weights = g0.static_weights;
  a.all_wait = env_int("XBIT_W4P_ALLWAIT", 0);
  a.prefetch_delay = env_int("XBIT_W4P_DELAY", 0);
  a.count = count;
  a.nb_shift = 31;
  while ((1ll << (a.nb_shift - 31)) < a.nb) ++a.nb_shift;          // 2^nb_shift > 2^30 * nb >= j * nb
  a.nb_magic = (unsigned int)(((1ull << a.nb_shift) + (unsigned)a.nb - 1) / (unsigned)a.nb);
  a.ll_out = g0.ll_out;
  a.a_is_ll = g0.a_is_ll;
  a.ll_chain_index = g0.ll_chain_index;
  a.ll_state = g0.sig_state;
  a.trace = g0.trace;
  a.debug_skip = g0.debug_skip;
  for (int i = 0; i < count; ++i) {
    const GemvArgs& g = gs[i];
    W4PPlan pi;                                      // decomposition of this matrix: what a separate call would use
    if (!plan_w4p(g, have_ws, pi)) return cudaErrorInvalidValue;
    if (count > 1 && (pi.nw != p.nw || pi.i8 != p.i8)) {
      // the launch's shape came from matrix 0 and the combined share: take matrix i's own warp count if all agree on it
      if (i == 0) p.nw = pi.nw;
      else return cudaErrorNotSupported;
    }
    if (count > 1 && pi.grid != (pi.unit == 1 ? (int)std::min<long long>((long long)(g.N / 32) * a.nb, sms) : sms)) return cudaErrorNotSupported;
    if (i == 0) p.grid = pi.grid;
    else if (pi.grid != p.grid) return cudaErrorNotSupported;
    W4PProblem& P = a.prob[i];
    P.scales = g.scales;
    P.qzeros = g.qzeros;
    for (int r = 0; r < kMaxPeers; ++r) P.out[r] = g.out[r];
    P.ldo = g.ldo;
    P.col_offset = g.col_offset;
    P.N = g.N;
    P.zwords = g.zwords;
    P.total = (g.N / 32) * a.nb;
    P.unit = pi.unit;
    P.uq = (P.total / pi.unit) / p.grid;
    P.ur = (P.total / pi.unit) % p.grid;
    P.ws = pi.unit == 1 ? reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(workspace) + (size_t)i * region) : nullptr;
    // qweight [qrows, N] u32: box = two blocks (32 word-rows) x 32 columns (128 B), 128-byte swizzle; one block for odd tails
    cudaError_t e = encode_2d(&maps.m[2 * i], CU_TENSOR_MAP_DATA_TYPE_UINT32, g.qweight, (uint64_t)g.N, (uint64_t)g.qrows, (uint64_t)g.N * 4, 32, (uint32_t)(8 * g.bits),
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (e != cudaSuccess) return e;
    e = encode_2d(&maps.m[2 * i + 1], CU_TENSOR_MAP_DATA_TYPE_UINT32, g.qweight, (uint64_t)g.N, (uint64_t)g.qrows, (uint64_t)g.N * 4, 32, (uint32_t)(4 * g.bits),
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (e != cudaSuccess) return e;
  }
  if (count > 1) {
    // the warp count may have been replaced by the matrices' own: shared memory and register budget follow it
    W4PPlan q = p;
    GemvArgs probe = g0;
    if (!plan_w4p_nw(probe, have_ws, q, p.nw == 16, count, share_all) || q.nw != p.nw) return cudaErrorNotSupported;
    p.ring = q.ring; p.smem = q.smem; p.minb = q.minb;
    a.ring = p.ring;
  }
  a.g_magic = p.grid <= 1 ? 0xffffffffu : (unsigned int)(((1ull << 32) + (unsigned)p.grid - 1) / (unsigned)p.grid);
  const void* kern = nullptr;
#define XBIT_W4P_CASE(UPG_, I8_, GEN_)                                                                  \
  if (upg == UPG_ && (p.i8 != 0) == I8_) {                                                              \
    if (p.nw == 16) kern = (const void*)gemv_w4p_kernel<UPG_, 16, 1, I8_, 2, 1, GEN_>;                  \
    else if (p.minb == 2) kern = (const void*)gemv_w4p_kernel<UPG_, 8, 1, I8_, 2, 2, GEN_>;             \
    else kern = (const void*)gemv_w4p_kernel<UPG_, 8, 1, I8_, 2, 1, GEN_>;                              \
  }
#define XBIT_W4P_GEN(GEN_)                                                                              \
  XBIT_W4P_CASE(1, false, GEN_) XBIT_W4P_CASE(2, false, GEN_) XBIT_W4P_CASE(4, false, GEN_) XBIT_W4P_CASE(4, true, GEN_)
  const int gen = (a.ll_out || a.a_is_ll) ? 2 : (count > 1 ? 1 : 0);
  if (g0.bits == 8) {
    kern = g0.bf16 ? (const void*)gemv_w4p_kernel<4, 8, 1, true, 2, 1, 0, true, 8> : (const void*)gemv_w4p_kernel<4, 8, 1, true, 2, 1, 0, false, 8>;
  } else if (g0.bits == 2) {
    if (p.minb == 2) kern = g0.bf16 ? (const void*)gemv_w4p_kernel<4, 8, 1, true, 2, 2, 0, true, 2> : (const void*)gemv_w4p_kernel<4, 8, 1, true, 2, 2, 0, false, 2>;
    else kern = g0.bf16 ? (const void*)gemv_w4p_kernel<4, 8, 1, true, 2, 1, 0, true, 2> : (const void*)gemv_w4p_kernel<4, 8, 1, true, 2, 1, 0, false, 2>;
  } else if (g0.bf16) {
    if (p.nw == 16) kern = (const void*)gemv_w4p_kernel<4, 16, 1, true, 2, 1, 0, true>;
    else if (p.minb == 2) kern = (const void*)gemv_w4p_kernel<4, 8, 1, true, 2, 2, 0, true>;
    else kern = (const void*)gemv_w4p_kernel<4, 8, 1, true, 2, 1, 0, true>;
  } else if (gen == 0) { XBIT_W4P_GEN(0) }
  else if (gen == 1) { XBIT_W4P_GEN(1) }
  else { XBIT_W4P_GEN(2) }
#undef XBIT_W4P_GEN
#undef XBIT_W4P_CASE
  if (!kern) return cudaErrorInvalidValue;
  cudaError_t e = ensure_max_dyn_smem(kern);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)p.grid, 1, 1);
  cfg.blockDim = dim3((unsigned)(p.nw + 1) * 32, 1, 1);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  // the single-matrix instantiation takes the short parameter forms: prefixes of the long ones
  static_assert(offsetof(W4PArgsN<1>, prob) == offsetof(W4PArgs, prob), "short form must be a prefix");
  alignas(64) W4PMapsN<1> maps1;
  W4PArgsN<1> a1;
  void* params[2] = {&maps, &a};
  if (gen == 0) {
    memcpy(&maps1, &maps, sizeof(maps1));
    memcpy(&a1, &a, sizeof(a1));
    params[0] = &maps1;
    params[1] = &a1;
  }
  return cudaLaunchKernelExC(&cfg, kern, params);
}

cudaError_t launch_gemv_w4p(const GemvArgs& g, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  return launch_gemv_w4p_multi(&g, 1, workspace, workspace_bytes, stream);
}

}  // namespace xbit
