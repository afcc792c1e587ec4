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
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "gemv_prims.cuh"
#include "unpack.cuh"
#include "xbit_internal.h"
#include "../../include/xbitops_b200.h"

namespace xbit {

constexpr int kPMaxWarps = 16;                      // consumer warps per CTA: 8 (two blocks per step), 12 or 16 (one)
constexpr int kPMaxRing = 8;
constexpr long long kPSpinGuardClocks = 60000000000ll;   // ~30 s: see kSpinGuardClocks in gemv_sm100.cu

struct W4PArgs {
  const __half* a;          // [M, K]
  __half* out[kMaxPeers];   // world output buffers ([M, ldo] each)
  int world;
  long long ldo, col_offset;
  int M, K, N, zero_bias;
  int nb;                   // 128-k blocks per tile = K / 128
  int total;                // tiles * nb
  int unit;                 // CTA boundaries are multiples of `unit` blocks: 1 (stream-K, needs ws) or nb (tile aligned)
  int ring;                 // two-block slots per consumer warp
  int static_weights;
  int all_wait;             // every consumer warp executes griddepcontrol.wait (comparison knob)
  int prefetch_delay;       // SM clocks the producer waits before its first request (only when it starts ahead of the wait)
  const __half* scales;     // [G, N]: copied into the rings with cp.async (64 bytes per block and group: too small for TMA requests)
  const uint32_t* qzeros;   // [G, zwords]
  int zwords;
  unsigned long long* ws;   // [grid][M][32] {fp32 partial, flag} slots, zero outside a launch
  unsigned long long* trace;
  int debug_skip;
};

template <int UPG>
struct W4PCfg {
  static constexpr int GPB = 4 / UPG;                                   // scale groups per 128-k block
  static constexpr int kBlockBytes = 2048;                              // 16 word-rows x 32 columns
  static constexpr int kWSlot = 2 * kBlockBytes;                        // a ring slot holds the two blocks of a step
  static constexpr int kSSlot = 2 * GPB * 64;                           // their scale rows (32 columns x fp16)
  static constexpr int kZSlot = 2 * GPB * 16;                           // their zero rows (4 words)
};

__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
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
__device__ __forceinline__ void p_trace_stamp(const W4PArgs& a, int slot) {
  if (a.trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[(size_t)blockIdx.x * 16 + slot] = t;
  }
}
__device__ __forceinline__ void p_trace_value(const W4PArgs& a, int slot, unsigned long long v) {
  if (a.trace) a.trace[(size_t)blockIdx.x * 16 + slot] = v;
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

// slice (eighth of the CTA's range) of consumer warp w: warps w and w + 4 share an SM sub-partition and adjacent
// slices differ by at most one block, so every sub-partition gets two ADJACENT slices
template <int NW>
__device__ __forceinline__ int p_slice_of(int w) { return (w & 3) * (NW / 4) + (w >> 2); }

// NW consumer warps + one producer warp.  MODE 0: a ring per warp, the two blocks of a step one after the other;
// MODE 1 (DUAL): a ring per warp, the two blocks together (two accumulator sets); MODE 2 (PAIR): a ring per PAIR of
// warps, warp 2p takes the first block of every step and warp 2p+1 the second -- 16 consumer warps at 56 registers
// share the 8 rings (and the shared memory) of the 8-warp forms.  Two CTAs (of consecutive launches) per SM.
template <int UPG, int NW, int MODE>
__global__ void __launch_bounds__((NW + 1) * 32, 2)
gemv_w4p_kernel(const __grid_constant__ CUtensorMap wmap2, const __grid_constant__ CUtensorMap wmap1, const W4PArgs a) {
  using Cfg = W4PCfg<UPG>;
  constexpr int GPB = Cfg::GPB;
  constexpr bool DUAL = MODE == 1, PAIR = MODE == 2;
  constexpr int kPWarps = PAIR ? NW / 2 : NW;       // rings = slices of the CTA's range
  constexpr int kPConsumerThreads = NW * 32;
  constexpr int LPR = kPWarps <= 8 ? 4 : 2;         // producer lanes per ring
  auto slice_of_ring = [](int rg) { return PAIR ? (rg & 1) * (kPWarps / 2) + (rg >> 1) : p_slice_of<kPWarps>(rg); };
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
  int* cnt_sm = reinterpret_cast<int*>(empty_bar + kPWarps * kPMaxRing);              // [32] arrival counters of shared tiles
  int* bnd_sm = cnt_sm + 32;                                                          // [slices + 1] first block of every slice
  float* part_sm = reinterpret_cast<float*>(bnd_sm + 32);                             // [slices][2 warps][2][2][M][32] partial tiles
  uint32_t* zt_sm = reinterpret_cast<uint32_t*>(part_sm + kPWarps * 8 * a.M * 32);    // [groups][M][4]: (hi, lo) of sum_k a_k / 64, 3 zero words
  __half* act_sm = reinterpret_cast<__half*>(zt_sm + (size_t)nb * GPB * a.M * 4);     // [M][pitch]

  // this CTA's range of the tile-major block list
  const long long U = a.total / a.unit;
  const int lo = (int)(U * c / G) * a.unit, hi = (int)(U * (c + 1) / G) * a.unit;
  const int len = hi - lo;

  if (tid == 0) {
    P_TRACE(0);
#ifdef XBIT_DEVTOOLS
    {
      unsigned int smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      P_TRACE_VALUE(12, (unsigned long long)smid + 1);
    }
#endif
    for (int w = 0; w < kPWarps; ++w)
      for (int s = 0; s < R; ++s) {
        mbar_init(&full_bar[w * kPMaxRing + s], 1 + LPR);   // the TMA issuer's expect_tx arrival + LPR cp.async arrivals
        mbar_init(&empty_bar[w * kPMaxRing + s], PAIR ? 2 : 1);
      }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    cnt_sm[tid] = 0;
    if (tid <= kPWarps) bnd_sm[tid] = lo + (int)(((long long)len * tid) / kPWarps);
  }
  __syncthreads();
  // the next kernel in the stream may become resident now: its producer streams ITS weights while this one runs
  griddep_launch_dependents();

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
      int j = lo + (int)((long long)len * rho / kPWarps);
      const int jend = lo + (int)((long long)len * (rho + 1) / kPWarps);
      int tile = j / nb, kb = j - tile * nb;
      uint64_t policy = 0;
      if (sub == 0) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      if (lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap1) : "memory");
        P_TRACE(1);
      }
      const unsigned char* sbase = reinterpret_cast<const unsigned char*>(a.scales) + sub * 16;
      const unsigned char* zbase = reinterpret_cast<const unsigned char*>(a.qzeros);
      int s = 0, ph = 0;
      for (int n = 0; j < jend; ++n) {
        const int nblk = min(2, min(jend - j, nb - kb));
        uint64_t* fb = &full_bar[w * kPMaxRing + s];
        if (n >= R) mbar_wait(&empty_bar[w * kPMaxRing + s], ph ^ 1);
        if (sub == 0) {
          mbar_arrive_expect_tx(fb, (uint32_t)(nblk * Cfg::kBlockBytes));
          tma_load_2d(wring + (w * R + s) * Cfg::kWSlot, nblk == 2 ? &wmap2 : &wmap1, tile * 32, kb * 16, fb, policy);
        }
        unsigned char* sdst = sring + (w * R + s) * Cfg::kSSlot + sub * 16;
        unsigned char* zdst = zring + (w * R + s) * Cfg::kZSlot;
        const int rows = nblk * GPB, row0 = kb * GPB;
#pragma unroll
        for (int rr = 0; rr < 2 * GPB; ++rr)
          if (rr < rows) {
#pragma unroll
            for (int ch = 0; ch < 4; ch += LPR)
              cp_async_16(sdst + rr * 64 + ch * 16, sbase + ((size_t)(row0 + rr) * a.N + tile * 32) * 2 + ch * 16);
          }
#pragma unroll
        for (int rr = 0; rr < 2 * GPB; ++rr)
          if (rr < rows && (rr % LPR) == sub) cp_async_16(zdst + rr * 16, zbase + ((size_t)(row0 + rr) * a.zwords + tile * 4) * 4);
        cp_async_mbar_arrive_noinc(fb);
        j += nblk;
        kb += nblk;
        if (kb == nb) { kb = 0; ++tile; }
        if (++s == R) { s = 0; ph ^= 1; }
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

  const int rg = PAIR ? warp >> 1 : warp, hh = PAIR ? warp & 1 : 0;   // ring, and which block of a step this warp takes
  const int rho = slice_of_ring(rg);
  int j = lo + (int)((long long)len * rho / kPWarps);
  const int jend = lo + (int)((long long)len * (rho + 1) / kPWarps);
  int tile = j / nb, kb = j - tile * nb;
  int s = 0, ph = 0;
  uint64_t* const my_full = full_bar + rg * kPMaxRing;
  uint64_t* const my_empty = empty_bar + rg * kPMaxRing;
  const unsigned char* const my_w = wring + rg * R * Cfg::kWSlot;
  const unsigned char* const my_s = sring + rg * R * Cfg::kSSlot;
  const unsigned char* const my_z = zring + rg * R * Cfg::kZSlot;
  const unsigned char* const zt_bytes = reinterpret_cast<const unsigned char*>(zt_sm);
  // The activations are the only data produced by the previous kernel.  ONE warp waits for it; the others block on a
  // hardware barrier behind that warp: warps parked in griddepcontrol.wait were measured to slow the co-resident CTA of
  // the previous launch down (XBIT_W4P_ALLWAIT=1 restores the plain form for the comparison).
  if (a.all_wait || warp == 0) griddep_wait();
  if (!a.all_wait) asm volatile("bar.sync 1, %0;" ::"n"(kPConsumerThreads) : "memory");
  if (tid == 0) P_TRACE(2);
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
          if (v < vecs) val[b] = __ldcg(arow + v);  // L2 only: may just have been written by the previous kernel or a peer GPU
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
  bool first_wait = true;
#endif

  while (j < jend) {
    const int cnt = min(jend - j, nb - kb);         // this warp's blocks of `tile`: [kb, kb + cnt)
    float tot[2][4];
#pragma unroll
    for (int tt = 0; tt < 2; ++tt)
#pragma unroll
      for (int i = 0; i < 4; ++i) tot[tt][i] = 0.f;

    for (int i = 0; i < cnt; i += 2) {
      const int s0 = s, ph0 = ph;
      if (++s == R) { s = 0; ph ^= 1; }
      mbar_wait(&my_full[s0], ph0);
#ifdef XBIT_DEVTOOLS
      if (first_wait && tid == 0) P_TRACE(4);
      first_wait = false;
      if (a.debug_skip != 1)
#endif
      {
        const unsigned char* const w0 = my_w + s0 * Cfg::kWSlot;
        const unsigned char* const sc0 = my_s + s0 * Cfg::kSSlot;
        const unsigned char* const z0 = my_z + s0 * Cfg::kZSlot;
        const __half* const a0 = act_sm + (kb + i) * 128;
        const unsigned char* const zt0 = zt_bytes + (size_t)(kb + i) * GPB * zt_group_bytes;
        if (PAIR) {
          if (i + hh < cnt) {
            const unsigned char* const wp[1] = {w0 + hh * Cfg::kBlockBytes};
            const unsigned char* const sp[1] = {sc0 + hh * GPB * 64};
            const unsigned char* const zp[1] = {z0 + hh * GPB * 16};
            const __half* const ap[1] = {a0 + hh * 128};
            const unsigned char* const ztp[1] = {zt0 + hh * GPB * zt_group_bytes};
            w4p_consume<UPG, 1>(wp, sp, zp, ap, ztp, zt_group_bytes, L, tot);
          }
        } else if (DUAL && i + 2 <= cnt) {
          const unsigned char* const wp[2] = {w0, w0 + Cfg::kBlockBytes};
          const unsigned char* const sp[2] = {sc0, sc0 + GPB * 64};
          const unsigned char* const zp[2] = {z0, z0 + GPB * 16};
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
            const unsigned char* const wp1[1] = {w0 + Cfg::kBlockBytes};
            const unsigned char* const sp1[1] = {sc0 + GPB * 64};
            const unsigned char* const zp1[1] = {z0 + GPB * 16};
            const __half* const ap1[1] = {a0 + 128};
            const unsigned char* const ztp1[1] = {zt0 + GPB * zt_group_bytes};
            w4p_consume<UPG, 1>(wp1, sp1, zp1, ap1, ztp1, zt_group_bytes, L, tot);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&my_empty[s0]);
    }

    if (tid == 0) P_TRACE(5);
    // ---- this warp's piece [j, j + cnt) of `tile` is done: partial tile -> shared memory; whoever completes the
    // CTA's portion of the tile sums the warps' partials in slice order and stores / publishes / finishes it
    const int t0 = tile * nb;
    const int p0 = max(lo, t0), p1 = min(hi, t0 + nb);              // the CTA's portion of the tile
    const bool is_first = (j == p0);                                // this piece starts the portion
    const bool contributes = !PAIR || hh < cnt;                     // the second warp of a pair has nothing in a one-block piece
    const int par = PAIR ? (tile & 1) : 0;                          // pair partners are at most one piece apart
    auto part_of = [&](int sl, int h, int first) { return part_sm + (size_t)((((sl * 2 + h) * 2 + first) * 2 + par) * a.M) * 32; };
    if (contributes) {
      float* mine = part_of(rho, hh, is_first ? 1 : 0);
#pragma unroll
      for (int tt = 0; tt < 2; ++tt)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int m = 2 * r + (q & 1);
          const int col = 4 * c8 + 2 * tt + (q >> 1);
          if (m < a.M) mine[m * 32 + col] = tot[tt][q] * 16777216.f;
        }
    }
    if (PAIR) asm volatile("bar.sync %0, 64;" ::"r"(2 + rg) : "memory");   // both warps of the pair have written their parts
    else __syncwarp();
    // slices of the CTA that hold blocks of the portion (a slice can be empty when the CTA has fewer blocks than slices:
    // the LAST slice that begins at or before a block holds it)
    int sl_first = 0, sl_last = 0;
#pragma unroll
    for (int sl = 1; sl < kPWarps; ++sl) {
      const int b = bnd_sm[sl];
      if (b <= p0) sl_first = sl;
      if (b < p1) sl_last = sl;
    }
    // blocks of the portion held by slice sl (<= 0: none)
    auto slice_cnt = [&](int sl) { return min(bnd_sm[sl + 1], p1) - max(bnd_sm[sl], p0); };
    bool finalize = contributes && hh == 0;
    if (finalize && sl_last > sl_first) {
      int expected = 0;
      for (int sl = sl_first; sl <= sl_last; ++sl) expected += slice_cnt(sl) > 0 ? 1 : 0;
      int old = 0;
      if (lane == 0) {
        __threadfence_block();
        old = atomicAdd(&cnt_sm[sl_first], 1);
      }
      old = __shfl_sync(0xffffffffu, old, 0);
      finalize = (old == expected - 1);
      __threadfence_block();
    }
    if (finalize) {
      const bool starts_tile = (p0 == t0), ends_tile = (p1 == t0 + nb);
      const int n = tile * 32 + lane;                               // this lane's output column
      for (int m = 0; m < a.M; ++m) {
        float v = 0.f;
        for (int sl = sl_first; sl <= sl_last; ++sl) {
          const int n = slice_cnt(sl);
          if (n > 0) v += part_of(sl, 0, sl == sl_first ? 1 : 0)[m * 32 + lane];
          if (PAIR && n >= 2) v += part_of(sl, 1, sl == sl_first ? 1 : 0)[m * 32 + lane];
        }
        if (starts_tile && !ends_tile) {
          // finisher: the CTAs after this one that hold the tile's later blocks published their parts (normally long ago)
          const long long ux = (long long)(t0 + nb - 1) / a.unit;
          const int c_last = (int)(((ux + 1) * G + U - 1) / U) - 1;
          for (int cc = c + 1; cc <= c_last; ++cc) {
            unsigned long long* slot = a.ws + ((size_t)cc * a.M + m) * 32 + lane;
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
          const __half h = __float2half_rn(v);
          const size_t off = (size_t)m * a.ldo + a.col_offset + n;
          a.out[0][off] = h;
          for (int p = 1; p < a.world; ++p) a.out[p][off] = h;      // fused all-gather: NVLink peer stores
        } else {
          // contributor: this CTA holds later blocks of a tile that starts in an earlier CTA
          asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(a.ws + ((size_t)c * a.M + m) * 32 + lane), "r"(__float_as_uint(v)), "r"(1u) : "memory");
        }
      }
    }
    if (tid == 0) P_TRACE(6);
    j += cnt;
    kb = 0;
    ++tile;
  }
#ifdef XBIT_DEVTOOLS
  if (tid == 0 && a.trace) {
    P_TRACE_VALUE(8, (unsigned long long)(clock64() - loop0));
    P_TRACE_VALUE(11, (unsigned long long)(jend - (lo + (int)((long long)len * rho / kPWarps))));
    P_TRACE(7);
  }
#endif
}

// ------------------------------------------------------------------------------------------------ host side

static int p_upg_of(int groupsize) { return groupsize == 32 ? 1 : (groupsize == 64 ? 2 : 4); }

struct W4PPlan {
  int grid, unit, ring, nw, mode;
  size_t smem;
};

static size_t w4p_smem_bytes(int upg, int nr, int m, int k, int ring) {
  const int gpb = 4 / upg;
  return 1024                                                        // alignment slack
         + (size_t)nr * ring * (4096 + 160 * gpb)                    // rings of two-block slots
         + (size_t)2 * nr * kPMaxRing * 8 + 256                      // mbarriers, counters, slice boundaries
         + (size_t)nr * 8 * m * 32 * sizeof(float)                   // part_sm
         + (size_t)(k / 128) * gpb * m * 16                          // zt_sm
         + (size_t)m * (k + 8) * sizeof(__half);                     // act_sm
}

size_t gemv_w4p_workspace_bytes(int M) {
  return (size_t)device_sm_count() * (size_t)(M < 1 ? 1 : (M > 8 ? 8 : M)) * 32 * sizeof(unsigned long long);
}

static bool plan_w4p(const GemvArgs& a, bool have_ws, W4PPlan& p) {
  if (!gemv_w4_supported(a) || a.M > 8) return false;
  const int sms = device_sm_count();
  const int upg = p_upg_of(a.groupsize);
  const long long tiles = a.N / 32, nb = a.K / 128;
  if (tiles * nb > 0x3fffffffLL) return false;
  // ring: as deep as half an SM allows (so that the next call's CTA is co-resident), 2..3 two-block slots per warp; when
  // the staged activations of a large M * K leave no room for that, one CTA per SM with whatever fits
  const size_t half = 112 * 1024;
  // 16 consumer warps in pairs (MODE 2) unless overridden: XBIT_W4P_WARPS = 8 (two blocks per warp and step), 12 (a ring per warp)
  const int env_nw = env_int("XBIT_W4P_WARPS", 16);
  p.nw = env_nw == 8 ? 8 : (env_nw == 12 ? 12 : 16);
  p.mode = p.nw == 8 ? 1 : (p.nw == 12 ? 0 : 2);
  const int nw = p.mode == 2 ? p.nw / 2 : p.nw;     // rings
  int ring = 0;
  for (int r = 3; r >= 2 && !ring; --r)
    if (w4p_smem_bytes(upg, nw, a.M, a.K, r) <= half) ring = r;
  for (int r = 4; r >= 2 && !ring; --r)
    if (w4p_smem_bytes(upg, nw, a.M, a.K, r) <= kMaxDynSmem) ring = r;
  if (!ring) return false;
  const int env_ring = env_int("XBIT_W4P_RING", 0);
  if (env_ring >= 2 && env_ring <= kPMaxRing && w4p_smem_bytes(upg, nw, a.M, a.K, env_ring) <= kMaxDynSmem) ring = env_ring;
  p.ring = ring;
  p.smem = w4p_smem_bytes(upg, nw, a.M, a.K, ring);
  // CTA boundaries: block granular (perfect balance, tiles shared between CTAs meet in the workspace) or tile aligned
  // (nothing crosses CTAs).  Cost model in blocks per CTA; the cross-CTA fix-up is worth about 4 blocks of time.
  const long long total = tiles * nb;
  const long long g_fine = total < sms ? total : sms;
  // tile-aligned: always one CTA per SM, also when there are fewer tiles (CTAs without work hold their slot until the
  // previous launch has finished, so that the hardware never stacks two working CTAs of one launch on an SM: measured
  // 2.7 us against 1.7 us for the stacked ones, profiles/r02_trace_*)
  const long long g_tile = sms;
  const long long cost_fine = (total + g_fine - 1) / g_fine + (total % g_fine == 0 && (total / g_fine) % nb == 0 ? 0 : 4);
  const long long cost_tile = (tiles + sms - 1) / sms * nb;
  bool fine = have_ws && cost_fine < cost_tile;
  const int env_unit = env_int("XBIT_W4P_FINE", -1);
  if (env_unit == 0) fine = false;
  if (env_unit == 1 && have_ws) fine = true;
  p.unit = fine ? 1 : (int)nb;
  p.grid = (int)(fine ? g_fine : g_tile);
  const int env_grid = env_int("XBIT_W4P_GRID", 0);
  if (env_grid > 0 && (env_grid <= total || !fine)) p.grid = env_grid;
  return true;
}

bool gemv_w4p_applicable(const GemvArgs& a) {
  W4PPlan p;
  return plan_w4p(a, false, p);
}

using W4PKernel = void (*)(const CUtensorMap, const CUtensorMap, const W4PArgs);

cudaError_t launch_gemv_w4p(const GemvArgs& g_in, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  GemvArgs g = g_in;
  apply_debug_knobs(g);
  const bool have_ws = workspace && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0 && workspace_bytes >= gemv_w4p_workspace_bytes(g.M);
  W4PPlan p;
  if (!plan_w4p(g, have_ws, p)) return cudaErrorInvalidValue;
  const int upg = p_upg_of(g.groupsize);
  W4PArgs a;
  memset(&a, 0, sizeof(a));
  a.a = g.a;
  for (int i = 0; i < kMaxPeers; ++i) a.out[i] = g.out[i];
  a.world = g.world;
  a.ldo = g.ldo;
  a.col_offset = g.col_offset;
  a.M = g.M; a.K = g.K; a.N = g.N; a.zero_bias = g.zero_bias;
  a.nb = g.K / 128;
  a.total = (g.N / 32) * a.nb;
  a.unit = p.unit;
  a.ring = p.ring;
  a.static_weights = g.static_weights;
  a.all_wait = env_int("XBIT_W4P_ALLWAIT", 0);
  a.prefetch_delay = env_int("XBIT_W4P_DELAY", 0);
  a.scales = g.scales;
  a.qzeros = g.qzeros;
  a.zwords = g.zwords;
  a.ws = p.unit == 1 ? reinterpret_cast<unsigned long long*>(workspace) : nullptr;
  a.trace = g.trace;
  a.debug_skip = g.debug_skip;
  W4PKernel kern = nullptr;
#define XBIT_W4P_CASE(UPG_) \
  if (upg == UPG_) kern = p.nw == 8 ? gemv_w4p_kernel<UPG_, 8, 1> : (p.nw == 16 ? gemv_w4p_kernel<UPG_, 16, 2> : gemv_w4p_kernel<UPG_, 12, 0>);
  XBIT_W4P_CASE(1) XBIT_W4P_CASE(2) XBIT_W4P_CASE(4)
#undef XBIT_W4P_CASE

  alignas(64) CUtensorMap wmap2, wmap1;
  // qweight [qrows, N] u32: box = two blocks (32 word-rows) x 32 columns (128 B), 128-byte swizzle; one block for odd tails
  cudaError_t e = encode_2d(&wmap2, CU_TENSOR_MAP_DATA_TYPE_UINT32, g.qweight, (uint64_t)g.N, (uint64_t)g.qrows, (uint64_t)g.N * 4, 32, 32,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  if (e != cudaSuccess) return e;
  e = encode_2d(&wmap1, CU_TENSOR_MAP_DATA_TYPE_UINT32, g.qweight, (uint64_t)g.N, (uint64_t)g.qrows, (uint64_t)g.N * 4, 32, 16,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  if (e != cudaSuccess) return e;
  e = ensure_max_dyn_smem(reinterpret_cast<const void*>(kern));
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
  return cudaLaunchKernelEx(&cfg, kern, wmap2, wmap1, a);
}

}  // namespace xbit
