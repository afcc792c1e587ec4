// gemv_sm100.cu -- fused dequant + A16Wx GEMV / skinny GEMM for sm_100a.
//
// Replaces /root/reference/src/cu/gemv_w4a16_pt.cu: gemv<T> (:35-145), warpReduceSum (:20-33) and
// lauch_Gemv_kernel (:149-173).  Semantics: y[m, n] = RN16( sum_k a[m, k] * DQ[k, n] ) where DQ is
// the dequantised weight of dq_sm100.cu; accumulation is fp32 (the reference: fp16 chains of 4).
//
// What the reference does and why it cannot approach the B200 roofline:
// CTA = 64 columns x all of K, so N = 4096 gives 64 CTAs on 148 SMs; 8 bytes of weights in flight
// per thread; one I2F per weight; M > 1 re-reads the weights M times; legacy default stream.
//
// This file (all kernels: 256 threads, 128-bit streaming loads, K split over warps, then over a
// thread-block cluster with a DSMEM reduction -- deterministic, no atomics, no workspace):
//
//   gemv_w4_kernel<kMma=false>  W4 SIMT GEMV, M <= 4.  Lane (r = lane%4, c = lane/4) streams
//       word-row 4u+r, columns 4c..4c+3 (LDG.128, 8 deep register ring = 128 B in flight per
//       thread); nibbles become exact fp16 integers with LOP3 (mask|magic) and one HSUB2 that also
//       removes the zero point; half2 FMA chains over (k, k+4) pairs, flushed to fp32 every 2
//       word-rows; per-group fp32 scale; split-K over the 4 r-lanes by warp shuffle, over warps by
//       shared memory, over the cluster by DSMEM.
//   gemv_w4_kernel<kMma=true>   W4 skinny GEMM, M <= 16.  Same streaming geometry; the unpacked
//       half2 pairs ARE the m16n8k16 A fragments (weights on the MMA M axis, batch on the MMA N
//       axis; the K permutation inside a fragment is absorbed by the activation layout in shared
//       memory), so no shuffle or shared-memory round trip for weights; fp32 accumulators.
//       Weights are read once for all M rows (the reference re-reads them M times, :158).
//   gemv_generic_kernel         any bits 2..8, any groupsize >= 16, any M, any N: one column per
//       thread, bit-reader over the LSB-first stream, fp32 math with the zero point folded per
//       group.  Correctness path for the combinations the reference aborts on (:152-155).
//
// Programmatic dependent launch: weights do not depend on the previous kernel in a decode step, so
// the weight ring is filled BEFORE griddepcontrol.wait and only the activation staging waits.
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "unpack.cuh"
#include "xbit_internal.h"
#include "../../include/xbitops_b200.h"

namespace cg = cooperative_groups;

namespace xbit {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// D += A(16x16, row) * B(16x8, col), fp16 inputs, fp32 accumulate
__device__ __forceinline__ void mma_m16n8k16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                             uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

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

// activations [8 consecutive k] (pairs (0,1)(2,3)(4,5)(6,7)) -> pairs (0,4)(1,5)(2,6)(3,7)
__device__ __forceinline__ uint4 permute_act8(uint4 v) {
  uint4 o;
  o.x = prmt(v.x, v.z, 0x5410);
  o.y = prmt(v.x, v.z, 0x7632);
  o.z = prmt(v.y, v.w, 0x5410);
  o.w = prmt(v.y, v.w, 0x7632);
  return o;
}

// kMma = false: MV = number of activation rows (1..4).  kMma = true: MV = number of n8 batch tiles
// (1: M <= 8, 2: M <= 16).  CT = 32-column chunks per CTA.
template <bool kMma, int MV, int CT>
__global__ void __launch_bounds__(kThreads)
gemv_w4_kernel(const GemvArgs a) {
  constexpr int NT = 32 * CT;                      // columns per CTA
  constexpr int PF = (CT == 1) ? 8 : 4;            // ring depth: 128 B of weights in flight per thread
  constexpr int MROWS = kMma ? 8 * MV : MV;        // activation rows this instantiation can take
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = lane & 3, c8 = lane >> 2;
  const int n_cta = blockIdx.x * NT;
  const int split = blockIdx.y;
  const int total_units = (a.K + 31) >> 5;
  const int u_begin = min(split * a.units_per_split, total_units);
  const int u_end = min(u_begin + a.units_per_split, total_units);
  const int upg = a.groupsize >> 5;                // 32-k units per group
  const int pitch = a.chunk_units * 32 + 32;       // halves per staged activation row (+64 B: bank spread)
  __half* act_sm = reinterpret_cast<__half*>(smem_raw);
  float* red_sm = reinterpret_cast<float*>(smem_raw + (size_t)MROWS * pitch * sizeof(__half));
  float* clus_sm = red_sm + kWarps * MROWS * NT;   // [splits][MROWS][NT], only the cluster leader's is used

  const bool clustered = a.splits > 1;
  if (clustered) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");   // "I have started"
  // Default: nothing is read before the previous kernel in the stream has completed.  With
  // XBIT_GEMV_FLAG_STATIC_WEIGHTS the caller promises the weights were not produced by that
  // kernel, and only the activation staging below waits.
  if (!a.static_weights) griddep_wait();

  int ncol[CT];
  bool cvalid[CT];
#pragma unroll
  for (int ct = 0; ct < CT; ++ct) {
    ncol[ct] = n_cta + 32 * ct + 4 * c8;
    cvalid[ct] = ncol[ct] < a.N;                   // N % 8 == 0: a lane's 4 columns are all in or all out
  }

  auto load_w = [&](int unit, int ct) -> uint4 {
    const int row = unit * 4 + r;
    if (cvalid[ct] && row < a.qrows) return ldg_stream_v4(a.qweight + (size_t)row * a.N + ncol[ct]);
    return make_uint4(0, 0, 0, 0);
  };

  // ---- accumulators
  //  SIMT: tot[ct][m][j]              column 4*c8+j, row m, partial over this lane's word-rows
  //  MMA : tot[ct][t][mt][i]  as [ct][t*MV+mt][i]   t = column pair, i = mma accumulator index
  constexpr int AV = kMma ? 2 * MV : MV;
  float tot[CT][AV][4];
  float grp[CT][AV][4];
#pragma unroll
  for (int ct = 0; ct < CT; ++ct)
#pragma unroll
    for (int v = 0; v < AV; ++v)
#pragma unroll
      for (int i = 0; i < 4; ++i) { tot[ct][v][i] = 0.f; grp[ct][v][i] = 0.f; }

  bool first_chunk = true;
  for (int cu0 = u_begin; cu0 < u_end || first_chunk; cu0 += a.chunk_units) {
    const int cu1 = min(cu0 + a.chunk_units, u_end);
    const int wq = (max(cu1 - cu0, 0) + kWarps - 1) / kWarps;
    const int my0 = min(cu0 + warp * wq, cu1);
    const int my1 = min(my0 + wq, cu1);

    // ---- fill the weight ring (independent of the previous kernel)
    uint4 ring[PF][CT];
#pragma unroll
    for (int p = 0; p < PF; ++p)
#pragma unroll
      for (int ct = 0; ct < CT; ++ct) ring[p][ct] = (my0 + p < my1) ? load_w(my0 + p, ct) : make_uint4(0, 0, 0, 0);

    // ---- group state for this warp's first unit
    int cur_g = (my0 < my1) ? my0 / upg : 0;
    int next_switch = (cur_g + 1) * upg;
    uint2 raw_s[CT];
    uint32_t raw_z[CT];
    auto fetch_group_raw = [&](int g) {
#pragma unroll
      for (int ct = 0; ct < CT; ++ct) {
        raw_s[ct] = make_uint2(0, 0);
        raw_z[ct] = 0;
        if (cvalid[ct] && g < a.groups) {
          raw_s[ct] = __ldg(reinterpret_cast<const uint2*>(a.scales + (size_t)g * a.N + ncol[ct]));
          raw_z[ct] = __ldg(a.qzeros + (size_t)g * a.zwords + (ncol[ct] >> 3)) >> (16 * (c8 & 1));
        }
      }
    };
    float sf[CT][4];
    uint32_t zc_lo[CT][4], zc_hi[CT][4];
    auto decode_group = [&]() {
#pragma unroll
      for (int ct = 0; ct < CT; ++ct) {
        const float2 s01 = __half22float2(u2h2(raw_s[ct].x));
        const float2 s23 = __half22float2(u2h2(raw_s[ct].y));
        sf[ct][0] = s01.x; sf[ct][1] = s01.y; sf[ct][2] = s23.x; sf[ct][3] = s23.y;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t zb = ((raw_z[ct] >> (4 * j)) & 0xFu) + (uint32_t)a.zero_bias;   // <= 16
          zc_lo[ct][j] = dup16(magic_base_bits(0) + zb);          // half(1024 + z)
          zc_hi[ct][j] = dup16(magic_base_bits(4) + (zb << 4));   // half(64 + z)
        }
      }
    };
    if (my0 < my1) {
      fetch_group_raw(cur_g);
      decode_group();
      fetch_group_raw(cur_g + 1);     // prefetch the next group's scale/zero words
    }

    // ---- stage this chunk's activations (the only data that depends on the previous kernel)
    if (first_chunk && a.static_weights) griddep_wait();
    {
      const int chunk_k0 = cu0 * 32;
      const int vecs_per_row = max(cu1 - cu0, 0) * 4;      // 8-half vectors
      for (int idx = tid; idx < MROWS * vecs_per_row; idx += kThreads) {
        const int m = idx / vecs_per_row, v = idx - m * vecs_per_row;
        const int k = chunk_k0 + v * 8;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (m < a.M && k < a.K) val = __ldg(reinterpret_cast<const uint4*>(a.a + (size_t)m * a.K + k));  // K % 8 == 0
        *reinterpret_cast<uint4*>(act_sm + (size_t)m * pitch + v * 8) = permute_act8(val);
      }
    }
    __syncthreads();

    // SIMT fp16 chains (flushed to fp32 every 2 units)
    __half2 acc16[CT][kMma ? 1 : MV][4];
    if constexpr (!kMma) {
#pragma unroll
      for (int ct = 0; ct < CT; ++ct)
#pragma unroll
        for (int m = 0; m < MV; ++m)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc16[ct][m][j] = __float2half2_rn(0.f);
    }
    auto flush16 = [&]() {
      if constexpr (!kMma) {
#pragma unroll
        for (int ct = 0; ct < CT; ++ct)
#pragma unroll
          for (int m = 0; m < MV; ++m)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __half22float2(acc16[ct][m][j]);
              grp[ct][m][j] += f.x + f.y;
              acc16[ct][m][j] = __float2half2_rn(0.f);
            }
      }
    };
    auto close_group = [&]() {       // tot += scale * group sum
      flush16();
#pragma unroll
      for (int ct = 0; ct < CT; ++ct)
#pragma unroll
        for (int v = 0; v < AV; ++v)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            // SIMT: i = column j.  MMA: v = t*MV+mt, accumulators 0,1 -> column 2t, 2,3 -> column 2t+1
            const float s = kMma ? sf[ct][2 * (v / MV) + (i >> 1)] : sf[ct][i];
            tot[ct][v][i] = fmaf(s, grp[ct][v][i], tot[ct][v][i]);
            grp[ct][v][i] = 0.f;
          }
    };

    for (int u = my0; u < my1; u += PF) {
#pragma unroll
      for (int p = 0; p < PF; ++p) {
        const int uu = u + p;
        if (uu < my1) {                                     // warp-uniform
          uint4 wv[CT];
#pragma unroll
          for (int ct = 0; ct < CT; ++ct) {
            wv[ct] = ring[p][ct];
            ring[p][ct] = (uu + PF < my1) ? load_w(uu + PF, ct) : make_uint4(0, 0, 0, 0);
          }
          if (uu == next_switch) {                          // warp-uniform: entering the next group
            close_group();
            ++cur_g;
            next_switch += upg;
            decode_group();
            fetch_group_raw(cur_g + 1);
          }
          const __half* arow = act_sm + (uu - cu0) * 32 + r * 8;
          if constexpr (kMma) {
            uint4 bfrag[MV];
#pragma unroll
            for (int mt = 0; mt < MV; ++mt) {
              bfrag[mt] = make_uint4(0, 0, 0, 0);
              const int m = c8 + 8 * mt;
              if (m < a.M) bfrag[mt] = *reinterpret_cast<const uint4*>(arow + (size_t)m * pitch);
            }
#pragma unroll
            for (int ct = 0; ct < CT; ++ct) {
              const uint32_t w4[4] = {wv[ct].x, wv[ct].y, wv[ct].z, wv[ct].w};
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                uint32_t ea[4], eb[4];
                unpack_w4_minus_zero(w4[2 * t], zc_lo[ct][2 * t], zc_hi[ct][2 * t], ea);
                unpack_w4_minus_zero(w4[2 * t + 1], zc_lo[ct][2 * t + 1], zc_hi[ct][2 * t + 1], eb);
#pragma unroll
                for (int mt = 0; mt < MV; ++mt) {
                  mma_m16n8k16(grp[ct][t * MV + mt], ea[0], eb[0], ea[1], eb[1], bfrag[mt].x, bfrag[mt].y);
                  mma_m16n8k16(grp[ct][t * MV + mt], ea[2], eb[2], ea[3], eb[3], bfrag[mt].z, bfrag[mt].w);
                }
              }
            }
          } else {
            uint4 av[MV];
#pragma unroll
            for (int m = 0; m < MV; ++m) av[m] = *reinterpret_cast<const uint4*>(arow + (size_t)m * pitch);
#pragma unroll
            for (int ct = 0; ct < CT; ++ct) {
              const uint32_t w4[4] = {wv[ct].x, wv[ct].y, wv[ct].z, wv[ct].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint32_t e[4];
                unpack_w4_minus_zero(w4[j], zc_lo[ct][j], zc_hi[ct][j], e);
#pragma unroll
                for (int m = 0; m < MV; ++m) {
                  __half2 c = acc16[ct][m][j];
                  c = __hfma2(u2h2(e[0]), u2h2(av[m].x), c);
                  c = __hfma2(u2h2(e[1]), u2h2(av[m].y), c);
                  c = __hfma2(u2h2(e[2]), u2h2(av[m].z), c);
                  c = __hfma2(u2h2(e[3]), u2h2(av[m].w), c);
                  acc16[ct][m][j] = c;
                }
              }
            }
            if (p & 1) flush16();
          }
        }
      }
    }
    if (my0 < my1) close_group();
    first_chunk = false;
    if (cu0 + a.chunk_units < u_end) __syncthreads();      // the next chunk overwrites act_sm
  }

  // Let the next kernel in the stream start its prologue / weight prefetch.
  griddep_launch_dependents();

  // ---- split-K reduction: r-lanes (shuffle) -> warps (smem) -> cluster (DSMEM)
  if constexpr (!kMma) {
#pragma unroll
    for (int ct = 0; ct < CT; ++ct)
#pragma unroll
      for (int m = 0; m < MV; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v = tot[ct][m][j];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          if (r == 0) red_sm[(warp * MROWS + m) * NT + 32 * ct + 4 * c8 + j] = v;
        }
  } else {
#pragma unroll
    for (int ct = 0; ct < CT; ++ct)
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int mt = 0; mt < MV; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int m = 8 * mt + 2 * r + (i & 1);
            const int col = 32 * ct + 4 * c8 + 2 * t + (i >> 1);
            red_sm[(warp * MROWS + m) * NT + col] = tot[ct][t * MV + mt][i];
          }
  }
  __syncthreads();

  const int nout = a.M * NT;                                // M <= MROWS
  if (clustered) {
    cg::cluster_group cluster = cg::this_cluster();
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // every CTA of the cluster has started
    float* leader = cluster.map_shared_rank(clus_sm, 0);
    for (int o = tid; o < nout; o += kThreads) {
      const int m = o / NT, col = o - m * NT;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) v += red_sm[(w * MROWS + m) * NT + col];
      leader[(split * MROWS + m) * NT + col] = v;
    }
    cluster.sync();
    if (split != 0) return;
  }
  for (int o = tid; o < nout; o += kThreads) {
    const int m = o / NT, col = o - m * NT;
    const int n = n_cta + col;
    float v = 0.f;
    if (clustered) {
      for (int s = 0; s < a.splits; ++s) v += clus_sm[(s * MROWS + m) * NT + col];
    } else {
#pragma unroll
      for (int w = 0; w < kWarps; ++w) v += red_sm[(w * MROWS + m) * NT + col];
    }
    if (n < a.N) {
      const __half h = __float2half_rn(v);
      const size_t off = (size_t)m * a.ldo + a.col_offset + n;
      a.out[0][off] = h;
      for (int p = 1; p < a.world; ++p) a.out[p][off] = h;   // fused all-gather: NVLink peer stores
    }
  }
}

// ------------------------------------------------------------------------------------------------
// generic: any bits / groupsize / M / N.  One column per thread, 32 columns x 8 K-slices per CTA.
__global__ void __launch_bounds__(kThreads)
gemv_generic_kernel(const GemvArgs a, int m_base) {
  __shared__ float red[kWarps][4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const bool valid = n < a.N;
  const int bits = a.bits;
  const uint32_t mask = (1u << bits) - 1u;
  const int total_units = (a.K + 31) >> 5;
  const int wq = (total_units + kWarps - 1) / kWarps;
  const int my0 = min(warp * wq, total_units), my1 = min(my0 + wq, total_units);
  const int mcount = min(4, a.M - m_base);

  float tot[4] = {0.f, 0.f, 0.f, 0.f};
  float dsum[4] = {0.f, 0.f, 0.f, 0.f};   // sum a_k * w_k inside the current group
  float asum[4] = {0.f, 0.f, 0.f, 0.f};   // sum a_k inside the current group
  int cur_g = -1;
  float s = 0.f, z = 0.f;
  griddep_wait();
  if (valid) {
    for (int u = my0; u < my1; ++u) {
      unsigned long long buf = 0;
      int avail = 0, next_word = u * bits;
      for (int i = 0; i < 32; ++i) {
        const int k = u * 32 + i;
        if (k >= a.K) break;
        if (avail < bits) {
          const uint32_t wd = next_word < a.qrows ? __ldg(a.qweight + (size_t)next_word * a.N + n) : 0u;
          buf |= (unsigned long long)wd << avail;
          avail += 32;
          ++next_word;
        }
        const float wv = (float)(uint32_t)(buf & mask);
        buf >>= bits;
        avail -= bits;
        const int g = k / a.groupsize;
        if (g != cur_g) {
          if (cur_g >= 0) {
#pragma unroll
            for (int m = 0; m < 4; ++m) { tot[m] = fmaf(s, dsum[m] - z * asum[m], tot[m]); dsum[m] = 0.f; asum[m] = 0.f; }
          }
          cur_g = g;
          s = __half2float(a.scales[(size_t)g * a.N + n]);
          const int zpos = n * bits, zi = zpos >> 5, zsh = zpos & 31;
          const uint32_t zlo = __ldg(a.qzeros + (size_t)g * a.zwords + zi);
          const uint32_t zhi = (zsh + bits > 32 && zi + 1 < a.zwords) ? __ldg(a.qzeros + (size_t)g * a.zwords + zi + 1) : 0u;
          z = (float)((__funnelshift_r(zlo, zhi, zsh) & mask) + (uint32_t)a.zero_bias);
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          if (m < mcount) {
            const float av = __half2float(a.a[(size_t)(m_base + m) * a.K + k]);
            dsum[m] = fmaf(av, wv, dsum[m]);
            asum[m] += av;
          }
        }
      }
    }
    if (cur_g >= 0) {
#pragma unroll
      for (int m = 0; m < 4; ++m) tot[m] = fmaf(s, dsum[m] - z * asum[m], tot[m]);
    }
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) red[warp][m][lane] = tot[m];
  __syncthreads();
  if (warp < mcount && valid) {
    const int m = warp;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) v += red[w][m][lane];
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
                       reinterpret_cast<uintptr_t>(a.scales);
  return a.bits == 4 && a.groupsize % 32 == 0 && a.K % 8 == 0 && a.N % 8 == 0 && (al & 15u) == 0 && a.M >= 1;
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// Decomposition: CT (tile = 32*CT columns), K splits (cluster size) and activation chunking.
static void plan_w4(GemvArgs& a, int mrows, int& ct, size_t& smem, dim3& grid) {
  const int sms = device_sm_count();
  const int total_units = (a.K + 31) / 32;
  // tuning knobs for the sweep harness (tools/sweep): XBIT_GEMV_CT / XBIT_GEMV_SPLITS
  ct = env_int("XBIT_GEMV_CT", 0);
  if (ct != 1 && ct != 2) ct = (a.N >= 16384) ? 2 : 1;
  const int tiles = (a.N + 32 * ct - 1) / (32 * ct);
  int splits = env_int("XBIT_GEMV_SPLITS", 0);
  if (splits < 1 || splits > 8 || (splits & (splits - 1))) {
    splits = 1;
    // enough CTAs for ~2 per SM, while every warp keeps at least 4 units (128 k) of work
    while (splits < 8 && tiles * splits < 2 * sms && total_units / (splits * 2 * kWarps) >= 4) splits *= 2;
  }
  a.splits = splits;
  a.units_per_split = (total_units + splits - 1) / splits;
  // activation staging: at most 32 Ki halves (64 KiB); whole split slab when it fits
  int chunk = a.units_per_split;
  const int max_units = (32768 / mrows) / 32;
  if (chunk > max_units) chunk = max_units / kWarps * kWarps;
  a.chunk_units = chunk < 1 ? 1 : chunk;
  const int nt = 32 * ct;
  smem = (size_t)mrows * (a.chunk_units * 32 + 32) * sizeof(__half)       // act_sm
         + (size_t)kWarps * mrows * nt * sizeof(float)                     // red_sm
         + (size_t)(splits > 1 ? splits : 0) * mrows * nt * sizeof(float); // clus_sm
  grid = dim3((unsigned)tiles, (unsigned)splits, 1);
}

constexpr size_t kMaxDynSmem = 200 * 1024;

template <typename Kern>
static cudaError_t launch_w4(Kern kern, const GemvArgs& a, dim3 grid, size_t smem, cudaStream_t stream) {
  if (smem > kMaxDynSmem) return cudaErrorInvalidValue;
  // opt in to large dynamic shared memory once per (kernel, device)
  static bool configured[64] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
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
  return cudaLaunchKernelEx(&cfg, kern, a);
}

cudaError_t launch_gemv_w4_simt(GemvArgs a, cudaStream_t stream) {
  int ct; size_t smem; dim3 grid;
  plan_w4(a, a.M, ct, smem, grid);
#define XBIT_SIMT_CASE(MV_, CT_) \
  if (a.M == MV_ && ct == CT_) return launch_w4(gemv_w4_kernel<false, MV_, CT_>, a, grid, smem, stream);
  XBIT_SIMT_CASE(1, 1) XBIT_SIMT_CASE(1, 2)
  XBIT_SIMT_CASE(2, 1) XBIT_SIMT_CASE(2, 2)
  XBIT_SIMT_CASE(3, 1) XBIT_SIMT_CASE(3, 2)
  XBIT_SIMT_CASE(4, 1) XBIT_SIMT_CASE(4, 2)
#undef XBIT_SIMT_CASE
  return cudaErrorInvalidValue;
}

cudaError_t launch_gemv_w4_mma(GemvArgs a, cudaStream_t stream) {
  const int mv = a.M <= 8 ? 1 : 2;
  int ct; size_t smem; dim3 grid;
  plan_w4(a, 8 * mv, ct, smem, grid);
#define XBIT_MMA_CASE(MV_, CT_) \
  if (mv == MV_ && ct == CT_) return launch_w4(gemv_w4_kernel<true, MV_, CT_>, a, grid, smem, stream);
  XBIT_MMA_CASE(1, 1) XBIT_MMA_CASE(1, 2)
  XBIT_MMA_CASE(2, 1) XBIT_MMA_CASE(2, 2)
#undef XBIT_MMA_CASE
  return cudaErrorInvalidValue;
}

cudaError_t launch_gemv_generic(GemvArgs a, cudaStream_t stream) {
  const unsigned grid = (unsigned)((a.N + 31) / 32);
  for (int m0 = 0; m0 < a.M; m0 += 4) {
    gemv_generic_kernel<<<grid, kThreads, 0, stream>>>(a, m0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace xbit
