// bw_probe.cu -- access-pattern probe for streaming a [rows x N] int32 matrix (the qweight layout)
// on B200.  Not part of the product: it only answers "which global->SM path and which contiguous
// run length reach HBM speed", to size the GEMV tiles.  Patterns:
//   A  quad pattern of the first GEMV kernel: a warp load = 4 rows x 128 B (LDG.128, ring of 8)
//   B  row pattern: a warp load = 1 row x 512 B (LDG.128, ring of 8)
//   C  cp.async.bulk (TMA 1-D) row segments of NT*4 bytes into an mbarrier ring, consumers LDS.128
// Usage: bw_probe [rows N]...   prints GB/s per pattern (rotating over > 1 GiB of buffers).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// ---- A: quad pattern. CTA = 32 columns, 8 warps split the rows.
__global__ void __launch_bounds__(256) probe_a(const uint32_t* __restrict__ q, uint32_t* __restrict__ out, int rows, int N, int splits) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = lane & 3, c8 = lane >> 2;
  const int n = blockIdx.x * 32 + 4 * c8;
  const int units = rows / 4, ups = (units + splits - 1) / splits;
  const int u0s = blockIdx.y * ups, u1s = min(u0s + ups, units);
  const int wq = (u1s - u0s + 7) / 8;
  const int my0 = min(u0s + warp * wq, u1s), my1 = min(my0 + wq, u1s);
  uint4 ring[8];
#pragma unroll
  for (int p = 0; p < 8; ++p) ring[p] = (my0 + p < my1) ? ldg_stream_v4(q + (size_t)((my0 + p) * 4 + r) * N + n) : make_uint4(0, 0, 0, 0);
  uint32_t acc = 0;
  for (int u = my0; u < my1; u += 8) {
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      if (u + p < my1) {
        uint4 v = ring[p];
        ring[p] = (u + p + 8 < my1) ? ldg_stream_v4(q + (size_t)((u + p + 8) * 4 + r) * N + n) : make_uint4(0, 0, 0, 0);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
      }
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}

// ---- B: row pattern. CTA = 128 columns, 8 warps split the rows; a warp load = 1 row x 512 B.
__global__ void __launch_bounds__(256) probe_b(const uint32_t* __restrict__ q, uint32_t* __restrict__ out, int rows, int N, int splits) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 128 + 4 * lane;
  const int rps = (rows + splits - 1) / splits;
  const int r0s = blockIdx.y * rps, r1s = min(r0s + rps, rows);
  const int wq = (r1s - r0s + 7) / 8;
  const int my0 = min(r0s + warp * wq, r1s), my1 = min(my0 + wq, r1s);
  uint4 ring[8];
#pragma unroll
  for (int p = 0; p < 8; ++p) ring[p] = (my0 + p < my1) ? ldg_stream_v4(q + (size_t)(my0 + p) * N + n) : make_uint4(0, 0, 0, 0);
  uint32_t acc = 0;
  for (int u = my0; u < my1; u += 8) {
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      if (u + p < my1) {
        uint4 v = ring[p];
        ring[p] = (u + p + 8 < my1) ? ldg_stream_v4(q + (size_t)(u + p + 8) * N + n) : make_uint4(0, 0, 0, 0);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
      }
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}

// ---- C: bulk-copy ring. CTA = NT columns; stage = SR rows x NT*4 bytes; 8 consumer warps + 1 producer warp.
template <int NT, int SR, int STAGES>
__global__ void __launch_bounds__(288) probe_c(const uint32_t* __restrict__ q, uint32_t* __restrict__ out, int rows, int N, int splits) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int PITCH = NT * 4 + 32;                 // bytes per staged row (+32: bank spread for the quad LDS pattern)
  constexpr int STAGE_BYTES = SR * PITCH;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * NT;
  const int rps = ((rows + splits - 1) / splits + SR - 1) / SR * SR;
  const int r0 = blockIdx.y * rps, r1 = min(r0 + rps, rows);
  const int ntiles = (max(r1 - r0, 0) + SR - 1) / SR;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 8) {
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    for (int t = 0; t < ntiles; ++t) {
      const int s = t % STAGES;
      if (t >= STAGES) mbar_wait(&empty[s], ((t / STAGES) - 1) & 1);
      const int row0 = r0 + t * SR;
      const int nrows = min(SR, r1 - row0);
      if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)nrows * NT * 4);
      __syncwarp();
      for (int rr = lane; rr < nrows; rr += 32)
        bulk_g2s(smem + s * STAGE_BYTES + rr * PITCH, q + (size_t)(row0 + rr) * N + n0, NT * 4, &full[s], policy);
    }
  } else {
    uint32_t acc = 0;
    const int r = lane & 3, c8 = lane >> 2;
    for (int t = 0; t < ntiles; ++t) {
      const int s = t % STAGES;
      mbar_wait(&full[s], (t / STAGES) & 1);
      const unsigned char* st = smem + s * STAGE_BYTES;
      // quad pattern over the stage: warp handles (unit, 32-col chunk) pairs round-robin
      constexpr int UNITS = SR / 4, CHUNKS = NT / 32;
      for (int job = warp; job < UNITS * CHUNKS; job += 8) {
        const int unit = job / CHUNKS, ch = job % CHUNKS;
        const uint4 v = *reinterpret_cast<const uint4*>(st + (unit * 4 + r) * PITCH + (ch * 32 + 4 * c8) * 4);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (acc == 0x12345678u) out[0] = acc;
  }
}

template <typename F>
static float time_rot(F launch, int R, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < R; ++i) launch(i);
  CK(cudaDeviceSynchronize());
  std::vector<float> ts;
  for (int rep = 0; rep < reps; ++rep) {
    CK(cudaEventRecord(e0));
    for (int i = 0; i < R; ++i) launch(i);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    ts.push_back(ms * 1e3f / R);
  }
  std::sort(ts.begin(), ts.end());
  return ts[ts.size() / 2];
}

template <int NT, int SR, int STAGES>
static void run_c(const char* name, uint32_t* buf, uint32_t* out, int rows, int N, int R, size_t per, int splits) {
  const size_t smem = (size_t)STAGES * SR * (NT * 4 + 32) + 2 * STAGES * 8;
  CK(cudaFuncSetAttribute(probe_c<NT, SR, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((N + NT - 1) / NT, splits);
  float us = time_rot([&](int i) { probe_c<NT, SR, STAGES><<<grid, 288, smem>>>(buf + (size_t)(i % R) * per, out, rows, N, splits); }, R, 5);
  CK(cudaGetLastError());
  printf("   %-22s splits=%d grid=%4d  %8.2f us  %7.0f GB/s\n", name, splits, grid.x * grid.y, us, (double)rows * N * 4 / us / 1e3);
}

int main(int argc, char** argv) {
  std::vector<std::pair<int, int>> shapes;
  for (int i = 1; i + 1 < argc; i += 2) shapes.push_back({atoi(argv[i]), atoi(argv[i + 1])});
  if (shapes.empty()) shapes = {{512, 4096}, {512, 11008}, {1376, 4096}, {1024, 8192}, {1024, 28672}, {3584, 8192}};
  uint32_t* out; CK(cudaMalloc(&out, 64));
  for (auto [rows, N] : shapes) {
    const size_t per = (size_t)rows * N;
    const int R = (int)std::max<size_t>(2, ((size_t)1 << 30) / (per * 4) + 1);
    uint32_t* buf; CK(cudaMalloc(&buf, per * 4 * R));
    CK(cudaMemset(buf, 1, per * 4 * R));
    printf("== rows=%d N=%d (%.1f MB, R=%d)\n", rows, N, per * 4 / 1e6, R);
    for (int splits : {1, 2, 4, 8}) {
      dim3 ga((N + 31) / 32, splits);
      float us = time_rot([&](int i) { probe_a<<<ga, 256>>>(buf + (size_t)(i % R) * per, out, rows, N, splits); }, R, 5);
      printf("   %-22s splits=%d grid=%4d  %8.2f us  %7.0f GB/s\n", "A quad LDG 4x128B", splits, ga.x * ga.y, us, (double)per * 4 / us / 1e3);
    }
    for (int splits : {1, 2, 4, 8, 16}) {
      dim3 gb((N + 127) / 128, splits);
      float us = time_rot([&](int i) { probe_b<<<gb, 256>>>(buf + (size_t)(i % R) * per, out, rows, N, splits); }, R, 5);
      printf("   %-22s splits=%d grid=%4d  %8.2f us  %7.0f GB/s\n", "B row LDG 1x512B", splits, gb.x * gb.y, us, (double)per * 4 / us / 1e3);
    }
    for (int splits : {1, 2, 4, 8}) {
      run_c<64, 64, 4>("C bulk NT=64 SR=64 S4", buf, out, rows, N, R, per, splits);
      run_c<128, 32, 4>("C bulk NT=128 SR=32 S4", buf, out, rows, N, R, per, splits);
      run_c<256, 16, 4>("C bulk NT=256 SR=16 S4", buf, out, rows, N, R, per, splits);
      run_c<128, 32, 6>("C bulk NT=128 SR=32 S6", buf, out, rows, N, R, per, splits);
      run_c<128, 64, 3>("C bulk NT=128 SR=64 S3", buf, out, rows, N, R, per, splits);
    }
    CK(cudaFree(buf));
  }
  printf("done\n");
  return 0;
}
