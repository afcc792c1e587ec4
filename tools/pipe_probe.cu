// Developer microbenchmark: issue rate of the instructions the W4 consumer is built from, on one SM
// sub-partition (SMSP).  Prints cycles per warp-instruction per SMSP for W warps/SM and C independent
// dependency chains per warp.      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int kIters = 512;

enum Op { HMMA_F32, HMMA_F16, HMMA_K8, IMMA_S8, HFMA2, FHFMA, FFMA, LOP3, SHF, PRMT, IMADHI, IMADLO, IADD3, LDS32, LDS128, NOPS };
static const char* kNames[NOPS] = {"HMMA.16816.F32", "HMMA.16816.F16", "HMMA.1688.F32", "IMMA.16832.S8", "HFMA2", "FHFMA (fma.rn.f32.f16)",
                                   "FFMA", "LOP3", "SHF", "PRMT", "IMAD.HI", "IMAD.LO", "IADD3", "LDS.32", "LDS.128"};

template <int OP, int C>
__global__ void probe(unsigned long long* cycles, uint32_t* sink, uint32_t seed) {
  __shared__ uint4 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_uint4(i, seed, i ^ seed, 7);
  __syncthreads();
  uint32_t x[C][4];
  float f[C][4];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[c][i] = seed * (c + 1) + i + threadIdx.x; f[c][i] = (float)(c + i) * 1e-3f; }
  const uint32_t a0 = seed | 0x3c003c00u, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = seed ^ 0x38003800u, b1 = b0 + 5;
  uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16;
  __syncthreads();
  const unsigned long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if constexpr (OP == HMMA_F32) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(f[c][0]), "+f"(f[c][1]), "+f"(f[c][2]), "+f"(f[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      } else if constexpr (OP == HMMA_F16) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
                     : "+r"(x[c][0]), "+r"(x[c][1]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      } else if constexpr (OP == HMMA_K8) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(f[c][0]), "+f"(f[c][1]), "+f"(f[c][2]), "+f"(f[c][3]) : "r"(a0), "r"(a1), "r"(b0));
      } else if constexpr (OP == IMMA_S8) {
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(x[c][0]), "+r"(x[c][1]), "+r"(x[c][2]), "+r"(x[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      } else if constexpr (OP == HFMA2) {
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x[c][0]) : "r"(a0), "r"(b0));
      } else if constexpr (OP == FHFMA) {
        asm volatile("{\n .reg .f16 l, h, p, q;\n mov.b32 {l, h}, %1;\n mov.b32 {p, q}, %2;\n fma.rn.f32.f16 %0, l, p, %0;\n}" : "+f"(f[c][0]) : "r"(a0), "r"(b0));
      } else if constexpr (OP == FFMA) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c][0]) : "f"(f[c][1]), "f"(f[c][2]));
      } else if constexpr (OP == LOP3) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x[c][0]) : "r"(a0), "r"(b0));
      } else if constexpr (OP == SHF) {
        asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[c][0]) : "r"(a0));
      } else if constexpr (OP == PRMT) {
        asm volatile("prmt.b32 %0, %0, %1, 0x3715;" : "+r"(x[c][0]) : "r"(a0));
      } else if constexpr (OP == IMADHI) {
        asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[c][0]) : "r"(a0));
      } else if constexpr (OP == IMADLO) {
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c][0]) : "r"(a0), "r"(b0));
      } else if constexpr (OP == IADD3) {
        asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c][0]) : "r"(a0));
      } else if constexpr (OP == LDS32) {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x[c][0]) : "r"(saddr + ((x[c][0] & 1) << 9) + c * 4));
      } else if constexpr (OP == LDS128) {
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x[c][0]), "=r"(x[c][1]), "=r"(x[c][2]), "=r"(x[c][3]) : "r"(saddr + ((x[c][0] & 1) << 9) + c * 512));
      }
    }
  }
  const unsigned long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc += x[c][i] + __float_as_uint(f[c][i]);
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP, int C>
static void run(int warps, unsigned long long* d_cycles, uint32_t* d_sink, int sms) {
  probe<OP, C><<<sms, warps * 32>>>(d_cycles, d_sink, 3);
  probe<OP, C><<<sms, warps * 32>>>(d_cycles, d_sink, 3);
  cudaDeviceSynchronize();
  unsigned long long h[256];
  cudaMemcpy(h, d_cycles, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; ++i) avg += (double)h[i];
  avg /= sms;
  const double per_smsp = (double)kIters * C * warps / 4.0;   // warp-instructions per SMSP
  printf("  W=%2d C=%d: %6.2f clk/inst/SMSP", warps, C, avg / per_smsp);
}

template <int OP>
static void sweep(unsigned long long* d_cycles, uint32_t* d_sink, int sms) {
  printf("%-24s\n", kNames[OP]);
  const int ws[3] = {4, 8, 16};
  for (int wi = 0; wi < 3; ++wi) {
    run<OP, 1>(ws[wi], d_cycles, d_sink, sms);
    run<OP, 2>(ws[wi], d_cycles, d_sink, sms);
    run<OP, 4>(ws[wi], d_cycles, d_sink, sms);
    run<OP, 8>(ws[wi], d_cycles, d_sink, sms);
    printf("\n");
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* d_cycles;
  uint32_t* d_sink;
  cudaMalloc(&d_cycles, 256 * sizeof(unsigned long long));
  cudaMalloc(&d_sink, 4096);
  sweep<HMMA_F32>(d_cycles, d_sink, sms);
  sweep<HMMA_F16>(d_cycles, d_sink, sms);
  sweep<HMMA_K8>(d_cycles, d_sink, sms);
  sweep<IMMA_S8>(d_cycles, d_sink, sms);
  sweep<HFMA2>(d_cycles, d_sink, sms);
  sweep<FHFMA>(d_cycles, d_sink, sms);
  sweep<FFMA>(d_cycles, d_sink, sms);
  sweep<LOP3>(d_cycles, d_sink, sms);
  sweep<SHF>(d_cycles, d_sink, sms);
  sweep<PRMT>(d_cycles, d_sink, sms);
  sweep<IMADHI>(d_cycles, d_sink, sms);
  sweep<IMADLO>(d_cycles, d_sink, sms);
  sweep<IADD3>(d_cycles, d_sink, sms);
  sweep<LDS32>(d_cycles, d_sink, sms);
  sweep<LDS128>(d_cycles, d_sink, sms);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
