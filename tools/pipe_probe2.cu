// Developer microbenchmark 2: do the ALU, FMA and tensor pipes of an SM sub-partition overlap?  Mixed bodies.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 512;

// MODE: 0 = 8 LOP3(imm); 1 = 8 FFMA; 2 = 4 LOP3 + 4 FFMA interleaved; 3 = 8 LOP3 + 8 FFMA interleaved (cost per pair)
//       4 = 8 LOP3 + 2 HMMA; 5 = 2 HMMA; 6 = 8 LOP3 + 2 IMAD.HI; 7 = 4 LOP3 + 4 IMAD(lo) ; 8 = LOP3 reg-reg-reg x8; 9 = 8 LOP3 + 8 FFMA + 2 HMMA
template <int MODE>
__global__ void probe(unsigned long long* cycles, uint32_t* sink, uint32_t seed) {
  uint32_t x[8];
  float f[8], acc[2][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = seed * (i + 1) + threadIdx.x; f[i] = (float)i * 1e-3f + seed; }
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;
  const uint32_t a0 = seed | 0x3c003c00u, b0 = seed ^ 0x38003800u;
  const float g0 = 1.0001f, g1 = 0.5f;
  __syncthreads();
  const unsigned long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || MODE == 3 || MODE == 4 || MODE == 6 || MODE == 9 || ((MODE == 2 || MODE == 7) && (i & 1) == 0))
        asm volatile("lop3.b32 %0, %0, 0x0f0f3355, %1, 0x6a;" : "+r"(x[i]) : "r"(b0));
      if (MODE == 8) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x[i]) : "r"(a0), "r"(b0));
      if (MODE == 1 || MODE == 3 || MODE == 9 || (MODE == 2 && (i & 1) == 1))
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(g0), "f"(g1));
      if (MODE == 7 && (i & 1) == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a0), "r"(b0));
      if (MODE == 6 && i < 2) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a0));
      if ((MODE == 4 || MODE == 5 || MODE == 9) && i < 2)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3]) : "r"(a0), "r"(a0 + 1), "r"(a0 + 2), "r"(a0 + 3), "r"(b0), "r"(b0 + 5));
    }
  }
  const unsigned long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + __float_as_uint(f[i]);
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) s += __float_as_uint(acc[c][i]);
  if (s == 0x12345678u) sink[threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, unsigned long long* d_cycles, uint32_t* d_sink, int sms) {
  printf("%-40s", name);
  const int ws[3] = {4, 8, 16};
  for (int wi = 0; wi < 3; ++wi) {
    probe<MODE><<<sms, ws[wi] * 32>>>(d_cycles, d_sink, 3);
    probe<MODE><<<sms, ws[wi] * 32>>>(d_cycles, d_sink, 3);
    cudaDeviceSynchronize();
    unsigned long long h[256];
    cudaMemcpy(h, d_cycles, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[i];
    avg /= sms;
    printf("  W=%2d: %7.2f clk/body/SMSP", ws[wi], avg / ((double)kIters * ws[wi] / 4.0));
  }
  printf("\n");
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* d_cycles;
  uint32_t* d_sink;
  cudaMalloc(&d_cycles, 256 * sizeof(unsigned long long));
  cudaMalloc(&d_sink, 4096);
  run<0>("8 LOP3 (imm mask)", d_cycles, d_sink, sms);
  run<8>("8 LOP3 (3 registers)", d_cycles, d_sink, sms);
  run<1>("8 FFMA", d_cycles, d_sink, sms);
  run<2>("4 LOP3 + 4 FFMA", d_cycles, d_sink, sms);
  run<3>("8 LOP3 + 8 FFMA", d_cycles, d_sink, sms);
  run<7>("4 LOP3 + 4 IMAD", d_cycles, d_sink, sms);
  run<5>("2 HMMA", d_cycles, d_sink, sms);
  run<4>("8 LOP3 + 2 HMMA", d_cycles, d_sink, sms);
  run<6>("8 LOP3 + 2 IMAD.HI", d_cycles, d_sink, sms);
  run<9>("8 LOP3 + 8 FFMA + 2 HMMA", d_cycles, d_sink, sms);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
