// TEST INFRASTRUCTURE ONLY.  Thin extern "C" driver around the UNMODIFIED reference CPU
// simulator: it #includes /root/reference/src/cpp_simulate.cc where it lies (path given on the
// command line as -DXBIT_REF_SIMULATE=...; nothing from the reference is copied into this repo)
// and exposes the usable pieces:
//   cpu::DequantizeAndUnpackWeight3567_v2<ushort, B>   (cpp_simulate.cc:568-691)  B = 2..8
//   cpu::cpu_gemv<ushort>                              (cpp_simulate.cc:88-221)   N = 11008 only
// MATRIX_K / MATRIX_N drive the simulator's loop bounds (cpp_simulate.cc:570) and are plain
// globals, so the DQ entry sets them per call; cpu::block_k (cpp_simulate.cc:12) is frozen at
// static-init time from the constant-initialised MATRIX_K below, which only cpu_gemv uses.
// Output: oracle/_ref/libxbit_refcpu.so (git-ignored).
#include <cstdint>
#ifndef XBIT_REF_K
#define XBIT_REF_K 4096
#endif
#ifndef XBIT_REF_N
#define XBIT_REF_N 11008
#endif
int MATRIX_M = 1;
int MATRIX_K = XBIT_REF_K;
int MATRIX_N = XBIT_REF_N;

#include XBIT_REF_SIMULATE

extern "C" {

// returns 0 on success, -1 for an unsupported bit width. NOTE: not re-entrant (globals).
__attribute__((visibility("default")))
int refcpu_dequant(uint16_t* out, const int32_t* qweight, const uint16_t* scales, const int32_t* qzeros,
                   int K, int N, int bits, int groupsize) {
  MATRIX_K = K;
  MATRIX_N = N;
  const uint32_t* qw = reinterpret_cast<const uint32_t*>(qweight);
  const uint32_t* qz = reinterpret_cast<const uint32_t*>(qzeros);
  switch (bits) {
    case 2: cpu::DequantizeAndUnpackWeight3567_v2<cpu::ushort, 2>(out, qw, scales, qz, groupsize, K, N); break;
    case 3: cpu::DequantizeAndUnpackWeight3567_v2<cpu::ushort, 3>(out, qw, scales, qz, groupsize, K, N); break;
    case 4: cpu::DequantizeAndUnpackWeight3567_v2<cpu::ushort, 4>(out, qw, scales, qz, groupsize, K, N); break;
    case 5: cpu::DequantizeAndUnpackWeight3567_v2<cpu::ushort, 5>(out, qw, scales, qz, groupsize, K, N); break;
    case 6: cpu::DequantizeAndUnpackWeight3567_v2<cpu::ushort, 6>(out, qw, scales, qz, groupsize, K, N); break;
    case 7: cpu::DequantizeAndUnpackWeight3567_v2<cpu::ushort, 7>(out, qw, scales, qz, groupsize, K, N); break;
    case 8: cpu::DequantizeAndUnpackWeight3567_v2<cpu::ushort, 8>(out, qw, scales, qz, groupsize, K, N); break;
    default: return -1;
  }
  return 0;
}

// The shipped CPU gemv, as is: valid only for K == XBIT_REF_K (block_k) and N == 11008
// (hard-coded loop bound, cpp_simulate.cc:90). Its output is the FIRST K-slab's partial sum only
//  -- it is timed for the record, never used as a numerical oracle.
// w_unpack must be the matching dequantised weights (the function aborts its loop otherwise,
// cpp_simulate.cc:186-189).
__attribute__((visibility("default")))
int refcpu_gemv_as_shipped(uint16_t* out, uint16_t* w_unpack, uint16_t* a, int32_t* qweight,
                           uint16_t* scales, int32_t* qzeros, int K, int N, int groupsize) {
  if (K != XBIT_REF_K || N != 11008) return -1;
  MATRIX_K = K;
  MATRIX_N = N;
  cpu::cpu_gemv<cpu::ushort>(out, w_unpack, a, reinterpret_cast<uint32_t*>(qweight), scales,
                             reinterpret_cast<uint32_t*>(qzeros), groupsize);
  return 0;
}

__attribute__((visibility("default"))) int refcpu_block_k() { return cpu::block_k; }
__attribute__((visibility("default"))) uint16_t refcpu_float_to_half(float f) { return cpu::float_to_half(f); }

}  // extern "C"
