// unpack.cuh -- register-level primitives shared by the sm_100a DQ and GEMV kernels.
//
// Replaces the reference's type helpers (/root/reference/src/cu/common.cuh:41-171), which go
// through I2F conversions (__short2half_rn / __int2half_rn), and its branchy bit-stream reader
// iterator_qweight_v2 (/root/reference/src/cu/unpack_weight_2_to_7.cu:196-217), which needs a
// shared-memory slice per thread because it indexes the words at run time.
//
// Here every value is materialised as an EXACT fp16 integer with full-rate ALU ops only:
//   field at mantissa bits [p, p+b) of a 16-bit lane, OR-ed with the exponent pattern whose ulp is
//   2^-p  ->  fp16 value (2^(10-p) + w)   ["magic number"], then one exact HSUB2 of the base.
// PRMT moves whole bytes, LOP3 does mask|magic in one op, SHF funnels across 32-bit words for
// the straddling widths (3, 5, 6, 7).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace xbit {

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
  return r;
}
// (a & b) | c  -- truth table 0xEA
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c) { return lop3<0xEA>(a, b, c); }

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}

__device__ __forceinline__ __half2 u2h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h22u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// streaming 128-bit global load: read-only path, do not allocate in L1 (each byte is used once)
__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// streaming 128-bit global store (evict-first: the dequantised matrix is not re-read by us)
__device__ __forceinline__ void stg_stream_v4(void* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// fp16 bit pattern whose value is 2^(10-p): a b-bit field at mantissa bits [p, p+b) OR-ed into it
// reads as 2^(10-p) + field.  Valid while p + b <= 10.
__host__ __device__ constexpr uint32_t magic_base_bits(int p) { return (uint32_t)((25 - p) << 10); }
__host__ __device__ constexpr uint32_t dup16(uint32_t x) { return x | (x << 16); }

// half2 constant 2^(10-p) in both lanes, as a bit pattern
__host__ __device__ constexpr uint32_t magic2(int p) { return dup16(magic_base_bits(p)); }

// Number of 32-bit words a 32-value block of b-bit fields occupies: exactly b.
// Field i of the block starts at bit i*b: word (i*b)>>5, shift (i*b)&31.

// Extract field i (compile-time) of a b-bit block held in words w[0..B), zero-extended, at bit 0.
template <int B, int I>
__device__ __forceinline__ uint32_t block_field(const uint32_t (&w)[B]) {
  constexpr int pos = I * B, wi = pos >> 5, sh = pos & 31;
  if constexpr (sh + B <= 32) {
    return w[wi] >> sh;                       // caller masks
  } else {
    return __funnelshift_r(w[wi], w[wi + 1], sh);
  }
}

}  // namespace xbit
