// gemv_prims.cuh -- device primitives shared by the sm_100a GEMV kernels (gemv_sm100.cu, gemv_w4p_sm100.cu):
// programmatic dependent launch, mma.sync wrappers, mbarrier / TMA, and the v2 W4 unpack helpers.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "unpack.cuh"

namespace xbit {

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
// D = A * B with a zero C operand (first MMA of a scale group)
__device__ __forceinline__ void mma_m16n8k16_zero(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                  uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}

// ---- mbarrier / TMA primitives
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// 2-D tiled TMA load (cp.async.bulk.tensor; SASS: UTMALDG): one instruction moves a whole box and
// zero-fills anything outside the tensor.  L2 evict-first: every weight byte is used exactly once.
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// activations [8 consecutive k] -> (a0 - a1/16, a4 - a5/16) (a1/16, a5/16) (a2 - a3/16, a6 - a7/16) (a3/16, a7/16)
__device__ __forceinline__ uint4 permute_act8_v2(uint4 v) {
  const __half2 sixteenth = u2h2(0x2C002C00u);   // 0.0625
  uint4 o;
  const __half2 p15 = __hmul2(u2h2(prmt(v.x, v.z, 0x7632)), sixteenth);
  const __half2 p37 = __hmul2(u2h2(prmt(v.y, v.w, 0x7632)), sixteenth);
  o.x = h22u(__hsub2(u2h2(prmt(v.x, v.z, 0x5410)), p15));
  o.y = h22u(p15);
  o.z = h22u(__hsub2(u2h2(prmt(v.y, v.w, 0x5410)), p37));
  o.w = h22u(p37);
  return o;
}

__device__ __forceinline__ void unpack_w4_bytes(uint32_t w, uint32_t (&e)[4]) {
  e[1] = prmt(w, 0u, 0x4240);      // bytes 0, 2 zero-extended into the two halves
  e[3] = prmt(w, 0u, 0x4341);      // bytes 1, 3
  e[0] = e[1] & 0x000F000Fu;
  e[2] = e[3] & 0x000F000Fu;
}

}  // namespace xbit
