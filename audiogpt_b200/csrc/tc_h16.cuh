// fp16 hi/lo helpers shared by the kind::f16 tcgen05 kernels (tcconv5.cu, tcconv6.cu).
#pragma once
#include <cuda_fp16.h>
#include "tapconv.cuh"
#include "tc_common.cuh"

namespace agpt {
namespace {

constexpr int H_KCH = 64;   // channels per K chunk = one 128-byte swizzle span of fp16

__device__ __forceinline__ float4 pro_apply5(const TapConvParams& P, float4 v, bool ok, const float* pv) {
  if (P.pro == PRO_LRELU) {
    v.x = lrelu(v.x, P.slope); v.y = lrelu(v.y, P.slope); v.z = lrelu(v.z, P.slope); v.w = lrelu(v.w, P.slope);
  } else if (P.pro == PRO_ADDVEC) {
    if (ok) {
      const float4 a = *reinterpret_cast<const float4*>(pv);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
  } else if (P.pro == PRO_SILU) {
    v.x = siluf_(v.x); v.y = siluf_(v.y); v.z = siluf_(v.z); v.w = siluf_(v.w);
  }
  return v;
}

// two floats -> packed f16x2 (round to nearest, saturate to +-65504): lower half = a, upper half = b
__device__ __forceinline__ uint32_t f2h2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// hi/lo split of two floats; returns the packed hi pair, writes the packed lo pair
__device__ __forceinline__ uint32_t split2(float a, float b, uint32_t& lo) {
  const uint32_t h = f2h2_sat(a, b);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h));
  lo = f2h2_sat(a - hf.x, b - hf.y);
  return h;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

}  // namespace
}  // namespace agpt
