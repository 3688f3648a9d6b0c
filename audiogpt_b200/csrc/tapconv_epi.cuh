// Shared fused epilogue of the tapconv kernels (fp32 FMA version and tcgen05 version):
// one call handles 4 consecutive output channels of one output row.
#pragma once
#include "tapconv.cuh"

namespace agpt {

__device__ __forceinline__ void tc_epilogue(const TapConvParams& P, int g, int p, int co, float4 v) {
  if (co >= P.Cout) return;
  if (P.bias) {
    const float4 b = *reinterpret_cast<const float4*>(P.bias + co);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  switch (P.epi) {
    case EPI_BIAS: break;
    case EPI_RES:
    case EPI_ACC: {
      if (P.res) {
        const float4 r = *reinterpret_cast<const float4*>(P.res + g * P.res_gstride + (long)p * P.res_pitch + co);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      if (P.epi == EPI_ACC) {
        v.x *= P.scale; v.y *= P.scale; v.z *= P.scale; v.w *= P.scale;
        if (P.accumulate) {
          const float4 o = *reinterpret_cast<const float4*>(P.out + g * P.out_gstride + (long)p * P.out_pitch + co);
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
      }
      break;
    }
    case EPI_RELU:
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      break;
    case EPI_TANH:
      v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w);
      break;
    case EPI_MISH:
      v.x = mishf_(v.x); v.y = mishf_(v.y); v.z = mishf_(v.z); v.w = mishf_(v.w);
      break;
    case EPI_SILU:
      v.x = siluf_(v.x); v.y = siluf_(v.y); v.z = siluf_(v.z); v.w = siluf_(v.w);
      break;
    case EPI_ADDVEC: {
      const float4 e = *reinterpret_cast<const float4*>(P.evec + (long)g * P.evec_gstride + co);
      v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
      break;
    }
    case EPI_GATE:
    case EPI_GEGLU: {
      if (P.res) {  // pre-activation additive term (DiffNet hoisted conditioner projection)
        const float4 r = *reinterpret_cast<const float4*>(P.res + g * P.res_gstride + (long)p * P.res_pitch + co);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      float2 o;
      if (P.epi == EPI_GATE) {
        o.x = sigmoidf_(v.x) * tanhf(v.y);
        o.y = sigmoidf_(v.z) * tanhf(v.w);
      } else {
        o.x = v.x * gelu_erf(v.y);
        o.y = v.z * gelu_erf(v.w);
      }
      *reinterpret_cast<float2*>(P.out + g * P.out_gstride + (long)p * P.out_pitch + (co >> 1)) = o;
      return;
    }
    case EPI_DIFFOUT: {
      if (co < P.csplit) {
        float4* o = reinterpret_cast<float4*>(P.out + g * P.out_gstride + (long)p * P.out_pitch + co);
        float4 x = *o;
        const float r2 = 0.70710678118654752440f;
        x.x = (x.x + v.x) * r2; x.y = (x.y + v.y) * r2; x.z = (x.z + v.z) * r2; x.w = (x.w + v.w) * r2;
        *o = x;
      } else {
        float4* o = reinterpret_cast<float4*>(P.out2 + g * P.out2_gstride + (long)p * P.out2_pitch + (co - P.csplit));
        if (P.accumulate) {
          float4 s = *o;
          v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
        }
        *o = v;
      }
      return;
    }
    case EPI_STORE_CF: {
      float* o = P.out + g * P.out_gstride + (long)co * P.L + p;
      o[0] = v.x;
      if (co + 1 < P.Cout) o[(long)P.L] = v.y;
      if (co + 2 < P.Cout) o[2 * (long)P.L] = v.z;
      if (co + 3 < P.Cout) o[3 * (long)P.L] = v.w;
      return;
    }
    default: break;
  }
  *reinterpret_cast<float4*>(P.out + g * P.out_gstride + (long)p * P.out_pitch + co) = v;
}

}  // namespace agpt
