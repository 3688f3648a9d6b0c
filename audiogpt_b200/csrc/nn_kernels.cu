// Non-contraction kernels of the UNet path (channels-last rows [N][HW][C]):
// GroupNorm(+SiLU), LayerNorm, attention (fp32, online softmax), timestep embedding,
// concat / nearest-upsample / stride-2 im2col gathers, DDIM update.
// Reference semantics: ldm/modules/diffusionmodules/util.py:151-171,214-216;
// ldm/modules/attention.py:76-77,170-193,203-215; ldm/models/diffusion/ddim.py:198-225.
#include "common.cuh"
#include "models.h"
#include "nn_kernels.h"

namespace agpt {

// ------------------------------------------------------------------ GroupNorm
// pass 1: per (n, row-split, group) partial sum / sum of squares, accumulated in double
// (per-thread channel sums -> shared-memory per-group reduction); grid (S, N)
constexpr int GN_MAX_SPLIT = 32;
__global__ void gn_partial_kernel(const float* __restrict__ x, double* __restrict__ part, int HW, int C, int S, int G) {
  __shared__ double sg[64][2];
  const int n = blockIdx.y, s = blockIdx.x;
  const int r0 = (int)((long)HW * s / S), r1 = (int)((long)HW * (s + 1) / S);
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) { sg[g][0] = 0.0; sg[g][1] = 0.0; }
  __syncthreads();
  const float* xb = x + (long)n * HW * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int r = r0; r < r1; ++r) {
      const double v = (double)__ldg(xb + (long)r * C + c);
      a += v; b += v * v;
    }
    atomicAdd(&sg[c / cpg][0], a);
    atomicAdd(&sg[c / cpg][1], b);
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double* o = part + (((long)n * S + s) * G + g) * 2;
    o[0] = sg[g][0]; o[1] = sg[g][1];
  }
}

// pass 2: every CTA folds the S x G partials of its sample (tiny) and normalises a slab of rows,
// 128-bit loads/stores along the channel axis
__global__ void gn_apply_kernel(const float* __restrict__ x, const double* __restrict__ part,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                float* __restrict__ y, int HW, int C, int S, int G, float eps, int silu, int rows_per_cta) {
  __shared__ float mean[64], rstd[64];
  const int n = blockIdx.y;
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int s = 0; s < S; ++s) {
      const double* p = part + (((long)n * S + s) * G + g) * 2;
      a += p[0]; b += p[1];
    }
    const double cnt = (double)HW * cpg;
    const double m = a / cnt;
    double var = b / cnt - m * m;
    if (var < 0.0) var = 0.0;
    mean[g] = (float)m;
    rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(HW, r0 + rows_per_cta);
  const float* xb = x + (long)n * HW * C;
  float* yb = y + (long)n * HW * C;
  const int c4n = C >> 2;
  const int total = (r1 - r0) * c4n;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int r = r0 + i / c4n, c = (i % c4n) << 2;
    const float4 v = __ldg(reinterpret_cast<const float4*>(xb + (long)r * C + c));
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
    const int g0 = c / cpg, g1 = (c + 1) / cpg, g2 = (c + 2) / cpg, g3 = (c + 3) / cpg;
    float4 o;
    o.x = (v.x - mean[g0]) * rstd[g0] * ga.x + be.x;
    o.y = (v.y - mean[g1]) * rstd[g1] * ga.y + be.y;
    o.z = (v.z - mean[g2]) * rstd[g2] * ga.z + be.z;
    o.w = (v.w - mean[g3]) * rstd[g3] * ga.w + be.w;
    if (silu) { o.x = siluf_(o.x); o.y = siluf_(o.y); o.z = siluf_(o.z); o.w = siluf_(o.w); }
    *reinterpret_cast<float4*>(yb + (long)r * C + c) = o;
  }
}

void groupnorm(const float* x, float* y, const float* gamma, const float* beta, int N, int HW, int C, int G,
               float eps, bool silu, double* scratch, cudaStream_t st) {
  AGPT_CHECK(C % G == 0 && G <= 64 && C % 4 == 0, "GroupNorm: channels must be divisible by groups (<= 64) and by 4");
  const int S = std::max(1, std::min(GN_MAX_SPLIT, HW / 8));
  gn_partial_kernel<<<dim3(S, N), 256, 0, st>>>(x, scratch, HW, C, S, G);
  // ~2 waves of CTAs over 148 SMs
  const int target_ctas = std::max(1, 296 / N);
  const int rows_per_cta = std::max(1, cdiv(HW, target_ctas));
  gn_apply_kernel<<<dim3(cdiv(HW, rows_per_cta), N), 256, 0, st>>>(
      x, scratch, gamma, beta, y, HW, C, S, G, eps, silu ? 1 : 0, rows_per_cta);
  count_launch(2);
  AGPT_CUDA(cudaGetLastError());
}
size_t groupnorm_scratch_doubles(int N, int C) { (void)C; return (size_t)N * GN_MAX_SPLIT * 64 * 2; }

// ------------------------------------------------------------------ LayerNorm (one warp per row)
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float* __restrict__ y, long rows, int C, float eps) {
  const long row = (long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; v += d * d; }
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rstd = rsqrtf(v / (float)C + eps);
  float* yr = y + row * C;
  for (int c = lane; c < C; c += 32) yr[c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
}

void layernorm(const float* x, float* y, const float* gamma, const float* beta, long rows, int C, float eps,
               cudaStream_t st) {
  const int wpb = 8;
  layernorm_kernel<<<(unsigned)cdivl(rows, wpb), wpb * 32, 0, st>>>(x, gamma, beta, y, rows, C, eps);
  count_launch(1);
  AGPT_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ attention
// softmax_j(q_i . k_j * scale) v_j, heads outermost in the channel dim ('b n (h d)').
// Two lanes share one query row (each holds half of the head dim); K/V tiles of 32 keys
// are staged in shared memory and read as warp-wide broadcasts.
template <int DH>
__global__ void __launch_bounds__(128) attention_kernel(
    const float* __restrict__ q, int q_pitch, const float* __restrict__ k, int k_pitch,
    const float* __restrict__ v, int v_pitch, float* __restrict__ o, int o_pitch,
    int Lq, int Lk, float scale) {
  constexpr int D = 2 * DH, KT = 32;
  __shared__ __align__(16) float Ks[KT][D];
  __shared__ __align__(16) float Vs[KT][D];
  const int n = blockIdx.z, h = blockIdx.y;
  const int half = threadIdx.x & 1;
  const int qi = blockIdx.x * 64 + (threadIdx.x >> 1);
  const bool valid = qi < Lq;
  float qr[DH], acc[DH];
  {
    const float* qp = q + ((long)n * Lq + (valid ? qi : 0)) * q_pitch + h * D + half * DH;
#pragma unroll
    for (int c = 0; c < DH; c += 4) {
      const float4 t = *reinterpret_cast<const float4*>(qp + c);
      qr[c] = t.x * scale; qr[c + 1] = t.y * scale; qr[c + 2] = t.z * scale; qr[c + 3] = t.w * scale;
    }
#pragma unroll
    for (int c = 0; c < DH; ++c) acc[c] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  const float* kb = k + (long)n * Lk * k_pitch + h * D;
  const float* vb = v + (long)n * Lk * v_pitch + h * D;
  for (int j0 = 0; j0 < Lk; j0 += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * (D / 4); i += blockDim.x) {
      const int j = i / (D / 4), c = (i % (D / 4)) * 4;
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (j0 + j < Lk) {
        kk = *reinterpret_cast<const float4*>(kb + (long)(j0 + j) * k_pitch + c);
        vv = *reinterpret_cast<const float4*>(vb + (long)(j0 + j) * v_pitch + c);
      }
      *reinterpret_cast<float4*>(&Ks[j][c]) = kk;
      *reinterpret_cast<float4*>(&Vs[j][c]) = vv;
    }
    __syncthreads();
    float s[KT];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < DH; c += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(&Ks[j][half * DH + c]);
        d = fmaf(qr[c], kk.x, d); d = fmaf(qr[c + 1], kk.y, d); d = fmaf(qr[c + 2], kk.z, d); d = fmaf(qr[c + 3], kk.w, d);
      }
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      if (j0 + j >= Lk) d = -INFINITY;
      s[j] = d;
      tmax = fmaxf(tmax, d);
    }
    const float mn = fmaxf(m, tmax);
    const float corr = (m == -INFINITY) ? 0.f : expf(m - mn);
    l *= corr;
#pragma unroll
    for (int c = 0; c < DH; ++c) acc[c] *= corr;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const float p = expf(s[j] - mn);   // exp(-inf) = 0 for masked keys
      l += p;
#pragma unroll
      for (int c = 0; c < DH; c += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(&Vs[j][half * DH + c]);
        acc[c] = fmaf(p, vv.x, acc[c]); acc[c + 1] = fmaf(p, vv.y, acc[c + 1]);
        acc[c + 2] = fmaf(p, vv.z, acc[c + 2]); acc[c + 3] = fmaf(p, vv.w, acc[c + 3]);
      }
    }
    m = mn;
  }
  if (valid) {
    const float inv = 1.f / l;
    float* op = o + ((long)n * Lq + qi) * o_pitch + h * D + half * DH;
#pragma unroll
    for (int c = 0; c < DH; c += 4)
      *reinterpret_cast<float4*>(op + c) = make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
  }
}

void attention(const float* q, int q_pitch, const float* k, int k_pitch, const float* v, int v_pitch,
               float* o, int o_pitch, int N, int heads, int d, int Lq, int Lk, cudaStream_t st) {
  const float scale = 1.0f / sqrtf((float)d);   // dim_head ** -0.5  (attention.py:158)
  dim3 grid(cdiv(Lq, 64), heads, N);
#define AGPT_ATT(DH_) attention_kernel<DH_><<<grid, 128, 0, st>>>(q, q_pitch, k, k_pitch, v, v_pitch, o, o_pitch, Lq, Lk, scale)
  switch (d) {
    case 8: AGPT_ATT(4); break;
    case 16: AGPT_ATT(8); break;
    case 32: AGPT_ATT(16); break;
    case 40: AGPT_ATT(20); break;
    case 64: AGPT_ATT(32); break;
    case 80: AGPT_ATT(40); break;
    default: throw Error("attention: unsupported head dim " + std::to_string(d) + " (supported: 8,16,32,40,64,80)");
  }
#undef AGPT_ATT
  count_launch(1);
  AGPT_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ timestep embedding (cos || sin)
struct TParam { int t[256]; };
__global__ void timestep_embed_kernel(float* __restrict__ out, const __grid_constant__ TParam tp, int dim) {
  const int n = blockIdx.x, half = dim / 2;
  for (int j = threadIdx.x; j < dim; j += blockDim.x) {
    float v = 0.f;
    if (j < 2 * half) {
      const int i = j < half ? j : j - half;
      const float f = expf(-9.210340371976184f * (float)i / (float)half);   // -ln(10000) * i / half
      const float a = (float)tp.t[n] * f;
      v = j < half ? cosf(a) : sinf(a);
    }
    out[(long)n * dim + j] = v;
  }
}
void timestep_embedding(float* out, const int* t_host, int N, int dim, cudaStream_t st) {
  AGPT_CHECK(N <= 256, "at most 256 samples per UNet call");
  TParam tp;
  for (int i = 0; i < N; ++i) tp.t[i] = t_host[i];
  timestep_embed_kernel<<<N, 128, 0, st>>>(out, tp, dim);
  count_launch(1);
  AGPT_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ gathers
__global__ void concat_kernel(const float* __restrict__ a, int Ca, const float* __restrict__ b, int Cb,
                              float* __restrict__ out, long rows) {
  const int C = Ca + Cb;
  const long total = rows * (C / 4);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / (C / 4);
    const int c = (int)(i % (C / 4)) * 4;
    const float4 v = c < Ca ? *reinterpret_cast<const float4*>(a + r * Ca + c)
                            : *reinterpret_cast<const float4*>(b + r * Cb + (c - Ca));
    *reinterpret_cast<float4*>(out + r * C + c) = v;
  }
}
void concat_channels(const float* a, int Ca, const float* b, int Cb, float* out, long rows, cudaStream_t st) {
  AGPT_CHECK(Ca % 4 == 0 && Cb % 4 == 0, "concat channels must be multiples of 4");
  const long total = rows * ((Ca + Cb) / 4);
  concat_kernel<<<(unsigned)std::min<long>(cdivl(total, 256), 4096), 256, 0, st>>>(a, Ca, b, Cb, out, rows);
  count_launch(1);
  AGPT_CUDA(cudaGetLastError());
}

// nearest x2: out[n][2H][2W][C] = in[n][h/2][w/2][C]   (F.interpolate(scale_factor=2, mode='nearest'))
__global__ void upsample2_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int H, int W, int C) {
  const long total = (long)N * 4 * H * W * (C / 4);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % (C / 4)) * 4;
    long r = i / (C / 4);
    const int wo = (int)(r % (2 * W)); r /= 2 * W;
    const int ho = (int)(r % (2 * H));
    const int n = (int)(r / (2 * H));
    const float4 v = *reinterpret_cast<const float4*>(in + (((long)n * H + ho / 2) * W + wo / 2) * C + c);
    *reinterpret_cast<float4*>(out + (((long)n * 2 * H + ho) * 2 * W + wo) * C + c) = v;
  }
}
void upsample_nearest2(const float* in, float* out, int N, int H, int W, int C, cudaStream_t st) {
  const long total = (long)N * 4 * H * W * (C / 4);
  upsample2_kernel<<<(unsigned)std::min<long>(cdivl(total, 256), 4096), 256, 0, st>>>(in, out, N, H, W, C);
  count_launch(1);
  AGPT_CUDA(cudaGetLastError());
}

// im2col for Conv2d(k3, stride 2, pad 1): col[n][ho][wo][tap*C + c] = x[n][2ho+kh-1][2wo+kw-1][c]
__global__ void im2col_s2_kernel(const float* __restrict__ in, float* __restrict__ col, int N, int H, int W, int C,
                                 int Ho, int Wo) {
  const long total = (long)N * Ho * Wo * 9 * (C / 4);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % (C / 4)) * 4;
    long r = i / (C / 4);
    const int tap = (int)(r % 9); r /= 9;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const int hi = 2 * ho + tap / 3 - 1, wi = 2 * wo + tap % 3 - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = *reinterpret_cast<const float4*>(in + (((long)n * H + hi) * W + wi) * C + c);
    *reinterpret_cast<float4*>(col + ((((long)n * Ho + ho) * Wo + wo) * 9 + tap) * C + c) = v;
  }
}
void im2col_stride2(const float* in, float* col, int N, int H, int W, int C, int Ho, int Wo, cudaStream_t st) {
  const long total = (long)N * Ho * Wo * 9 * (C / 4);
  im2col_s2_kernel<<<(unsigned)std::min<long>(cdivl(total, 256), 4096), 256, 0, st>>>(in, col, N, H, W, C, Ho, Wo);
  count_launch(1);
  AGPT_CUDA(cudaGetLastError());
}

// [N][C][HW] -> [N][HW][Cpad] with zero fill for c >= C (tiny C, e.g. 4 latent channels)
__global__ void cf_to_cl_pad_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int Cpad, int HW, long total) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    const long r = i / Cpad;
    const long n = r / HW, p = r % HW;
    out[i] = c < C ? in[(n * C + c) * HW + p] : 0.f;
  }
}
void cf_to_cl_pad(const float* in, float* out, int N, int C, int Cpad, int HW, cudaStream_t st) {
  const long total = (long)N * HW * Cpad;
  cf_to_cl_pad_kernel<<<(unsigned)std::min<long>(cdivl(total, 256), 4096), 256, 0, st>>>(in, out, C, Cpad, HW, total);
  count_launch(1);
  AGPT_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ DDIM update (ddim.py:198-225)
__global__ void ddim_update_kernel(const float* __restrict__ x, const float* __restrict__ eps2, int single, float s,
                                   float sqrt_at, float sqrt_aprev, float dir_coef, float sigma_t, float sqrt_om,
                                   const float* __restrict__ noise, float temperature, long total, long half_off,
                                   float* __restrict__ x_prev, float* __restrict__ pred_x0) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    float e;
    if (single) e = eps2[i];
    else { const float eu = eps2[i], ec = eps2[half_off + i]; e = eu + s * (ec - eu); }
    const float xv = x[i];
    const float p0 = (xv - sqrt_om * e) / sqrt_at;
    float o = sqrt_aprev * p0 + dir_coef * e;
    if (noise) o += sigma_t * noise[i] * temperature; else o += 0.f;
    x_prev[i] = o;
    if (pred_x0) pred_x0[i] = p0;
  }
}
void ddim_update(const float* x, const float* eps2, int single, float cfg_scale, float a_t, float a_prev,
                 float sigma_t, float sqrt_om, const float* noise, float temperature, int B, long n,
                 float* x_prev, float* pred_x0, cudaStream_t st) {
  const long total = (long)B * n;
  // fp32 scalar algebra exactly as torch.full(...).sqrt() etc. would do it
  const float sqrt_at = sqrtf(a_t), sqrt_aprev = sqrtf(a_prev);
  const float dir_coef = sqrtf(1.0f - a_prev - sigma_t * sigma_t);
  ddim_update_kernel<<<(unsigned)std::min<long>(cdivl(total, 256), 2048), 256, 0, st>>>(
      x, eps2, single, cfg_scale, sqrt_at, sqrt_aprev, dir_coef, sigma_t, sqrt_om, noise, temperature, total, total,
      x_prev, pred_x0);
  count_launch(1);
  AGPT_CUDA(cudaGetLastError());
}

}  // namespace agpt
