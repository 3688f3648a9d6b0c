// DiffNet epsilon-predictor + GaussianDiffusion step arithmetic on sm_100a.
// Reference: NeuralSeq/modules/diff/net.py:58-130 (DiffNet, ResidualBlock, SinusoidalPosEmb),
//            NeuralSeq/modules/diff/shallow_diffusion_tts.py:134-204 (p_sample / PLMS algebra).
// Parity: tests/test_diffusion_gpu.py against oracle/diffusion_ref.py and tests/golden/diffusion_*.npz.
#include "common.cuh"
#include "tapconv.cuh"
#include "models.h"

namespace agpt {

constexpr int kMaxBatchParam = 256;
struct StepT { int t[kMaxBatchParam]; };

// SinusoidalPosEmb (net.py:37-44): sin || cos, exponent divisor (half-1)
__global__ void diff_step_embed_kernel(float* __restrict__ out, const __grid_constant__ StepT st, int B, int C, float neg_emb) {
  const int b = blockIdx.x;
  const int half = C / 2;
  for (int j = threadIdx.x; j < C; j += blockDim.x) {
    const int i = j < half ? j : j - half;
    const float f = expf((float)i * neg_emb);
    const float a = (float)st.t[b] * f;
    out[(long)b * C + j] = j < half ? sinf(a) : cosf(a);
  }
}

// x_out = c1*clamp(A*x - Bc*eps, -1, 1) + c2*x + s*noise      (shallow_diffusion_tts.py:134-166)
// coef[b] = {A, Bc, c1, c2, s}
__global__ void p_sample_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ noise,
                                const float* __restrict__ coef, int clip, long n, float* __restrict__ out) {
  const int b = blockIdx.y;
  const float A = coef[b * 5 + 0], Bc = coef[b * 5 + 1], c1 = coef[b * 5 + 2], c2 = coef[b * 5 + 3], s = coef[b * 5 + 4];
  const long base = (long)b * n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float xv = x[base + i];
    float x0 = A * xv - Bc * eps[base + i];
    if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    float o = c1 * x0 + c2 * xv;
    if (noise) o += s * noise[base + i];
    out[base + i] = o;
  }
}

__global__ void axpby5_kernel(const float* __restrict__ x, const float* __restrict__ e0, const float* __restrict__ e1,
                              const float* __restrict__ e2, const float* __restrict__ e3,
                              const float* __restrict__ coef, long n, float* __restrict__ out) {
  const int b = blockIdx.y;
  const float a0 = coef[b * 5], a1 = coef[b * 5 + 1], a2 = coef[b * 5 + 2], a3 = coef[b * 5 + 3], a4 = coef[b * 5 + 4];
  const long base = (long)b * n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float o = a0 * x[base + i];
    if (e0) o += a1 * e0[base + i];
    if (e1) o += a2 * e1[base + i];
    if (e2) o += a3 * e2[base + i];
    if (e3) o += a4 * e3[base + i];
    out[base + i] = o;
  }
}

// staging for per-sample coefficient rows, one ring per device (stream-ordered reuse)
static DevBuf g_coef[16];
static int g_coef_slot[16] = {0};
constexpr int kCoefSlots = 8, kCoefMaxB = 1024;

static const float* upload_coef(const float* coef_host, int B, cudaStream_t st) {
  int dev = 0;
  AGPT_CUDA(cudaGetDevice(&dev));
  AGPT_CHECK(dev < 16 && B <= kCoefMaxB, "coefficient staging");
  float* base = g_coef[dev].ensure((size_t)kCoefSlots * kCoefMaxB * 5);
  float* slot = base + (size_t)(g_coef_slot[dev]++ % kCoefSlots) * kCoefMaxB * 5;
  AGPT_CUDA(cudaMemcpyAsync(slot, coef_host, (size_t)B * 5 * sizeof(float), cudaMemcpyHostToDevice, st));
  return slot;
}

void axpby5(const float* x, const float* e0, const float* e1, const float* e2, const float* e3,
            const float* coef_host, int B, long n, float* out, cudaStream_t st) {
  const float* coef = upload_coef(coef_host, B, st);
  dim3 grid((unsigned)std::min<long>(cdivl(n, 256), 1184), B);
  axpby5_kernel<<<grid, 256, 0, st>>>(x, e0, e1, e2, e3, coef, n, out);
  count_launch(1);
  AGPT_CUDA(cudaGetLastError());
}

struct Diffnet : Handle {
  agpt_diffnet_cfg cfg;
  PackedConv in_proj, mlp0, mlp2, dproj_all, cond_all, skip_proj, out_proj;
  std::vector<PackedConv> dil, outp;
  // state
  int B = 0, T = 0;
  DevBuf condT, condp, xT, xcur, z, skip, hbuf, emb, e1, e2, dproj, eps_tmp;

  void set_cond(const float* cond, int B_, int T_, cudaStream_t st) {
    const int H = cfg.hidden_size, C = cfg.residual_channels, L = cfg.residual_layers;
    AGPT_CHECK(B_ >= 1 && B_ <= kMaxBatchParam && T_ >= 1, "batch size must be in [1,256]");
    B = B_; T = T_;
    condT.ensure((size_t)B * T * H);
    condp.ensure((size_t)B * T * L * 2 * C);
    launch_cf_to_cl(cond, condT.p, B, H, T, st);
    TapConvParams P = tapconv_params(cond_all, B, T, 0, 1);
    P.in = condT.p; P.in_gstride = (long)T * H; P.in_pitch = H;
    P.out = condp.p; P.out_gstride = (long)T * L * 2 * C; P.out_pitch = L * 2 * C;
    P.epi = EPI_BIAS;
    tapconv_launch(P, st);
  }

  void eps(const float* x, const int* t_host, float* out, cudaStream_t st) {
    AGPT_CHECK(B > 0, "agpt_diffnet_set_cond must be called first");
    const int C = cfg.residual_channels, L = cfg.residual_layers, M = cfg.in_dims;
    const size_t rows = (size_t)B * T;
    xT.ensure(rows * M); xcur.ensure(rows * C); z.ensure(rows * C); skip.ensure(rows * C); hbuf.ensure(rows * C);
    emb.ensure((size_t)B * C); e1.ensure((size_t)B * 4 * C); e2.ensure((size_t)B * C); dproj.ensure((size_t)B * L * C);

    StepT stp;
    for (int b = 0; b < B; ++b) stp.t[b] = t_host[b];
    const float neg_emb = (float)(-(std::log(10000.0) / (double)(C / 2 - 1)));
    diff_step_embed_kernel<<<B, 128, 0, st>>>(emb.p, stp, B, C, neg_emb);
    count_launch(1);
    AGPT_CUDA(cudaGetLastError());
    auto lin = [&](const PackedConv& pc, const float* in, int cin, float* o, int cout, int epi) {
      TapConvParams P = tapconv_params(pc, 1, B, 0, 1);
      P.in = in; P.in_gstride = 0; P.in_pitch = cin;
      P.out = o; P.out_gstride = 0; P.out_pitch = cout;
      P.epi = epi;
      tapconv_launch(P, st);
    };
    lin(mlp0, emb.p, C, e1.p, 4 * C, EPI_MISH);
    lin(mlp2, e1.p, 4 * C, e2.p, C, EPI_BIAS);
    lin(dproj_all, e2.p, C, dproj.p, L * C, EPI_BIAS);

    launch_cf_to_cl(x, xT.p, B, M, T, st);
    {
      TapConvParams P = tapconv_params(in_proj, B, T, 0, 1);
      P.in = xT.p; P.in_gstride = (long)T * M; P.in_pitch = M;
      P.out = xcur.p; P.out_gstride = (long)T * C; P.out_pitch = C;
      P.epi = EPI_RELU;
      tapconv_launch(P, st);
    }
    const long gs = (long)T * C;
    for (int l = 0; l < L; ++l) {
      const int d = 1 << (l % cfg.dilation_cycle_length);
      {  // y = dilated_conv(x + dproj) + cond_proj ; z = sigmoid(gate)*tanh(filter)   (net.py:67-74)
        TapConvParams P = tapconv_params(dil[l], B, T, 0, d);
        P.in = xcur.p; P.in_gstride = gs; P.in_pitch = C;
        P.pro = PRO_ADDVEC; P.pvec = dproj.p + (long)l * C; P.pvec_gstride = L * C;
        P.epi = EPI_GATE;
        P.res = condp.p + (long)l * 2 * C; P.res_gstride = (long)T * L * 2 * C; P.res_pitch = L * 2 * C;
        P.out = z.p; P.out_gstride = gs; P.out_pitch = C;
        tapconv_launch(P, st);
      }
      {  // output_projection; x <- (x + residual)/sqrt2 ; skip += skip_l    (net.py:76-78)
        TapConvParams P = tapconv_params(outp[l], B, T, 0, 1);
        P.in = z.p; P.in_gstride = gs; P.in_pitch = C;
        P.epi = EPI_DIFFOUT; P.csplit = C; P.accumulate = (l > 0);
        P.out = xcur.p; P.out_gstride = gs; P.out_pitch = C;
        P.out2 = skip.p; P.out2_gstride = gs; P.out2_pitch = C;
        tapconv_launch(P, st);
      }
    }
    {  // relu(skip_projection(sum skip / sqrt(L)))   (1/sqrt(L) folded into the weights)
      TapConvParams P = tapconv_params(skip_proj, B, T, 0, 1);
      P.in = skip.p; P.in_gstride = gs; P.in_pitch = C;
      P.out = hbuf.p; P.out_gstride = gs; P.out_pitch = C;
      P.epi = EPI_RELU;
      tapconv_launch(P, st);
    }
    {
      TapConvParams P = tapconv_params(out_proj, B, T, 0, 1);
      P.in = hbuf.p; P.in_gstride = gs; P.in_pitch = C;
      P.out = out; P.out_gstride = (long)M * T; P.out_pitch = 0;
      P.epi = EPI_STORE_CF;
      tapconv_launch(P, st);
    }
  }
};

Handle* diffnet_create(const agpt_diffnet_cfg* cfg, const float* const* W, int nW, int device) {
  AGPT_CUDA(cudaSetDevice(device));
  auto* h = new Diffnet();
  h->magic = kMagicDiffnet; h->device = device; h->cfg = *cfg;
  const int C = cfg->residual_channels, H = cfg->hidden_size, M = cfg->in_dims, L = cfg->residual_layers;
  AGPT_CHECK(C % 8 == 0 && C >= 8 && L >= 1 && cfg->dilation_cycle_length >= 1, "bad DiffNet config");
  AGPT_CHECK(nW == 6 + 8 * L + 4, "weight array count does not match the config");
  int idx = 0;
  auto next = [&]() { return W[idx++]; };
  { auto w = next(); auto b = next(); pack_conv(h->in_proj, w, b, C, M, 1, false); }
  { auto w = next(); auto b = next(); pack_conv(h->mlp0, w, b, 4 * C, C, 1, false); }
  { auto w = next(); auto b = next(); pack_conv(h->mlp2, w, b, C, 4 * C, 1, false); }
  h->dil.resize(L); h->outp.resize(L);
  std::vector<float> dpw((size_t)L * C * C), dpb((size_t)L * C), cw((size_t)L * 2 * C * H), cb((size_t)L * 2 * C);
  for (int l = 0; l < L; ++l) {
    { auto w = next(); auto b = next(); pack_conv_pairs(h->dil[l], w, b, 2 * C, C, 3); }
    { auto w = next(); auto b = next();
      memcpy(&dpw[(size_t)l * C * C], w, sizeof(float) * C * C); memcpy(&dpb[(size_t)l * C], b, sizeof(float) * C); }
    { auto w = next(); auto b = next();   // conditioner: interleave (gate,filter) like the dilated conv
      for (int co = 0; co < 2 * C; ++co) {
        const int dst = 2 * (co % C) + co / C;
        memcpy(&cw[((size_t)l * 2 * C + dst) * H], w + (size_t)co * H, sizeof(float) * H);
        cb[(size_t)l * 2 * C + dst] = b[co];
      } }
    { auto w = next(); auto b = next(); pack_conv(h->outp[l], w, b, 2 * C, C, 1, false); }
  }
  pack_conv(h->dproj_all, dpw.data(), dpb.data(), L * C, C, 1, false);
  pack_conv(h->cond_all, cw.data(), cb.data(), L * 2 * C, H, 1, false);
  { auto w = next(); auto b = next(); pack_conv(h->skip_proj, w, b, C, C, 1, false, 1.f / std::sqrt((float)L)); }
  { auto w = next(); auto b = next(); pack_conv(h->out_proj, w, b, M, C, 1, false); }
  return h;
}

void diffnet_set_cond(Handle* hh, const float* cond, int B, int T, cudaStream_t st) {
  auto* h = static_cast<Diffnet*>(hh);
  AGPT_CUDA(cudaSetDevice(h->device));
  h->set_cond(cond, B, T, st);
}

void diffnet_eps(Handle* hh, const float* x, const int* t_host, float* eps, cudaStream_t st) {
  auto* h = static_cast<Diffnet*>(hh);
  AGPT_CUDA(cudaSetDevice(h->device));
  h->eps(x, t_host, eps, st);
}

void gd_p_sample(Handle* hh, const float* x, const float* eps_or_null, const int* t_host, const float* coef_host,
                 const float* noise, int clip, int B, long n, float* x_out, cudaStream_t st) {
  const float* e = eps_or_null;
  if (!e) {
    auto* h = static_cast<Diffnet*>(hh);
    AGPT_CUDA(cudaSetDevice(h->device));
    AGPT_CHECK(B == h->B && n == (long)h->cfg.in_dims * h->T, "shape differs from the cond set by agpt_diffnet_set_cond");
    h->eps_tmp.ensure((size_t)B * n);
    h->eps(x, t_host, h->eps_tmp.p, st);
    e = h->eps_tmp.p;
  }
  const float* coef = upload_coef(coef_host, B, st);
  dim3 grid((unsigned)std::min<long>(cdivl(n, 256), 1184), B);
  p_sample_kernel<<<grid, 256, 0, st>>>(x, e, noise, coef, clip, n, x_out);
  count_launch(1);
  AGPT_CUDA(cudaGetLastError());
}

}  // namespace agpt
