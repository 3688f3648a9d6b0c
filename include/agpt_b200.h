/* libagpt_b200 -- C ABI of the B200-native AudioGPT generative back-end.
 *
 * The reference (AIGC-Audio/AudioGPT) is pure Python: it has no FFI layer, its
 * "plugin boundary" is the Python class surface listed in SURVEY.md 8(b).  The
 * drop-in Python classes in audiogpt_b200/ bind these entry points through
 * ctypes (audiogpt_b200/_lib.py); INTEGRATION.md shows the stub.  Each entry
 * point cites the reference interface it replaces.
 *
 * Conventions: every function returns 0 on success, non-zero on error with a
 * thread-local message retrievable through agpt_last_error().  Device pointers
 * are plain fp32 arrays, valid for the stream-ordered duration of the call;
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 * Weights are copied and re-laid-out at create time (the handle owns them).
 * One handle per device; a handle is not re-entrant.  There is no CPU fallback.
 */
#ifndef AGPT_B200_H
#define AGPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct agpt_handle_s* agpt_handle;

const char* agpt_last_error(void);
int agpt_version(void);
/* number of kernels launched by this library in this process so far (bench.py's gpu_launches) */
long long agpt_launch_count(void);
void agpt_destroy(agpt_handle h);
/* Measurement helpers (bench.py): per-launch CUDA-event timing of the tapconv kernel,
 * summed per variant v = {0,1,2: fp32-FMA tiles BN=128/64/32; 3: tcgen05 version}; an fp32-FMA
 * saturation probe returning the measured TFLOP/s of the current device.       */
int agpt_profile_enable(int on);
int agpt_profile_collect(double ms[4], double flops[4], double bytes[4], long long launches[4]);
/* dev tooling: one text line per recorded launch ("variant G L Cin Cout ntaps span epi Wreal ms flops"); returns bytes written or -1 */
long agpt_profile_dump(char* out, long cap);
double agpt_fma_peak_tflops(void);
/* 1 (default): contractions run on the tcgen05 tensor cores with error-compensated fp16 parts
 * ("3xfp16": x = hi + lo, products hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM; weights pre-scaled
 * by a power of two per layer); 0: the fp32-FMA kernels only.  Environment: AGPT_TENSOR_CORES.        */
int agpt_set_tensor_cores(int on);
/* tcgen05 kernel schedule: 6 (default) = persistent kernel (tcconv6) where a CTA gets more than one tile,
 * one-tile-per-CTA kernel (tcconv5) otherwise; 5 = tcconv5 only; 7 = tcconv6 forced; -1 = environment
 * (AGPT_TC_V) / default.                                                                               */
int agpt_set_tc_version(int v);
/* Multi-head attention softmax_j(q_i . k_j * d^-0.5) v_j, heads outermost in the channel dim ('b n (h d)'),
 * CrossAttention.forward (ldm/modules/attention.py:170-193): q [N][Lq][q_pitch], k / v [N][Lk][pitch] device rows with
 * head h at channels [h*d, (h+1)*d); o [N][Lq][o_pitch].  Default: QK^T and PV on the tcgen05 tensor cores
 * (error-compensated fp16 parts, fp32 accumulation, exact online softmax; d in {8,16,32,40,64,80});
 * agpt_set_attention_tc(0) / AGPT_ATTN_TC=0 selects the fp32-FMA kernel, (2) / =2 routes the call through the
 * plane-fed kernel the UNet uses internally (q / k / v split into fp16 hi/lo planes first; test / A-B route).      */
int agpt_attention(const float* q, int q_pitch, const float* k, int k_pitch, const float* v, int v_pitch, float* o,
                   int o_pitch, int N, int heads, int d, int Lq, int Lk, void* stream);
int agpt_set_attention_tc(int on);
/* Micro-benchmark of one tapconv layer (random data): out3 = {ms per launch, algorithmic TFLOP/s,
 * max |tcgen05 - fp32 FMA| when check != 0}; dbg8 (tcgen05 only) = average per-CTA phase cycles
 * {setup, first activation tile, MMA issue loop, drain, epilogue, total, wait-on-activations,
 * wait-on-weights}.  Wreal > 0 selects a 3x3 conv on an (L/Wreal) x Wreal image.              */
int agpt_bench_tapconv(int G, int L, int Cin, int Cout, int K, int dil, int Wreal, int epi_res, int use_tc,
                       int reps, int check, double* out3, double* dbg8_or_null);
/* Numerics probe of one layer: activations ~ N(0,1) * x_scale, weights with a weight-norm-like gain spread
 * (output channel gains log-uniform over a factor w_spread); runs the selected tcgen05 kernel and the fp32-FMA
 * kernel on the same data.  rel2 = {max |diff| / rms(ref), rms(diff) / rms(ref)} (1e30 if anything is not finite). */
int agpt_check_tapconv(int G, int L, int Cin, int Cout, int K, int dil, int Wreal, int epi_res, double x_scale,
                       double w_spread, double* rel2);

/* ------------------------------------------------------------------ HiFi-GAN
 * Replaces HifiGanGenerator.__init__/forward/remove_weight_norm
 * (NeuralSeq/modules/hifigan/hifigan.py:104-178) and the twin
 * text_to_audio/Make_An_Audio/vocoder/hifigan/modules.py:86-136.            */
#define AGPT_MAX_UPS 8
#define AGPT_MAX_RBK 8
#define AGPT_MAX_DIL 8
typedef struct {
  int n_mels;                  /* 80 */
  int c_out;                   /* 1 */
  int upsample_initial_channel;
  int num_upsamples;
  int upsample_rates[AGPT_MAX_UPS];
  int upsample_kernel_sizes[AGPT_MAX_UPS];
  int resblock_type;           /* 1 = ResBlock1 (hifigan.py:30-67), 2 = ResBlock2 (:70-91) */
  int num_kernels;
  int resblock_kernel_sizes[AGPT_MAX_RBK];
  int resblock_num_dilations[AGPT_MAX_RBK];
  int resblock_dilations[AGPT_MAX_RBK][AGPT_MAX_DIL];
  int use_nsf;                 /* h['use_pitch_embed']: noise_convs present (hifigan.py:111-132) */
  /* BigVGAN (text_to_audio/Make_An_Audio/vocoder/bigvgan/models.py:133-203) is the same generator with
   * anti-aliased periodic activations (Activation1d(Snake|SnakeBeta)) instead of leaky-relu:
   * activation 0 = leaky-relu (HiFi-GAN), 1 = 'snake', 2 = 'snakebeta'; snake_logscale = h.snake_logscale.
   * With activation != 0 the weight list is: conv_pre w,b; ups w,b ...; per resblock: convs1 w,b x nd,
   * convs2 w,b x nd, then alpha[, beta] for activations.0 .. (2 nd - 1); activation_post alpha[, beta];
   * conv_post w,b; the 12 filter taps (Activation1d's registered buffer).                          */
  int activation;
  int snake_logscale;
} agpt_hifigan_cfg;

/* host_weights: fp32 HOST arrays in the key order of
 * audiogpt_b200.specs.hifigan_param_shapes(h) (weight-norm already folded;
 * m_source.l_linear.* entries are skipped by the library).                   */
int agpt_hifigan_create(const agpt_hifigan_cfg* cfg, const float* const* host_weights,
                        int n_weights, int device, agpt_handle* out);
/* mel [B,n_mels,T] (device) -> wav [B,c_out,T*prod(rates)] (device).
 * har_source: NULL or the merged NSF excitation [B,1,T*prod(rates)] (device),
 * i.e. SourceModuleHnNSF's first output (hifigan.py:145-149).                */
int agpt_hifigan_forward(agpt_handle h, const float* mel, const float* har_source,
                         int B, int T, float* wav, void* stream);
/* Same through HOST buffers: H2D copy, forward, D2H copy, synchronise.
 * This is what HifiGAN.spec2wav (NeuralSeq/vocoders/hifigan.py:55-69) calls. */
int agpt_hifigan_vocode_host(agpt_handle h, const float* mel_host, const float* har_host,
                             int B, int T, float* wav_host);

/* NSF harmonic source: replaces SineGen.forward + SourceModuleHnNSF.forward
 * (NeuralSeq/modules/parallel_wavegan/models/source.py:311-441,484-532; call site hifigan.py:145-149).
 * f0 [B][L] device, ALREADY upsampled to the sample rate (hifigan.py:147 f0_upsamp); dim = harmonic_num + 1 (<= 16);
 * lin_w_host [dim], lin_b: m_source.l_linear; rand_ini [B][dim] (torch.rand, entry 0 ignored) and noise [B][L][dim]
 * (torch.randn_like(sines)) are drawn by the caller in the reference's order, NULL = zeros.
 * har_source [B][L] = tanh(l_linear(sines * uv + noise_amp * noise)): the array agpt_hifigan_forward takes.
 * The phase prefix sum over the whole utterance is a three-level scan in fp64.                            */
int agpt_nsf_source(const float* f0, int B, int L, int dim, float sampling_rate, const float* lin_w_host, float lin_b,
                    const float* rand_ini_or_null, const float* noise_or_null, float sine_amp, float noise_std,
                    float voiced_threshold, float* har_source, void* stream);

/* ------------------------------------------------------------------ DiffNet + GaussianDiffusion
 * Replaces DiffNet.forward (NeuralSeq/modules/diff/net.py:107-130) and the
 * elementwise part of GaussianDiffusion.p_sample / p_sample_plms
 * (NeuralSeq/modules/diff/shallow_diffusion_tts.py:134-204).                 */
typedef struct {
  int in_dims;                 /* 80 mel bins */
  int hidden_size;             /* encoder_hidden: channels of cond */
  int residual_layers;
  int residual_channels;
  int dilation_cycle_length;
} agpt_diffnet_cfg;

int agpt_diffnet_create(const agpt_diffnet_cfg* cfg, const float* const* host_weights,
                        int n_weights, int device, agpt_handle* out);
/* Hoist the step-invariant conditioner projections of all layers
 * (net.py:68: conditioner_projection(cond)) for cond [B,hidden,T] (device).  */
int agpt_diffnet_set_cond(agpt_handle h, const float* cond, int B, int T, void* stream);
/* eps [B,1,M,T] = DiffNet(x [B,1,M,T], t [B] (host ints), cond set above)    */
int agpt_diffnet_eps(agpt_handle h, const float* x, const int* t_host, float* eps, void* stream);
/* x_out = p_sample(x, t, noise): eps is taken from `eps` when non-NULL (any
 * denoise_fn), else computed by the DiffNet handle `h` (cond set above).
 * coef_host[b] = {sqrt_recip_ac[t], sqrt_recipm1_ac[t], post_mean_coef1[t],
 * post_mean_coef2[t], exp(0.5*post_log_var[t]) * (t!=0)} gathered on the host
 * from the fp32 tables (shallow_diffusion_tts.py:108-123,159-166).  clip: clamp
 * x0 to [-1,1] (:153-154).  n_per_sample = M*T.  x_out may alias x.           */
int agpt_gd_p_sample(agpt_handle h_or_null, const float* x, const float* eps_or_null, const int* t_host,
                     const float* coef_host /*[B][5]*/, const float* noise_or_null, int clip_denoised,
                     int B, long n_per_sample, float* x_out, void* stream);
/* The whole ancestral loop of GaussianDiffusion.forward(infer=True) (shallow_diffusion_tts.py:263-272) on the
 * device: for t = t_hi-1 .. t_lo: x_io <- p_sample(x_io, t, noises[t - t_lo]) with every sample at the same t, the
 * DiffNet handle `h` (cond set by agpt_diffnet_set_cond) predicting eps.  coef_host [t_hi-t_lo][5]: the rows of
 * agpt_gd_p_sample in SAMPLING order (row k belongs to t = t_hi-1-k).  noises_or_null: device
 * [t_hi-t_lo][noise_step_stride] floats, pre-drawn by the caller in the reference's RNG call order.  The
 * step-embedding MLP and the per-layer diffusion projections run once for all steps; step 0 runs as plain
 * launches, then ONE captured step (CUDA graph + device-side step counter) is replayed.  AGPT_GRAPH=0 disables. */
int agpt_gd_sample_loop(agpt_handle h, float* x_io, int t_hi, int t_lo, const float* coef_host,
                        const float* noises_or_null, long noise_step_stride, int clip_denoised, void* stream);
long agpt_diffnet_launches_per_step(agpt_handle h);
/* generic elementwise: out = a0*x + a1*e0 + a2*e1 + a3*e2 + a4*e3 (per-sample
 * coefficient rows coef_host[B][5]; NULL e_i are skipped) -- the PLMS
 * combinations of shallow_diffusion_tts.py:174-204.                          */
int agpt_axpby5(const float* x, const float* e0, const float* e1, const float* e2, const float* e3,
                const float* coef_host, int B, long n_per_sample, float* out, void* stream);

/* ------------------------------------------------------------------ UNet + DDIM
 * Replaces UNetModel.forward (ldm/modules/diffusionmodules/openaimodel.py:711-744)
 * and DDIMSampler.p_sample_ddim's arithmetic (ldm/models/diffusion/ddim.py:168-225). */
#define AGPT_MAX_LEVELS 8
typedef struct {
  int in_channels, out_channels, model_channels;
  int num_res_blocks;
  int num_levels;
  int channel_mult[AGPT_MAX_LEVELS];
  int attn_at_level[AGPT_MAX_LEVELS];   /* 1 if ds=2^level is in attention_resolutions */
  int num_heads;                        /* -1 => use num_head_channels */
  int num_head_channels;
  int transformer_depth;
  int context_dim;
} agpt_unet_cfg;

int agpt_unet_create(const agpt_unet_cfg* cfg, const float* const* host_weights,
                     int n_weights, int device, agpt_handle* out);
/* Hoist the step-invariant to_k/to_v(context) of every cross-attention
 * (attention.py:174-176) for context [N,S,context_dim] (device).             */
int agpt_unet_set_context(agpt_handle h, const float* context, int N, int S, void* stream);
/* eps [N,Cout,H,W] = UNet(x [N,Cin,H,W], t [N] host ints, context set above) */
int agpt_unet_forward(agpt_handle h, const float* x, const int* t_host, int N, int H, int W,
                      float* eps, void* stream);
/* One DDIM step with classifier-free guidance on a doubled batch
 * (ddim.py:177-225): eps2 = [e_uncond ; e_cond] (2B samples), e = e_u + s*(e_c-e_u);
 * pred_x0 = (x - sqrt_om*e)/sqrt(a_t); x_prev = sqrt(a_prev)*pred_x0 +
 * sqrt(1-a_prev-sigma^2)*e + sigma*noise*temperature.  If eps2_is_single != 0
 * eps2 holds B samples and no guidance is applied.                           */
int agpt_ddim_update(const float* x, const float* eps2, int eps2_is_single, float cfg_scale,
                     float a_t, float a_prev, float sigma_t, float sqrt_one_minus_at,
                     const float* noise, float temperature, int B, long n_per_sample,
                     float* x_prev, float* pred_x0_or_null, void* stream);
/* Whole DDIM loop on device (ddim.py:117-166, eta = 0) with CFG; context = [uncond ; cond]
 * set through agpt_unet_set_context (2B rows) or B rows when cfg_scale == 1.
 * tables: host arrays of length S in *sampling order* (index S-1 first).  The time-embedding
 * MLP and the ResBlock embedding projections run once for all S timesteps; step 0 runs as plain
 * launches, then ONE captured step (CUDA graph, device-side step counter and coefficient tables)
 * is replayed S-1 times; the step's x_prev update reads its scalars from the table (no host sync,
 * no per-step H2D).  pred_x0_or_null receives the last step's pred_x0 (ddim.py:216).
 * AGPT_GRAPH=0 in the environment replaces the replays by plain launches.                     */
int agpt_unet_ddim_sample(agpt_handle h, const float* x_T, int B, int H, int W, int S,
                          const int* t_steps_host, const float* a_t, const float* a_prev,
                          const float* sigma, const float* sqrt_om, float cfg_scale,
                          float* x_out, float* pred_x0_or_null, void* stream);
/* kernels launched per DDIM step by the last agpt_unet_ddim_sample call (bench.py reports it) */
long agpt_unet_launches_per_step(agpt_handle h);

/* ------------------------------------------------------------------ AutoencoderKL.decode (first stage)
 * Replaces AutoencoderKL.decode = post_quant_conv -> Decoder.forward
 * (text_to_audio/Make_An_Audio/ldm/models/autoencoder.py:351-354,
 *  ldm/modules/diffusionmodules/model.py:462-568; ResnetBlock :121-143, AttnBlock :150-203, Upsample :43-49):
 * the step between DDIMSampler.sample and the vocoder on every text-to-audio call
 * (ddpm.py decode_first_stage; audio-chatgpt.py:174).                                                  */
typedef struct {
  int embed_dim, z_channels;            /* 4, 4 */
  int ch, out_ch;                       /* 128, 1 */
  int num_levels;                       /* len(ch_mult) */
  int ch_mult[AGPT_MAX_LEVELS];
  int num_res_blocks;                   /* the decoder uses num_res_blocks + 1 blocks per level */
  int attn_at_level[AGPT_MAX_LEVELS];   /* 1 if resolution / 2^level is in attn_resolutions (model.py:481,517) */
} agpt_vae_cfg;
/* host_weights: fp32 HOST arrays in the key order of audiogpt_b200.specs.vae_decoder_param_shapes(cfg). */
int agpt_vae_create(const agpt_vae_cfg* cfg, const float* const* host_weights, int n_weights, int device,
                    agpt_handle* out);
/* z [B, embed_dim, H, W] (device) -> out [B, out_ch, H * 2^(levels-1), W * 2^(levels-1)] (device) */
int agpt_vae_decode(agpt_handle h, const float* z, int B, int H, int W, float* out, void* stream);

/* ------------------------------------------------------------------ PitchExtractor
 * Replaces PitchExtractor.forward (NeuralSeq/modules/fastspeech/pe.py:119-148: Prenet :7-42, ConvStacks :82-116,
 * PitchPredictor modules/fastspeech/tts_modules.py:217-260, denorm_f0 utils/pitch_utils.py:63-76): F0 from a generated
 * mel for the NSF vocoder (inference/svs/base_svs_infer.py run_vocoder).                                          */
typedef struct {
  int n_mel_bins;          /* 80 */
  int hidden_size;         /* hparams['hidden_size'] */
  int conv_layers;         /* ConvStacks layers (2) */
  int predictor_hidden;    /* hparams['predictor_hidden'] or hidden_size */
  int predictor_layers;    /* 5 */
  int predictor_kernel;    /* hparams['predictor_kernel'] */
} agpt_pe_cfg;
/* host_weights: fp32 HOST arrays in the key order of audiogpt_b200.specs.pe_param_shapes(cfg) (state-dict order incl. the
 * BatchNorm buffers; num_batches_tracked and embed_positions._float_tensor are passed and ignored).               */
int agpt_pe_create(const agpt_pe_cfg* cfg, const float* const* host_weights, int n_weights, int device, agpt_handle* out);
/* mel [B][T][n_mel_bins] (device, channels-last as the reference passes it) -> pitch_pred [B][T][2], f0_denorm [B][T].
 * pitch_norm: 0 none, 1 'standard' (f0 * std + mean), 2 'log' (2 ** f0); use_uv: zero F0 where pitch_pred[..., 1] > 0;
 * all-zero mel frames (padding) give F0 = 0.                                                                      */
int agpt_pe_forward(agpt_handle h, const float* mel, int B, int T, float* pitch_pred, float* f0_denorm, int use_uv,
                    int pitch_norm, float f0_mean, float f0_std, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AGPT_B200_H */
