// tcconv: the tapconv contraction on the 5th-generation tensor cores (tcgen05 + TMEM).
//
//   out[g, p, co] = epi( bias[co] + sum_tap sum_ci pro(in[g, p + off_tap, ci]) * W[tap][ci][co] )
//
// Same TapConvParams contract (and the same fused prologue/epilogue table) as the fp32-FMA kernel
// in tapconv.cu, but the inner product runs as tcgen05.mma.kind::tf32 with the accumulator tile
// [128 rows x BN cols] in tensor memory.  To keep fp32-grade parity (waveform RMSE <= 1e-4 through
// 78 stacked convs) every product is error-compensated ("3xTF32"):
//        x = x_hi + x_lo ,  w = w_hi + w_lo   (hi = top 19 bits, lo = exact remainder)
//        D += x_hi*w_hi + x_lo*w_hi + x_hi*w_lo            (the dropped lo*lo term is ~2^-22)
//
// Roles (192 threads, 1 CTA per SM):
//   warps 0-3  transform: raw activation rows (cp.async, zero-filled halo, XOR-swizzled) -> apply the
//              prologue (LeakyReLU / +vec / SiLU), split hi/lo, write the two K-major SWIZZLE_128B
//              operand tiles of the current tap; afterwards the same warps run the epilogue
//              (tcgen05.ld TMEM -> registers -> fused epilogue -> global).
//   warp 4     one elected thread issues tcgen05.mma (12 per (chunk, tap): 4 k-steps x 3 products)
//              and tcgen05.commit to free operand buffers.
//   warp 5     one elected thread streams pre-swizzled weight tiles (hi|lo) with cp.async.bulk
//              (TMA engine, 1-D) onto an mbarrier.
// Pipelines: A tiles (2 buffers) and W tiles (2 stages) through full/empty mbarriers.
#include "tapconv.cuh"
#include "tapconv_epi.cuh"
#include "tc_common.cuh"
#include "models.h"

namespace agpt {

namespace {


struct TcSmem {
  // dynamic shared memory layout (offsets from a 1024-byte aligned base)
  uint32_t a_hi[2], a_lo[2], w[2], raw[2], rowinfo, bars, tmem_slot, total;
};
__host__ __device__ inline TcSmem tc_layout(int BN, int RR) {
  TcSmem s;
  uint32_t o = 0;
  for (int i = 0; i < 2; ++i) { s.a_hi[i] = o; o += TC_ROWS * 128; }
  for (int i = 0; i < 2; ++i) { s.a_lo[i] = o; o += TC_ROWS * 128; }
  for (int i = 0; i < 2; ++i) { s.w[i] = o; o += 2 * BN * 128; }
  for (int i = 0; i < 2; ++i) { s.raw[i] = o; o += RR * 128; }
  s.rowinfo = o; o += RR * 4;
  o = (o + 15) & ~15u;
  s.bars = o; o += 16 * 8;
  s.tmem_slot = o; o += 16;
  s.total = o;
  return s;
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1) tcconv_kernel(const __grid_constant__ TapConvParams P) {
  extern __shared__ uint8_t smem_raw_[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~(uintptr_t)1023);
  const int RR = P.R;                 // raw rows (multiple of 8)
  const TcSmem S = tc_layout(BN, RR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S.bars);
  uint64_t* a_full = bars + 0;   // [2]
  uint64_t* a_empty = bars + 2;  // [2]
  uint64_t* w_full = bars + 4;   // [2]
  uint64_t* w_empty = bars + 6;  // [2]
  uint64_t* acc_full = bars + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + S.tmem_slot);
  int* rowinfo = reinterpret_cast<int*>(smem + S.rowinfo);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = blockIdx.z, co0 = blockIdx.y * BN, q0 = blockIdx.x * TC_ROWS;
  const int Wv = P.Wreal > 0 ? P.Wreal + 1 : 0;
  const int Lv = Wv ? (P.L / P.Wreal) * Wv : P.L;
  const int nchunks = P.tc_chunks, ntaps = P.ntaps, total = nchunks * ntaps;
  const int lo = P.lo_al;             // min tap offset (not rounded here)

  if (tid == 0) {
    mbar_init(&a_full[0], 128); mbar_init(&a_full[1], 128);
    mbar_init(&a_empty[0], 1); mbar_init(&a_empty[1], 1);
    mbar_init(&w_full[0], 1); mbar_init(&w_full[1], 1);
    mbar_init(&w_empty[0], 1); mbar_init(&w_empty[1], 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void*)tmem_slot)), "r"((uint32_t)(BN < 32 ? 32 : BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp < 4) {
    for (int i = tid; i < RR; i += 128) {
      const int q = q0 + lo + i;
      int a = -1;
      if (q >= 0 && q < Lv) {
        if (Wv) {
          const int h = q / Wv, w = q - h * Wv;
          if (w < P.Wreal) a = (h * P.Wreal + w) * P.in_pitch;
        } else {
          a = q * P.in_pitch;
        }
      }
      rowinfo[i] = a;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =========================== transform warps ===========================
    const float* __restrict__ ing = P.in + g * P.in_gstride;
    auto issue_raw = [&](int c, int buf) {
      uint8_t* dst = smem + S.raw[buf];
      for (int idx = tid; idx < RR * 8; idx += 128) {
        const int row = idx >> 3, j = idx & 7;
        const int ch = c * TC_KCH + 4 * j;
        const int a = rowinfo[row];
        const bool ok = (a >= 0) && (ch < P.Cin);
        const float* src = ok ? (ing + a + ch) : P.in;
        cp_async16_zfill(dst + sw128(row, j), src, ok ? 16u : 0u);
      }
      cp_async_commit_();
    };
    issue_raw(0, 0);
    int it = 0;
    for (int c = 0; c < nchunks; ++c) {
      cp_async_wait_all_();
      named_bar_sync(1, 128);                       // raw[c&1] complete and visible to the 4 warps
      if (c + 1 < nchunks) issue_raw(c + 1, (c + 1) & 1);
      const uint8_t* rawb = smem + S.raw[c & 1];
      float pv[TC_KCH];
      if (P.pro == PRO_ADDVEC) {
#pragma unroll
        for (int k = 0; k < TC_KCH; ++k) {
          const int ch = c * TC_KCH + k;
          pv[k] = ch < P.Cin ? P.pvec[(long)g * P.pvec_gstride + ch] : 0.f;
        }
      }
      for (int t = 0; t < ntaps; ++t, ++it) {
        const int b = it & 1, n = it >> 1;
        if (n >= 1) mbar_wait(&a_empty[b], (uint32_t)((n - 1) & 1));
        const int rr = tid + (P.tap_off[t] - lo);
        const bool rvalid = rowinfo[rr] >= 0;
        uint8_t* ahi = smem + S.a_hi[b];
        uint8_t* alo = smem + S.a_lo[b];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v = *reinterpret_cast<const float4*>(rawb + sw128(rr, j));
          float x[4] = {v.x, v.y, v.z, v.w};
          float hi[4], lw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float xv = x[e];
            if (P.pro == PRO_LRELU) xv = lrelu(xv, P.slope);
            else if (P.pro == PRO_ADDVEC) xv = rvalid ? xv + pv[4 * j + e] : 0.f;
            else if (P.pro == PRO_SILU) xv = siluf_(xv);
            const float h = __uint_as_float(__float_as_uint(xv) & 0xffffe000u);
            hi[e] = h;
            lw[e] = xv - h;
          }
          *reinterpret_cast<float4*>(ahi + sw128(tid, j)) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(alo + sw128(tid, j)) = make_float4(lw[0], lw[1], lw[2], lw[3]);
        }
        fence_proxy_async();
        mbar_arrive(&a_full[b]);
      }
    }
    // =========================== epilogue ===========================
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int r = tid;                 // TMEM lane == output row of the tile
    const int q = q0 + r;
    bool valid = q < Lv;
    int p = q;
    if (valid && Wv) {
      const int h = q / Wv, w = q - h * Wv;
      valid = w < P.Wreal;
      p = h * P.Wreal + w;
    }
#pragma unroll 1
    for (int cb = 0; cb < BN; cb += 32) {
      uint32_t rg[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(rg[0]), "=r"(rg[1]), "=r"(rg[2]), "=r"(rg[3]), "=r"(rg[4]), "=r"(rg[5]), "=r"(rg[6]), "=r"(rg[7]),
            "=r"(rg[8]), "=r"(rg[9]), "=r"(rg[10]), "=r"(rg[11]), "=r"(rg[12]), "=r"(rg[13]), "=r"(rg[14]), "=r"(rg[15]),
            "=r"(rg[16]), "=r"(rg[17]), "=r"(rg[18]), "=r"(rg[19]), "=r"(rg[20]), "=r"(rg[21]), "=r"(rg[22]), "=r"(rg[23]),
            "=r"(rg[24]), "=r"(rg[25]), "=r"(rg[26]), "=r"(rg[27]), "=r"(rg[28]), "=r"(rg[29]), "=r"(rg[30]), "=r"(rg[31])
          : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (valid) {
#pragma unroll
        for (int qd = 0; qd < 8; ++qd) {
          tc_epilogue(P, g, p, co0 + cb + 4 * qd,
                      make_float4(__uint_as_float(rg[4 * qd]), __uint_as_float(rg[4 * qd + 1]),
                                  __uint_as_float(rg[4 * qd + 2]), __uint_as_float(rg[4 * qd + 3])));
        }
      }
    }
  } else if (warp == 4) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      // instruction descriptor: c=F32 (bit4), a=b=TF32 (2<<7, 2<<10), K-major both, N>>3 @17, M>>4 @24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
      for (int it = 0; it < total; ++it) {
        const int b = it & 1, n = it >> 1;
        mbar_wait(&w_full[b], (uint32_t)(n & 1));
        mbar_wait(&a_full[b], (uint32_t)(n & 1));
        tc_fence_after();
        const uint64_t dah = make_desc(smem_u32(smem + S.a_hi[b]));
        const uint64_t dal = make_desc(smem_u32(smem + S.a_lo[b]));
        const uint64_t dwh = make_desc(smem_u32(smem + S.w[b]));
        const uint64_t dwl = make_desc(smem_u32(smem + S.w[b] + BN * 128));
#pragma unroll
        for (int k = 0; k < TC_KCH / 8; ++k) {
          const uint64_t ko = (uint64_t)((k * 32) >> 4);     // advance 32 bytes (8 tf32) along K inside the swizzle span
          umma_tf32(tmem_base, dah + ko, dwh + ko, idesc, (it > 0 || k > 0) ? 1u : 0u);
          umma_tf32(tmem_base, dal + ko, dwh + ko, idesc, 1u);
          umma_tf32(tmem_base, dah + ko, dwl + ko, idesc, 1u);
        }
        umma_commit(&a_empty[b]);
        umma_commit(&w_empty[b]);
      }
      umma_commit(acc_full);
    }
  } else {
    // =========================== weight producer ===========================
    if (lane == 0) {
      const uint32_t bytes = 2u * BN * 128u;
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(P.w_tc) + (size_t)blockIdx.y * (size_t)total * bytes;
      for (int it = 0; it < total; ++it) {
        const int s = it & 1, n = it >> 1;
        if (n >= 1) mbar_wait(&w_empty[s], (uint32_t)((n - 1) & 1));
        mbar_arrive_expect_tx(&w_full[s], bytes);
        bulk_g2s(smem + S.w[s], wsrc + (size_t)it * bytes, bytes, &w_full[s]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(BN < 32 ? 32 : BN)) : "memory");
  }
}

}  // namespace

// ---------------------------------------------------------------- host side
static int tcgen_pick_bn(int cout) { return cout > 64 ? 128 : (cout > 32 ? 64 : 32); }

// Build the tensor-core weight image from the FMA-layout host array h[tap][cin_pad][cout_pad]:
//   [co-tile][chunk][tap][hi | lo][BN rows (co) x 32 ci], each [BN][128 B] block in SWIZZLE_128B order.
static void build_tc_image(const PackedConv& pc, const std::vector<float>& h, int BN, DevBuf& dst);

void pack_tc_weights(PackedConv& pc, const std::vector<float>& h) {
  build_tc_image(pc, h, tcgen_pick_bn(pc.Cout), pc.w_tc);
  pc.tc_bn = tcgen_pick_bn(pc.Cout);
  pc.tc_chunks = cdiv(pc.Cin, TC_KCH);
  // wide layers also get a BN=256 image: half the activation-operand bytes per FLOP (tcconv2 picks it
  // when the tile fits the shared-memory budget)
  if (pc.Cout % 256 == 0) build_tc_image(pc, h, 256, pc.w_tc256);
  pack_h_weights(pc, h);   // fp16 hi/lo image (tcconv5.cu)
}

static void build_tc_image(const PackedConv& pc, const std::vector<float>& h, int BN, DevBuf& dst) {
  const int nct = cdiv(pc.Cout, BN), nch = cdiv(pc.Cin, TC_KCH), nt = pc.ntaps;
  const size_t blk = (size_t)BN * 32;  // floats per hi (or lo) block
  std::vector<float> img((size_t)nct * nch * nt * 2 * blk, 0.f);
  for (int ct = 0; ct < nct; ++ct)
    for (int c = 0; c < nch; ++c)
      for (int t = 0; t < nt; ++t) {
        float* hi = &img[((((size_t)ct * nch + c) * nt + t) * 2) * blk];
        float* lo = hi + blk;
        for (int j = 0; j < BN; ++j) {
          const int co = ct * BN + j;
          if (co >= pc.Cout) continue;
          for (int k = 0; k < 32; ++k) {
            const int ci = c * TC_KCH + k;
            if (ci >= pc.Cin) continue;
            const float w = h[((size_t)t * pc.cin_pad + ci) * pc.cout_pad + co];
            uint32_t u;
            memcpy(&u, &w, 4);
            u &= 0xffffe000u;
            float wh;
            memcpy(&wh, &u, 4);
            const size_t off = (size_t)j * 32 + (size_t)(((k >> 2) ^ (j & 7)) << 2) + (k & 3);
            hi[off] = wh;
            lo[off] = w - wh;
          }
        }
      }
  dst.upload(img);
}

static int g_tc_version = -1;  // -1: AGPT_TC_V or the default; 1 = per-tap tiles, 2 = shifted descriptors, 3/4 = persistent
void tc_set_version(int v) { g_tc_version = v; }
static int g_tc_enabled = -1;   // -1: read AGPT_TENSOR_CORES from the environment on first use
void tc_set_enabled(int on) { g_tc_enabled = on != 0 ? 1 : 0; }
bool tc_enabled() {
  if (g_tc_enabled < 0) {
    const char* e = getenv("AGPT_TENSOR_CORES");
    g_tc_enabled = (e && e[0] == '0') ? 0 : 1;   // default ON (validated on B200: tests/test_*_gpu.py)
  }
  return g_tc_enabled == 1;
}

bool tcconv_supported(const TapConvParams& P) {
  if (!tc_enabled() || !P.w_tc || P.tc_bn == 0) return false;
  if (P.in_pitch % 4 != 0 || P.Cin % 4 != 0) return false;        // 16-byte cp.async granularity
  if ((reinterpret_cast<uintptr_t>(P.in) & 15) != 0 || (P.in_gstride % 4) != 0) return false;
  return true;
}

static int g_tc_flags_env = 0;
int tc_get_version() {
  int& ver = g_tc_version;
  if (ver < 0) {
    const char* e = getenv("AGPT_TC_V");
    ver = e ? atoi(e) : 6;   // default: v6 (persistent fp16 hi/lo, TMA epilogue) with v5 for single-wave grids; 5 = v5 only;
                             // 2 = the tf32 hi/lo kernel; 3/4 = persistent tf32 (experimental); 7 = v6 forced
  }
  static bool env_done = false;
  if (!env_done) {
    env_done = true;
    const char* b = getenv("AGPT_TC_BO");
    g_tc_flags_env = (b && b[0] == '1') ? 1 : 0;
    const char* d = getenv("AGPT_TC_DBGFLAGS");     // experiment switches (bits 2..): see tcconv2.cu
    if (d) g_tc_flags_env |= atoi(d) & ~3;
  }
  return ver;
}

void tcconv_launch(TapConvParams P, cudaStream_t st) {
  const int ver = tc_get_version();
  const int bo = g_tc_flags_env;
  P.tc_flags = bo | P.tc_flags_user;
  // v3 (persistent, overlapped epilogue) wins when an activation tile is reused by many taps (k >= 5:
  // measured +10..60 % on the k=7/11 HiFi-GAN convs); for k <= 3 and 1-tap GEMM-like layers the
  // concurrent transform/epilogue starve on shared-memory bandwidth and v2 is faster
  // (profiles/r1b_conv_microbench.txt).  AGPT_TC_V=3x forces v3 everywhere, =2 disables it.
  if ((ver == 6 || ver == 7) && tcconv6_launch(P, st, ver == 7)) return;   // persistent fp16 (7: also for single-wave grids)
  if (ver >= 5 && tcconv5_launch(P, st)) return;
  if ((ver == 3 && P.ntaps >= 5) || ver == 4) {
    if (tcconv3_launch(P, st)) return;
  }   // persistent, overlapped epilogue
  if (ver >= 2 && tcconv2_launch(P, st)) return;
  int lo = P.tap_off[0], hi = P.tap_off[0];
  for (int t = 1; t < P.ntaps; ++t) { lo = std::min(lo, P.tap_off[t]); hi = std::max(hi, P.tap_off[t]); }
  P.lo_al = lo;
  const int RR = round_up(TC_ROWS + (hi - lo), 8);
  P.R = RR;
  const int BN = P.tc_bn;
  const TcSmem S = tc_layout(BN, RR);
  const size_t smem = (size_t)S.total + 1024;
  AGPT_CHECK(smem <= 227 * 1024, "tcconv: shared memory (image too wide for the halo tile)");
  const int Wv = P.Wreal > 0 ? P.Wreal + 1 : 0;
  const int Lv = Wv ? (P.L / P.Wreal) * Wv : P.L;
  dim3 grid(cdiv(Lv, TC_ROWS), cdiv(P.Cout, BN), P.G);
  int dev = 0;
  AGPT_CUDA(cudaGetDevice(&dev));
  static bool attr_done_dev[64] = {false};
  if (!attr_done_dev[dev & 63]) {
    AGPT_CUDA(cudaFuncSetAttribute(tcconv_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AGPT_CUDA(cudaFuncSetAttribute(tcconv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AGPT_CUDA(cudaFuncSetAttribute(tcconv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done_dev[dev & 63] = true;
  }
  if (BN == 128) tcconv_kernel<128><<<grid, TC_THREADS, smem, st>>>(P);
  else if (BN == 64) tcconv_kernel<64><<<grid, TC_THREADS, smem, st>>>(P);
  else tcconv_kernel<32><<<grid, TC_THREADS, smem, st>>>(P);
}

}  // namespace agpt
