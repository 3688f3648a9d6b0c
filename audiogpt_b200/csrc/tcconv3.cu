// tcconv v3: persistent tcgen05 tapconv with the epilogue overlapped on the next tile's main loop.
//
// Same arithmetic as tcconv2.cu (K-major SWIZZLE_128B operand tiles, conv taps as row-shifted UMMA
// descriptors, 3xTF32 error compensation, accumulators in TMEM), but:
//   * one CTA per SM loops over output tiles (tile = 128 rows x BN channels of one sample);
//   * TWO accumulators live in TMEM (2 x BN columns): while the 4 epilogue warps drain tile i
//     (tcgen05.ld -> swizzled staging -> coalesced fused epilogue), the MMA warp already runs tile i+1;
//   * the transform warps prefetch the next activation chunk (also across tile boundaries) into
//     registers with 128-bit global loads, so no raw staging buffer is needed and the weight ring
//     gets the shared memory;
//   * all pipelines (activation buffers, weight stages, accumulators) are mbarrier rings whose phase
//     comes from counters that run across tiles.
// Warp roles (320 threads): 0-3 transform, 4 MMA issuer, 5 weight producer (cp.async.bulk), 6-9 epilogue.
#include "tapconv.cuh"
#include "tapconv_epi.cuh"
#include "tc_common.cuh"
#include "models.h"

namespace agpt {

namespace {

constexpr int V3_THREADS = 320;
constexpr int MAX_NA3 = 3, MAX_NW3 = 6, MAX_ITEMS = 20;      // 20 items/thread => up to 320 staged rows
constexpr int kMaxDyn3 = 227 * 1024 - 512;

struct Tc3Smem {
  uint32_t a_hi[MAX_NA3], a_lo[MAX_NA3], w[MAX_NW3], stg, rowinfo, rowp, bars, tmem_slot, total;
};
__host__ __device__ inline void tc3_layout(Tc3Smem& s, int BN, int RRA, int NA, int NW) {
  uint32_t o = 0;
  for (int i = 0; i < MAX_NA3; ++i) { s.a_hi[i] = o; if (i < NA) o += RRA * 128; }
  for (int i = 0; i < MAX_NA3; ++i) { s.a_lo[i] = o; if (i < NA) o += RRA * 128; }
  for (int i = 0; i < MAX_NW3; ++i) { s.w[i] = o; if (i < NW) o += 2 * BN * 128; }
  s.stg = o; o += 2 * TC_ROWS * 128;
  s.rowinfo = o; o += 4 * RRA * 4;     // ring of 4 tiles
  s.rowp = o; o += 2 * TC_ROWS * 4;    // ring of 2 tiles
  o = (o + 15) & ~15u;
  s.bars = o; o += 32 * 8;
  s.tmem_slot = o; o += 16;
  s.total = o;
}

__device__ __forceinline__ float4 pro_apply3(const TapConvParams& P, float4 v, bool ok, const float* pv) {
  if (P.pro == PRO_LRELU) {
    v.x = lrelu(v.x, P.slope); v.y = lrelu(v.y, P.slope); v.z = lrelu(v.z, P.slope); v.w = lrelu(v.w, P.slope);
  } else if (P.pro == PRO_ADDVEC) {
    if (ok) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(pv));
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
  } else if (P.pro == PRO_SILU) {
    v.x = siluf_(v.x); v.y = siluf_(v.y); v.z = siluf_(v.z); v.w = siluf_(v.w);
  }
  return v;
}
__device__ __forceinline__ float tf32_hi3(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

struct TileId { int g, q0, ct; };
__device__ __forceinline__ TileId tile_of(int t, int nct, int nrt) {
  TileId r;
  r.ct = t % nct;
  const int u = t / nct;
  r.q0 = (u % nrt) * TC_ROWS;
  r.g = u / nrt;
  return r;
}

template <int BN>
__global__ void __launch_bounds__(V3_THREADS, 1) tcconv3_kernel(const __grid_constant__ TapConvParams P) {
  extern __shared__ uint8_t smem_raw_[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~(uintptr_t)1023);
  const int RRA = P.R, NA = P.tc_na, NW = P.tc_nw;
  __shared__ Tc3Smem S;
  if (threadIdx.x == 0) tc3_layout(S, BN, RRA, NA, NW);
  __syncthreads();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S.bars);
  uint64_t* a_full = bars + 0;             // [MAX_NA3]
  uint64_t* a_empty = bars + MAX_NA3;      // [MAX_NA3]
  uint64_t* w_full = bars + 2 * MAX_NA3;   // [MAX_NW3]
  uint64_t* w_empty = w_full + MAX_NW3;    // [MAX_NW3]
  uint64_t* acc_full = w_empty + MAX_NW3;  // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + S.tmem_slot);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Wv = P.Wreal > 0 ? P.Wreal + 1 : 0;
  const int Lv = Wv ? (P.L / P.Wreal) * Wv : P.L;
  const int nchunks = P.tc_chunks, ntaps = P.ntaps, iters_per_tile = nchunks * ntaps;
  const int lo = P.lo_al;
  const int nct = (P.Cout + BN - 1) / BN, nrt = (Lv + TC_ROWS - 1) / TC_ROWS;
  const int ntiles = nct * nrt * P.G;
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;

  if (tid == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void*)tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // optional per-CTA wait accounting (tc_flags & 2): [0] total, [1] MMA wait a_full, [2] MMA wait w_full,
  // [3] MMA wait acc_empty, [4] transform wait a_empty, [5] epilogue wait acc_full, [6] epilogue busy, [7] producer wait
  const bool dbg_on = (P.tc_flags & 2) && P.dbg;
  long long* dbg = dbg_on ? P.dbg + 8 * (long)blockIdx.x : nullptr;
  const long long t_begin = dbg_on ? clock64() : 0;
#define DBG_WAIT(slot, stmt) do { if (dbg_on) { const long long _t = clock64(); stmt; dbgacc[slot] += clock64() - _t; } else { stmt; } } while (0)
  long long dbgacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  if (warp < 4) {
    // =========================== transform warps ===========================
    int* rowinfo_ring = reinterpret_cast<int*>(smem + S.rowinfo);
    const int items = RRA * 8;
    const int nper = (items + 127) / 128;        // <= MAX_ITEMS (host-checked)
    float4 v[MAX_ITEMS];
    uint32_t okmask = 0;
    const int jcol = tid & 7, r0 = tid >> 3;
    const int total_gc = my_tiles * nchunks;

    // prefetch global chunk k (tile k / nchunks, chunk k % nchunks) into registers
    auto prefetch = [&](int k) {
      const int tl = k / nchunks, c = k - tl * nchunks;
      const TileId T = tile_of((int)blockIdx.x + tl * (int)gridDim.x, nct, nrt);
      int* rowinfo = rowinfo_ring + (tl & 3) * RRA;
      if (c == 0) {
        for (int i = tid; i < RRA; i += 128) {
          const int q = T.q0 + lo + i;
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
        named_bar_sync(1, 128);
      }
      const float* __restrict__ ing = P.in + T.g * P.in_gstride;
      // each thread owns chunk column j of rows r0, r0+16, ...: branch-free, table read with an explicit
      // shared-space load, so that all 128-bit global loads of the chunk are in flight together
      const uint32_t ri_sh = smem_u32(rowinfo);
      const int ch = c * TC_KCH + 4 * jcol;
      const bool chok = ch < P.Cin;
      okmask = 0;
#pragma unroll
      for (int u = 0; u < MAX_ITEMS; ++u) {
        if (u < nper) {
          const int row = r0 + 16 * u;
          int a = -1;
          if (row < RRA) asm volatile("ld.shared.s32 %0, [%1];" : "=r"(a) : "r"(ri_sh + 4u * (uint32_t)row));
          const bool ok = chok && (a >= 0);
          const float4* src = reinterpret_cast<const float4*>(ok ? (ing + a + ch) : P.in);
          if (!(P.tc_flags & 16)) v[u] = __ldg(src);   // (experiment bit 16: no global reads) zero-fill select happens at use
          okmask |= (ok ? 1u : 0u) << u;
        }
      }
    };

    if (total_gc > 0) prefetch(0);
    for (int k = 0; k < total_gc; ++k) {
      const int buf = k % NA, n = k / NA;
      const int tl = k / nchunks, c = k - tl * nchunks;
      if (n >= 1) DBG_WAIT(4, mbar_wait(&a_empty[buf], (uint32_t)((n - 1) & 1)));
      uint8_t* ahi = smem + S.a_hi[buf];
      uint8_t* alo = smem + S.a_lo[buf];
      const float* pvg = nullptr;
      if (P.pro == PRO_ADDVEC) {
        const TileId T = tile_of((int)blockIdx.x + tl * (int)gridDim.x, nct, nrt);
        pvg = P.pvec + (long)T.g * P.pvec_gstride + c * TC_KCH;
      }
#pragma unroll
      for (int u = 0; u < MAX_ITEMS; ++u) {
        if (u < nper) {
          const int row = r0 + 16 * u;
          if (row < RRA) {
            const bool okv = (okmask >> u) & 1u;
            const float4 vin = okv ? v[u] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 x = pro_apply3(P, vin, okv, pvg ? pvg + 4 * jcol : nullptr);
            const float4 h = make_float4(tf32_hi3(x.x), tf32_hi3(x.y), tf32_hi3(x.z), tf32_hi3(x.w));
            const float4 l = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
            const uint32_t o = sw128(row, jcol);
            *reinterpret_cast<float4*>(ahi + o) = h;
            *reinterpret_cast<float4*>(alo + o) = l;
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(&a_full[buf]);
      if (k + 1 < total_gc) prefetch(k + 1);     // in flight while this thread waits for the next a_empty
    }
  } else if (warp == 4) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
      int gc = 0, it = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int acc = tl & 1, na = tl >> 1;
        if (na >= 1) { DBG_WAIT(3, mbar_wait(&acc_empty[acc], (uint32_t)((na - 1) & 1))); tc_fence_after(); }
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        bool first = true;
        for (int c = 0; c < nchunks; ++c, ++gc) {
          const int buf = gc % NA;
          DBG_WAIT(1, mbar_wait(&a_full[buf], (uint32_t)((gc / NA) & 1)));
          tc_fence_after();
          const uint32_t ahi0 = smem_u32(smem + S.a_hi[buf]), alo0 = smem_u32(smem + S.a_lo[buf]);
          for (int t = 0; t < ntaps; ++t, ++it) {
            const int s = it % NW;
            DBG_WAIT(2, mbar_wait(&w_full[s], (uint32_t)((it / NW) & 1)));
            tc_fence_after();
            const uint32_t shift = (uint32_t)(P.tap_off[t] - lo) * 128u;
            const uint64_t dah = make_desc(ahi0 + shift), dal = make_desc(alo0 + shift);
            const uint64_t dwh = make_desc(smem_u32(smem + S.w[s]));
            const uint64_t dwl = make_desc(smem_u32(smem + S.w[s] + BN * 128));
#pragma unroll
            for (int k = 0; k < TC_KCH / 8; ++k) {
              const uint64_t ko = (uint64_t)((k * 32) >> 4);
              umma_tf32(tmem_d, dah + ko, dwh + ko, idesc, first ? 0u : 1u);
              first = false;
              umma_tf32(tmem_d, dal + ko, dwh + ko, idesc, 1u);
              umma_tf32(tmem_d, dah + ko, dwl + ko, idesc, 1u);
            }
            umma_commit(&w_empty[s]);
          }
          umma_commit(&a_empty[buf]);
        }
        umma_commit(&acc_full[acc]);
      }
    }
  } else if (warp == 5) {
    // =========================== weight producer ===========================
    if (lane == 0) {
      const uint32_t bytes = 2u * BN * 128u;
      int it = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const TileId T = tile_of((int)blockIdx.x + tl * (int)gridDim.x, nct, nrt);
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(P.w_tc) + (size_t)T.ct * (size_t)iters_per_tile * bytes;
        for (int i = 0; i < iters_per_tile; ++i, ++it) {
          const int s = it % NW, n = it / NW;
          if (n >= 1) DBG_WAIT(7, mbar_wait(&w_empty[s], (uint32_t)((n - 1) & 1)));
          mbar_arrive_expect_tx(&w_full[s], bytes);
          bulk_g2s(smem + S.w[s], wsrc + (size_t)i * bytes, bytes, &w_full[s]);
        }
      }
    }
  } else {
    // =========================== epilogue warps (6..9) ===========================
    const int et = tid - 6 * 32;                 // 0..127
    const int quad = warp & 3;                   // TMEM lane quadrant this warp may access
    const int myrow = quad * 32 + lane;          // accumulator row held by this thread's TMEM lane
    int* rowp_ring = reinterpret_cast<int*>(smem + S.rowp);
    uint8_t* stg0 = smem + S.stg;
    for (int tl = 0; tl < my_tiles; ++tl) {
      const TileId T = tile_of((int)blockIdx.x + tl * (int)gridDim.x, nct, nrt);
      const int acc = tl & 1;
      int* rowp = rowp_ring + acc * TC_ROWS;
      {
        const int q = T.q0 + et;
        int p = -1;
        if (q < Lv) {
          if (Wv) {
            const int h = q / Wv, w = q - h * Wv;
            if (w < P.Wreal) p = h * P.Wreal + w;
          } else {
            p = q;
          }
        }
        rowp[et] = p;
      }
      named_bar_sync(2, 128);
      const int co0 = T.ct * BN;
      EpiPre pre[8];
      int pp[8];
      auto load_block = [&](int cb) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = et + i * 128;
          pp[i] = rowp[idx >> 3];
          if (P.tc_flags & 4) { pre[i].a = make_float4(0.f, 0.f, 0.f, 0.f); pre[i].b = pre[i].a; continue; }   // experiment
          if (pp[i] >= 0) epi_load(P, T.g, pp[i], co0 + cb + 4 * (idx & 7), pre[i]);
        }
      };
      load_block(0);                              // global reads in flight while the tile is still accumulating
      DBG_WAIT(5, mbar_wait(&acc_full[acc], (uint32_t)((tl >> 1) & 1)));
      tc_fence_after();
      const long long t_epi0 = dbg_on ? clock64() : 0;
#pragma unroll 1
      for (int cb = 0, blk = 0; cb < BN; cb += 32, ++blk) {
        uint32_t rg[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + cb);
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
        if (cb + 32 >= BN) {                      // last TMEM read of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(&acc_empty[acc]);
        }
        uint8_t* stg = stg0 + (blk & 1) * (TC_ROWS * 128);
#pragma unroll
        for (int qd = 0; qd < 8; ++qd)
          *reinterpret_cast<float4*>(stg + sw128(myrow, qd)) =
              make_float4(__uint_as_float(rg[4 * qd]), __uint_as_float(rg[4 * qd + 1]),
                          __uint_as_float(rg[4 * qd + 2]), __uint_as_float(rg[4 * qd + 3]));
        named_bar_sync(2, 128);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = et + i * 128;
          const int row = idx >> 3, j = idx & 7;
          if (pp[i] >= 0 && !((P.tc_flags & 8) && (row & 63) != 0))   // experiment bit 8: store 2 rows only
            epi_store(P, T.g, pp[i], co0 + cb + 4 * j, *reinterpret_cast<const float4*>(stg + sw128(row, j)), pre[i]);
        }
        if (cb + 32 < BN) load_block(cb + 32);
      }
      if (dbg_on) dbgacc[6] += clock64() - t_epi0;
      // staging halves alternate per block; with an odd block count (BN=32) the same half would be
      // reused by the next tile without an intervening barrier -> the tile-start barrier covers it
    }
  }

  if (dbg_on) {
    if (tid == 0) { dbg[4] = dbgacc[4]; }
    if (warp == 4 && lane == 0) { dbg[1] = dbgacc[1]; dbg[2] = dbgacc[2]; dbg[3] = dbgacc[3]; }
    if (warp == 5 && lane == 0) dbg[7] = dbgacc[7];
    if (tid == 6 * 32) { dbg[5] = dbgacc[5]; dbg[6] = dbgacc[6]; }
  }
  tc_fence_before();
  __syncthreads();
  if (dbg_on && tid == 0) dbg[0] = clock64() - t_begin;
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace

bool tcconv3_launch(TapConvParams P, cudaStream_t st) {
  int lo = P.tap_off[0], hi = P.tap_off[0];
  for (int t = 1; t < P.ntaps; ++t) { lo = std::min(lo, P.tap_off[t]); hi = std::max(hi, P.tap_off[t]); }
  P.lo_al = lo;
  const int RRA = round_up(TC_ROWS + (hi - lo), 8);
  P.R = RRA;
  if ((RRA * 8 + 127) / 128 > MAX_ITEMS) return false;
  const int BN = P.tc_bn;
  const long avail = (long)kMaxDyn3 - 1024 - (2 * TC_ROWS * 128) /*staging*/ - (4 * RRA * 4 + 2 * TC_ROWS * 4 + 512);
  const long abytes = 2L * RRA * 128, wbytes = 2L * BN * 128;
  const int iters = P.tc_chunks * P.ntaps;
  int NA = 2;
  if (NA * abytes + 2 * wbytes > avail) NA = 1;
  if (NA * abytes + 2 * wbytes > avail) return false;
  int NW = (int)std::min<long>(MAX_NW3, (avail - NA * abytes) / wbytes);
  if (NA == 2 && P.ntaps == 1 && NW >= 5 && 3 * abytes + 3 * wbytes <= avail) {   // GEMM-like: a third activation buffer
    NA = 3;
    NW = (int)std::min<long>(MAX_NW3, (avail - NA * abytes) / wbytes);
  }
  NW = std::max(2, NW);
  P.tc_na = NA; P.tc_nw = NW;
  Tc3Smem S;
  tc3_layout(S, BN, RRA, NA, NW);
  const size_t smem = (size_t)S.total + 1024;
  if (smem > (size_t)kMaxDyn3) return false;
  const int Wv = P.Wreal > 0 ? P.Wreal + 1 : 0;
  const int Lv = Wv ? (P.L / P.Wreal) * Wv : P.L;
  const int ntiles = cdiv(Lv, TC_ROWS) * cdiv(P.Cout, BN) * P.G;
  int dev = 0;
  AGPT_CUDA(cudaGetDevice(&dev));
  static bool attr_done_dev[64] = {false};
  static int sms_dev[64] = {0};
  if (!attr_done_dev[dev & 63]) {
    AGPT_CUDA(cudaFuncSetAttribute(tcconv3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn3));
    AGPT_CUDA(cudaFuncSetAttribute(tcconv3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn3));
    AGPT_CUDA(cudaFuncSetAttribute(tcconv3_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn3));
    AGPT_CUDA(cudaDeviceGetAttribute(&sms_dev[dev & 63], cudaDevAttrMultiProcessorCount, dev));
    attr_done_dev[dev & 63] = true;
  }
  (void)iters;
  const int grid = std::min(ntiles, sms_dev[dev & 63]);
  if (BN == 128) tcconv3_kernel<128><<<grid, V3_THREADS, smem, st>>>(P);
  else if (BN == 64) tcconv3_kernel<64><<<grid, V3_THREADS, smem, st>>>(P);
  else tcconv3_kernel<32><<<grid, V3_THREADS, smem, st>>>(P);
  return true;
}

}  // namespace agpt
