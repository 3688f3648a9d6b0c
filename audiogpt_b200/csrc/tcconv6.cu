// tcconv v6: PERSISTENT fp16-hi/lo tap-GEMM -- load / transform, tensor-core main loop and fused epilogue
// of different output tiles run concurrently on one SM.
//
// Arithmetic and operand layout are those of tcconv5.cu (fp16 hi/lo parts in K-major SWIZZLE_128B tiles,
// 64-channel chunks, a conv tap = a row-shifted UMMA descriptor, three products accumulated in fp32
// in TMEM, weights pre-scaled by a power of two).  What changes is the schedule: v5 runs one tile per
// CTA and its phases are serial (first activation tile ~7 k cycles, main loop, epilogue 5-25 k cycles;
// profiles/r1c_conv_microbench.txt: the tensor pipe is busy only 35-55 % of a tile), and all 148 CTAs
// of a wave hit HBM at the same time.  Here one CTA per SM loops over tiles with dedicated warps:
//
//   warps 0-3   epilogue: TMEM lane quadrant w -> registers -> warp-private swizzled staging rows ->
//               coalesced fused epilogue (residual / accumulate reads issued one 32-column block ahead)
//   warp  4     MMA issuer (one thread): TWO accumulators in TMEM (2 x BN columns); tile i+1 starts
//               as soon as its first operand chunk is ready, while tile i is being drained
//   warp  5     weight producer (cp.async.bulk ring, runs ahead across tiles)
//   warps 6-13  transform: 128-bit global loads of the NEXT activation chunk (also across tile
//               boundaries) into registers -> prologue -> fp16 hi/lo split -> operand ring (NA buffers)
//
// All pipelines are mbarrier rings whose phases come from counters that run across tiles.
#include "tapconv.cuh"
#include "tapconv_epi.cuh"
#include "tc_common.cuh"
#include "tc_h16.cuh"
#include "tc_tma.cuh"
#include "models.h"

namespace agpt {

namespace {

constexpr int V6_THREADS = 448;
constexpr int V6_NT = 256;                    // transform threads (warps 6-13)
constexpr int MAX_NA6 = 3, MAX_NW6 = 6;
constexpr int kMaxDyn6 = 227 * 1024 - 512;

struct Tc6Smem {
  uint32_t a_hi[MAX_NA6], a_lo[MAX_NA6], w[MAX_NW6], stg, cvs, rowinfo, rowp, bars, tmem_slot, total;
};
constexpr int V6_EBLK = TC_ROWS * 128;   // one epilogue block buffer: 128 rows x 32 fp32 columns (4 KB per epilogue warp)
constexpr uint32_t V6_FLAG_WBLK = 2048;  // tc_flags bit: TMA epilogue with whole-block [128 x 32] boxes
constexpr uint32_t V6_FLAG_TMA = 64;     // tc_flags bit: epilogue through tensor maps (TMA load / store / reduce-add)
__host__ __device__ inline void tc6_layout(Tc6Smem& s, int BN, int RRA, int NA, int NW, int NB, int tps) {
  uint32_t o = 0;
  for (int i = 0; i < MAX_NA6; ++i) { s.a_hi[i] = o; if (i < NA) o += RRA * 128; }
  for (int i = 0; i < MAX_NA6; ++i) { s.a_lo[i] = o; if (i < NA) o += RRA * 128; }
  for (int i = 0; i < MAX_NW6; ++i) { s.w[i] = o; if (i < NW) o += tps * 2 * BN * 128; }
  s.stg = o; o += (NB > 1 ? NB : 1) * V6_EBLK;   // epilogue block buffers (warp-private 32-row slices), 1 KB aligned
  s.cvs = o; o += 4 * 256 * 4;         // per-epilogue-warp copy of the tile's bias (+ per-sample vector)
  s.rowinfo = o; o += 4 * RRA * 4;     // ring of 4 tiles
  s.rowp = o; o += 2 * TC_ROWS * 4;    // ring of 2 tiles
  o = (o + 15) & ~15u;
  s.bars = o; o += 48 * 8;
  s.tmem_slot = o; o += 16;
  s.total = o;
}

struct TileId6 { int g, q0, ct; };
__device__ __forceinline__ TileId6 tile_of6(int t, int nct, int nrt) {
  TileId6 r;
  r.ct = t % nct;
  const int u = t / nct;
  r.q0 = (u % nrt) * TC_ROWS;
  r.g = u / nrt;
  return r;
}

// NI = register-prefetched 8-channel items per transform thread (rows r0, r0+32, ...): covers RRA <= 32 * NI
// accumulate 32 more fp32 TMEM columns of this lane into rg (the A_hi x W_lo range of a stacked accumulator)
__device__ __forceinline__ void tc6_ld32_add(uint32_t taddr, uint32_t* rg) {
  uint32_t r2[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3]), "=r"(r2[4]), "=r"(r2[5]), "=r"(r2[6]), "=r"(r2[7]),
        "=r"(r2[8]), "=r"(r2[9]), "=r"(r2[10]), "=r"(r2[11]), "=r"(r2[12]), "=r"(r2[13]), "=r"(r2[14]), "=r"(r2[15]),
        "=r"(r2[16]), "=r"(r2[17]), "=r"(r2[18]), "=r"(r2[19]), "=r"(r2[20]), "=r"(r2[21]), "=r"(r2[22]), "=r"(r2[23]),
        "=r"(r2[24]), "=r"(r2[25]), "=r"(r2[26]), "=r"(r2[27]), "=r"(r2[28]), "=r"(r2[29]), "=r"(r2[30]), "=r"(r2[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) rg[i] = __float_as_uint(__uint_as_float(rg[i]) + __uint_as_float(r2[i]));
}


template <int BN, int NI, bool NARROW>
__global__ void __launch_bounds__(V6_THREADS, 1) tcconv6_kernel(const __grid_constant__ TapConvParams P,
                                                                const __grid_constant__ CUtensorMap tm_res,
                                                                const __grid_constant__ CUtensorMap tm_out) {
  extern __shared__ uint8_t smem_raw_[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~(uintptr_t)1023);
  const int RRA = P.R, NA = P.tc_na, NW = P.tc_nw;
  __shared__ Tc6Smem S;
  const int tps = P.tc_tps;                  // taps per weight stage
  if (threadIdx.x == 0) tc6_layout(S, BN, RRA, NA, NW, P.tc_nb, tps);
  __syncthreads();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S.bars);
  uint64_t* a_full = bars + 0;             // [MAX_NA6]
  uint64_t* a_empty = bars + MAX_NA6;      // [MAX_NA6]
  uint64_t* w_full = bars + 2 * MAX_NA6;   // [MAX_NW6]
  uint64_t* w_empty = w_full + MAX_NW6;    // [MAX_NW6]
  uint64_t* acc_full = w_empty + MAX_NW6;  // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint64_t* e_full = acc_empty + 2;        // [4 epilogue warps][4 block buffers] (TMA epilogue)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + S.tmem_slot);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Wv = P.Wreal > 0 ? P.Wreal + 1 : 0;
  const int Lv = Wv ? (P.L / P.Wreal) * Wv : P.L;
  const int nchunks = P.tc_chunks_h, ntaps = P.ntaps, iters_per_tile = nchunks * ntaps;
  const int lo = P.lo_al;
  const int nct = (P.Cout + BN - 1) / BN, nrt = (Lv + TC_ROWS - 1) / TC_ROWS;
  const int ntiles = nct * nrt * P.G;
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  // Two accumulators.  Tiles up to 128 columns use STACKED weight parts: one MMA of width 2 BN reads the hi and lo
  // blocks of a weight stage as one tile and leaves [A_hi W_hi | A_hi W_lo] in two column ranges, a second of
  // width BN adds A_lo W_hi -- 2 instructions / 20 KB of operand reads per k-step instead of 3 / 24 KB; the epilogue
  // adds the ranges (tcconv5.cu has the measurement behind it).
  constexpr bool STK = BN <= 128;
  constexpr uint32_t ACCW = STK ? 2 * BN : BN;                   // accumulator stride in TMEM columns
  constexpr uint32_t TMEM_COLS = (2 * ACCW < 32) ? 32 : 2 * ACCW;
  const bool stk = STK && !(P.tc_flags & 4);

  if (tid == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], V6_NT); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
    for (int i = 0; i < 16; ++i) mbar_init(&e_full[i], 1);
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
  pdl_wait();          // everything above overlaps the previous kernel's tail
  // optional per-CTA wait accounting (tc_flags & 2): [0] total, [1] MMA wait a_full, [2] MMA wait w_full,
  // [3] MMA wait acc_empty, [4] transform wait a_empty, [5] epilogue wait acc_full, [6] epilogue busy, [7] producer wait
  const bool dbg_on = (P.tc_flags & 2) && P.dbg;
  long long* dbg = dbg_on ? P.dbg + 8 * (long)blockIdx.x : nullptr;
  const long long t_begin = dbg_on ? clock64() : 0;
#define DBG_WAIT6(slot, stmt) do { if (dbg_on) { const long long _t = clock64(); stmt; dbgacc[slot] += clock64() - _t; } else { stmt; } } while (0)
  long long dbgacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  if (warp >= 6) {
    // =========================== transform warps ===========================
    // Thread mapping: QN 16-byte fp16 chunks (8 channels each) per row, 256 / QN rows per pass.  NARROW
    // (Cin <= 32: only 4 chunks per row are ever touched) uses all threads on 64 rows per pass and keeps
    // TWO register sets in flight (chunks k+1 and k+2), because with one tile = one chunk the single-depth
    // prefetch leaves the global-load latency exposed once per tile.
    constexpr int QN = NARROW ? 4 : 8, RSTR = V6_NT / QN;
    const int xt = tid - 6 * 32;                   // 0..255
    int* rowinfo_ring = reinterpret_cast<int*>(smem + S.rowinfo);
    struct RSet { float4 v0[NI], v1[NI]; uint32_t ok0, ok1; };
    RSet RA, RB;
    const int q = xt % QN, r0 = xt / QN;
    const int total_gc = my_tiles * nchunks;

    // prefetch global chunk k (tile k / nchunks, chunk k % nchunks) into a register set
    auto prefetch = [&](int k, RSet& R) {
      const int tl = k / nchunks, c = k - tl * nchunks;
      const TileId6 T = tile_of6((int)blockIdx.x + tl * (int)gridDim.x, nct, nrt);
      int* rowinfo = rowinfo_ring + (tl & 3) * RRA;
      if (c == 0 && Wv != 0) {
        // (2-D layers only: a 1-D layer computes the row offset directly)
        // ring of 4 tiles: a slot is rewritten three tiles later; every thread passes this barrier only
        // after its own reads of the older tables were issued
        for (int i = xt; i < RRA; i += V6_NT) {
          const int qq = T.q0 + lo + i;
          int a = -1;
          if (qq >= 0 && qq < Lv) {
            if (Wv) {
              const int h = qq / Wv, w = qq - h * Wv;
              if (w < P.Wreal) a = (h * P.Wreal + w) * P.in_pitch;
            } else {
              a = qq * P.in_pitch;
            }
          }
          rowinfo[i] = a;
        }
        named_bar_sync(1, V6_NT);
      }
      const float* __restrict__ ing = P.in + T.g * P.in_gstride;
      const uint32_t ri_sh = smem_u32(rowinfo);
      const int ch = c * H_KCH + 8 * q;
      const bool chok0 = ch < P.Cin, chok1 = ch + 4 < P.Cin;
      R.ok0 = 0; R.ok1 = 0;
#pragma unroll
      for (int u = 0; u < NI; ++u) {
        const int row = r0 + RSTR * u;
        int a = -1;
        if (Wv == 0) {
          const int qq = T.q0 + lo + row;
          if (row < RRA && qq >= 0 && qq < Lv) a = qq * P.in_pitch;
        } else if (row < RRA) {
          asm volatile("ld.shared.s32 %0, [%1];" : "=r"(a) : "r"(ri_sh + 4u * (uint32_t)row));
        }
        const bool k0 = chok0 && (a >= 0), k1 = chok1 && (a >= 0);
        R.v0[u] = ldg_stream(k0 ? (ing + a + ch) : P.in);      // zero-select happens at use
        R.v1[u] = ldg_stream(k1 ? (ing + a + ch + 4) : P.in);
        R.ok0 |= (k0 ? 1u : 0u) << u;
        R.ok1 |= (k1 ? 1u : 0u) << u;
      }
    };
    // prologue + fp16 hi/lo split of register set R -> operand buffer of global chunk k
    auto convert = [&](int k, const RSet& R) {
      const int buf = k % NA, n = k / NA;
      const int tl = k / nchunks, c = k - tl * nchunks;
      const int kv = min(H_KCH, P.Cin - c * H_KCH);
      const int nq = ((kv + 15) >> 4) << 1;          // 16-byte chunks the MMA k-steps of this chunk touch
      if (n >= 1) mbar_wait(&a_empty[buf], (uint32_t)((n - 1) & 1));
      uint8_t* ahi = smem + S.a_hi[buf];
      uint8_t* alo = smem + S.a_lo[buf];
      const float* pvg = nullptr;
      if (P.pro == PRO_ADDVEC) {
        const TileId6 T = tile_of6((int)blockIdx.x + tl * (int)gridDim.x, nct, nrt);
        pvg = P.pvec + (long)T.g * P.pvec_gstride + c * H_KCH + 8 * q;
      }
      if (q < nq) {
#pragma unroll
        for (int u = 0; u < NI; ++u) {
          const int row = r0 + RSTR * u;
          if (row < RRA) {
            const bool k0 = (R.ok0 >> u) & 1u, k1 = (R.ok1 >> u) & 1u;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 x0 = pro_apply5(P, k0 ? R.v0[u] : z, k0, pvg);
            const float4 x1 = pro_apply5(P, k1 ? R.v1[u] : z, k1, pvg ? pvg + 4 : nullptr);
            uint4 h, l;
            h.x = split2(x0.x, x0.y, l.x);
            h.y = split2(x0.z, x0.w, l.y);
            h.z = split2(x1.x, x1.y, l.z);
            h.w = split2(x1.z, x1.w, l.w);
            const uint32_t o = sw128(row, q);
            *reinterpret_cast<uint4*>(ahi + o) = h;
            *reinterpret_cast<uint4*>(alo + o) = l;
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(&a_full[buf]);
    };

    if (NARROW) {
      if (total_gc > 0) prefetch(0, RA);
      if (total_gc > 1) prefetch(1, RB);
      for (int k = 0; k < total_gc; k += 2) {
        convert(k, RA);
        if (k + 2 < total_gc) prefetch(k + 2, RA);
        if (k + 1 < total_gc) {
          convert(k + 1, RB);
          if (k + 3 < total_gc) prefetch(k + 3, RB);
        }
      }
    } else {
      if (total_gc > 0) prefetch(0, RA);
      for (int k = 0; k < total_gc; ++k) {
        convert(k, RA);
        if (k + 1 < total_gc) prefetch(k + 1, RA);     // in flight while this thread waits for the next a_empty
      }
    }
  } else if (warp == 4) {
    // =========================== MMA issuer ===========================
    // The whole warp runs the loop (converged waits); one ELECTED lane issues, so that ptxas keeps the
    // descriptors in uniform registers instead of a per-MMA divergence "waterfall".  A weight stage holds
    // `tps` taps (~32 KB), so the per-stage handshake (wait, commit, scalar bookkeeping -- a few hundred
    // cycles of dependent single-warp code) is paid once per 12-48 MMAs even on the narrow layers.
    {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | ((uint32_t)((STK ? 2 * BN : BN) >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
      const uint64_t DC = make_desc(0);                         // descriptor constants; the low 14 bits take (address >> 4)
      const uint32_t a16 = smem_u32(smem + S.a_hi[0]) >> 4;     // operand ring: hi tiles, then lo tiles
      const uint32_t abuf16 = (uint32_t)(RRA * 128) >> 4;       // one hi (or lo) tile
      const uint32_t alo16 = (uint32_t)NA * abuf16;             // hi -> lo distance
      const uint32_t w16 = smem_u32(smem + S.w[0]) >> 4;
      const uint32_t wtap16 = (uint32_t)(2 * BN * 128) >> 4;    // one tap (hi | lo) inside a stage
      const uint32_t wstage16 = (uint32_t)tps * wtap16;
      constexpr uint32_t wlo16 = (uint32_t)(BN * 128) >> 4;
      int gc = 0, it = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int acc = tl & 1, na = tl >> 1;
        if (na >= 1) { DBG_WAIT6(3, mbar_wait(&acc_empty[acc], (uint32_t)((na - 1) & 1))); tc_fence_after(); }
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * ACCW;
        uint32_t nz = 0;                                        // 0 for the very first MMA of the tile
        for (int c = 0; c < nchunks; ++c, ++gc) {
          const int buf = gc % NA;
          const int kv = min(H_KCH, P.Cin - c * H_KCH);
          const int ksteps = (kv + 15) >> 4;
          DBG_WAIT6(1, mbar_wait(&a_full[buf], (uint32_t)((gc / NA) & 1)));
          tc_fence_after();
          const uint64_t dA = DC + (uint64_t)(a16 + (uint32_t)buf * abuf16);
          for (int t0 = 0; t0 < ntaps; t0 += tps, ++it) {
            const int s = it % NW;
            const int t1 = min(ntaps, t0 + tps);
            DBG_WAIT6(2, mbar_wait(&w_full[s], (uint32_t)((it / NW) & 1)));
            tc_fence_after();
            const uint64_t dW = DC + (uint64_t)(w16 + (uint32_t)s * wstage16);
            if (elect_one()) {
              for (int t = t0; t < t1; ++t) {
                const uint64_t dah = dA + (uint64_t)((uint32_t)(P.tap_off[t] - lo) * 8u);   // one row = 128 B = 8 x 16 B
                const uint64_t dal = dah + alo16;
                const uint64_t dwh = dW + (uint64_t)((uint32_t)(t - t0) * wtap16);
                const uint64_t dwl = dwh + wlo16;
                for (int k = 0; k < ksteps; ++k) {
                  const uint64_t ko = (uint64_t)(2 * k);        // 32 bytes per k-step
                  if (stk) {
                    umma_f16(tmem_d, dah + ko, dwh + ko, idesc2, nz);      // [hi x hi | hi x lo]
                    nz = 1u;
                    umma_f16(tmem_d, dal + ko, dwh + ko, idesc, 1u);       // += lo x hi
                  } else {
                    umma_f16(tmem_d, dah + ko, dwh + ko, idesc, nz);
                    nz = 1u;
                    umma_f16(tmem_d, dal + ko, dwh + ko, idesc, 1u);
                    umma_f16(tmem_d, dah + ko, dwl + ko, idesc, 1u);
                  }
                }
              }
              umma_commit(&w_empty[s]);
              if (t1 == ntaps) {
                umma_commit(&a_empty[buf]);
                if (c == nchunks - 1) umma_commit(&acc_full[acc]);
              }
            }
            __syncwarp();
            nz = 1u;
          }
        }
      }
    }
  } else if (warp == 5) {
    // =========================== weight producer ===========================
    if (lane == 0) {
      const uint32_t tapbytes = 2u * BN * 128u;
      int it = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const TileId6 T = tile_of6((int)blockIdx.x + tl * (int)gridDim.x, nct, nrt);
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(P.w_h) + (size_t)T.ct * (size_t)iters_per_tile * tapbytes;
        for (int c = 0; c < nchunks; ++c) {
          for (int t0 = 0; t0 < ntaps; t0 += tps, ++it) {
            const int s = it % NW, n = it / NW;
            const uint32_t bytes = (uint32_t)(min(ntaps, t0 + tps) - t0) * tapbytes;
            if (n >= 1) DBG_WAIT6(7, mbar_wait(&w_empty[s], (uint32_t)((n - 1) & 1)));
            mbar_arrive_expect_tx(&w_full[s], bytes);
            bulk_g2s(smem + S.w[s], wsrc + (size_t)(c * ntaps + t0) * tapbytes, bytes, &w_full[s]);
          }
        }
      }
    }
  } else if (P.tc_flags & V6_FLAG_TMA) {
    // =========================== epilogue warps (0..3), TMA version ===========================
    // Every warp owns the 32 accumulator rows of its TMEM lane quadrant and works on [32 rows x 32 columns]
    // blocks held in 4 KB SWIZZLE_128B buffers: the residual block is TMA-loaded (LA blocks ahead, also
    // across tile boundaries), combined IN PLACE with the accumulator (lane = row) and TMA-stored
    // (reduce-add when the layer accumulates into its output).  No LSU global traffic, no cross-warp sync.
    const int quad = warp;
    const int NB = P.tc_nb;
    const int LA = NB - 2;                                  // residual look-ahead in blocks (one store may be in flight)
    const bool has_res = (P.epi == EPI_RES || P.epi == EPI_ACC) && P.res != nullptr;
    const bool red_add = (P.epi == EPI_ACC) && P.accumulate;
    // whole-block mode (default): ONE tensor-map op moves a [128 rows x 32 columns] block for all four warps
    // (4x fewer TMA ops; the warps meet at a named barrier before the store).  Per-warp mode: [32 x 32] boxes.
    const bool wb = (P.tc_flags & V6_FLAG_WBLK) != 0;
    const bool leader = wb ? (tid == 0) : (lane == 0);
    uint8_t* ebuf = smem + S.stg + quad * 4096;             // this warp's 32 rows inside block buffer 0
    uint64_t* efull = wb ? e_full : e_full + quad * 4;
    float* cvs = reinterpret_cast<float*>(smem + S.cvs) + quad * 256;
    const float dsc = P.tc_descale;
    constexpr int nblk = BN / 32;
    const int total_blk = my_tiles * nblk;
    auto issue_load = [&](int m) {              // lane 0: residual block of global block index m
      const int tl = m / nblk, b = m - tl * nblk;
      const TileId6 T = tile_of6((int)blockIdx.x + tl * (int)gridDim.x, nct, nrt);
      const int bi = m % NB;
      if (wb) {
        mbar_arrive_expect_tx(&efull[bi], 16384u);
        tma_load_3d(smem + S.stg + bi * V6_EBLK, &tm_res, T.ct * BN + 32 * b, T.q0, T.g, &efull[bi]);
      } else {
        mbar_arrive_expect_tx(&efull[bi], 4096u);
        tma_load_3d(ebuf + bi * V6_EBLK, &tm_res, T.ct * BN + 32 * b, T.q0 + quad * 32, T.g, &efull[bi]);
      }
    };
    if (leader) {
      tma_prefetch_desc(&tm_out);
      if (has_res) {
        tma_prefetch_desc(&tm_res);
        for (int m = 0; m < LA && m < total_blk; ++m) issue_load(m);
      }
    }
    int j = 0;
    for (int tl = 0; tl < my_tiles; ++tl) {
      const TileId6 T = tile_of6((int)blockIdx.x + tl * (int)gridDim.x, nct, nrt);
      const int acc = tl & 1;
      const int co0 = T.ct * BN;
      __syncwarp();
      for (int c = lane; c < BN; c += 32) {
        const int co = co0 + c;
        float v = 0.f;
        if (co < P.Cout) {
          if (P.bias) v = __ldg(P.bias + co);
          if (P.epi == EPI_ADDVEC) v += __ldg(P.evec + (long)T.g * P.evec_gstride + co);
        }
        cvs[c] = v;
      }
      __syncwarp();
      bool acc_ready = false;
      long long t_epi0 = 0;
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 32, ++j) {
        const int bi = j % NB;
        if (leader) {
          if (has_res) {
            tma_wait_group_read<1>();             // store j-2 has left its buffer == the buffer of block j+LA
            if (j + LA < total_blk) issue_load(j + LA);
          } else if (!wb) {                       // buffer bi was last used by store j-NB
            if (NB >= 4) tma_wait_group_read<3>();
            else if (NB == 3) tma_wait_group_read<2>();
            else tma_wait_group_read<1>();
          }
        }
        if (!wb) __syncwarp();
        if (!acc_ready) {
          DBG_WAIT6(5, mbar_wait(&acc_full[acc], (uint32_t)((tl >> 1) & 1)));
          tc_fence_after();
          acc_ready = true;
          t_epi0 = dbg_on ? clock64() : 0;
        }
        uint32_t rg[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * ACCW + (uint32_t)cb;
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
        if (stk) tc6_ld32_add(taddr + (uint32_t)BN, rg);
        if (cb + 32 >= BN) {                      // last TMEM read of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(&acc_empty[acc]);
        }
        if (has_res) DBG_WAIT6(4, mbar_wait(&efull[bi], (uint32_t)((j / NB) & 1)));
        uint8_t* buf = ebuf + bi * V6_EBLK;
        // two halves of four 16-byte cells: all shared-memory loads of a half are issued before any store
        // (the compiler cannot prove that the in-place stores do not alias the loads and would serialise them)
        const uint32_t buf_sh = smem_u32(buf), cvs_sh = smem_u32(cvs + cb);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float4 rr[4], cc[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int qd = 4 * hh + i;
            if (has_res) rr[i] = lds128(buf_sh + sw128(lane, qd));
            else rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            cc[i] = lds128(cvs_sh + 16u * (uint32_t)qd);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int qd = 4 * hh + i;
            float4 v = make_float4(fmaf(__uint_as_float(rg[4 * qd]), dsc, cc[i].x), fmaf(__uint_as_float(rg[4 * qd + 1]), dsc, cc[i].y),
                                   fmaf(__uint_as_float(rg[4 * qd + 2]), dsc, cc[i].z), fmaf(__uint_as_float(rg[4 * qd + 3]), dsc, cc[i].w));
            v.x += rr[i].x; v.y += rr[i].y; v.z += rr[i].z; v.w += rr[i].w;
            switch (P.epi) {
              case EPI_ACC: v.x = __fmul_rn(v.x, P.scale); v.y = __fmul_rn(v.y, P.scale); v.z = __fmul_rn(v.z, P.scale); v.w = __fmul_rn(v.w, P.scale); break;
              case EPI_RELU: v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); break;
              case EPI_TANH: v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); break;
              case EPI_MISH: v.x = mishf_(v.x); v.y = mishf_(v.y); v.z = mishf_(v.z); v.w = mishf_(v.w); break;
              case EPI_SILU: v.x = siluf_(v.x); v.y = siluf_(v.y); v.z = siluf_(v.z); v.w = siluf_(v.w); break;
              default: break;
            }
            sts128(buf_sh + sw128(lane, qd), v);
          }
        }
        fence_proxy_async();
        if (wb) {
          if (leader && !has_res) {               // the NEXT block's buffer was last read by store j+1-NB: make sure
            if (NB >= 4) tma_wait_group_read<2>();       // it is free before anybody passes the barrier below
            else if (NB == 3) tma_wait_group_read<1>();
            else tma_wait_group_read<0>();
          }
          named_bar_sync(2, 128);
          if (leader) {
            const uint8_t* blk = smem + S.stg + bi * V6_EBLK;
            if (red_add) tma_reduce_add_3d(&tm_out, co0 + cb, T.q0, T.g, blk);
            else tma_store_3d(&tm_out, co0 + cb, T.q0, T.g, blk);
            tma_commit_group();
          }
        } else {
          __syncwarp();
          if (lane == 0) {
            if (red_add) tma_reduce_add_3d(&tm_out, co0 + cb, T.q0 + quad * 32, T.g, buf);
            else tma_store_3d(&tm_out, co0 + cb, T.q0 + quad * 32, T.g, buf);
            tma_commit_group();
          }
        }
      }
      if (dbg_on) dbgacc[6] += clock64() - t_epi0;
    }
    if (leader) tma_wait_group<0>();
  } else {
    // =========================== epilogue warps (0..3) ===========================
    const int quad = warp;                       // TMEM lane quadrant this warp may access
    int* rowp_ring = reinterpret_cast<int*>(smem + S.rowp);
    uint8_t* stg = smem + S.stg + quad * (32 * 128);      // this warp's 32 staging rows
    const float dsc = P.tc_descale;
    const float* pf0 = nullptr; long gs0 = 0; int pitch0 = 0;   // epilogue operands worth an L2 prefetch
    const float* pf1 = nullptr; long gs1 = 0; int pitch1 = 0;
    if ((P.epi == EPI_RES || P.epi == EPI_ACC || P.epi == EPI_GATE || P.epi == EPI_GEGLU) && P.res) {
      pf0 = P.res; gs0 = P.res_gstride; pitch0 = P.res_pitch;
    }
    if (P.epi == EPI_ACC && P.accumulate) { pf1 = P.out; gs1 = P.out_gstride; pitch1 = P.out_pitch; }
    if (P.epi == EPI_DIFFOUT) { pf0 = P.out; gs0 = P.out_gstride; pitch0 = P.out_pitch; }
    for (int tl = 0; tl < my_tiles; ++tl) {
      const TileId6 T = tile_of6((int)blockIdx.x + tl * (int)gridDim.x, nct, nrt);
      const int acc = tl & 1;
      int* rowp = rowp_ring + acc * TC_ROWS + quad * 32;    // warp-private slice
      {
        const int qq = T.q0 + quad * 32 + lane;
        int p = -1;
        if (qq < Lv) {
          if (Wv) {
            const int h = qq / Wv, w = qq - h * Wv;
            if (w < P.Wreal) p = h * P.Wreal + w;
          } else {
            p = qq;
          }
        }
        rowp[lane] = p;
      }
      __syncwarp();
      const int co0 = T.ct * BN;
      if (pf0 || pf1) {   // this warp's 32 rows x BN columns of the residual / old output -> L2
        constexpr int lines = (BN * 4) / 128 > 0 ? (BN * 4) / 128 : 1;
        for (int idx = lane; idx < 32 * lines; idx += 32) {
          const int p = rowp[idx / lines];
          const int co = co0 + (idx % lines) * 32;
          if (p >= 0 && co < P.Cout) {
            if (pf0) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf0 + T.g * gs0 + (long)p * pitch0 + co));
            if (pf1) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf1 + T.g * gs1 + (long)p * pitch1 + co));
          }
        }
      }
      float4 prea[8];                             // additive epilogue operand (residual / old x), one block ahead
      int pp[8];
      auto load_block = [&](int cb) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = lane + i * 32;          // (row = idx >> 3, 16-byte chunk j = idx & 7) of this warp's 32 x 32 block
          pp[i] = rowp[idx >> 3];
          if (pp[i] >= 0) prea[i] = epi_load_a(P, T.g, pp[i], co0 + cb + 4 * (idx & 7));
        }
      };
      load_block(0);                              // global reads in flight while the tile is still accumulating
      DBG_WAIT6(5, mbar_wait(&acc_full[acc], (uint32_t)((tl >> 1) & 1)));
      tc_fence_after();
      const long long t_epi0 = dbg_on ? clock64() : 0;
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 32) {
        uint32_t rg[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * ACCW + (uint32_t)cb;
        const long long t_ld0 = dbg_on ? clock64() : 0;
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
        if (stk) tc6_ld32_add(taddr + (uint32_t)BN, rg);
        if (dbg_on) dbgacc[4] += clock64() - t_ld0;
        if (cb + 32 >= BN) {                      // last TMEM read of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(&acc_empty[acc]);
        }
        __syncwarp();                             // the previous block's staging rows have been consumed
#pragma unroll
        for (int qd = 0; qd < 8; ++qd)
          *reinterpret_cast<float4*>(stg + sw128(lane, qd)) =
              make_float4(__uint_as_float(rg[4 * qd]) * dsc, __uint_as_float(rg[4 * qd + 1]) * dsc,
                          __uint_as_float(rg[4 * qd + 2]) * dsc, __uint_as_float(rg[4 * qd + 3]) * dsc);
        __syncwarp();
        const int jc = lane & 7;                  // all 8 items of this lane share the 4-channel group
        float4 cv = epi_colvec(P, T.g, co0 + cb + 4 * jc);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = (lane >> 3) + 4 * i;
          if (pp[i] >= 0) {
            EpiPre e;
            e.a = prea[i];
            e.b = epi_load_b(P, T.g, pp[i], co0 + cb + 4 * jc);     // old accumulator: L2-prefetched at tile start
            epi_store_cv(P, T.g, pp[i], co0 + cb + 4 * jc, *reinterpret_cast<const float4*>(stg + sw128(row, jc)), e, cv);
          }
        }
        if (cb + 32 < BN) load_block(cb + 32);
      }
      if (dbg_on) dbgacc[6] += clock64() - t_epi0;
    }
  }

  if (dbg_on) {
    if (warp == 4 && lane == 0) { dbg[1] = dbgacc[1]; dbg[2] = dbgacc[2]; dbg[3] = dbgacc[3]; }
    if (warp == 5 && lane == 0) dbg[7] = dbgacc[7];
    if (tid == 0) { dbg[4] = dbgacc[4]; dbg[5] = dbgacc[5]; dbg[6] = dbgacc[6]; }   // [4]: epilogue time inside tcgen05.ld + wait::ld
  }
  tc_fence_before();
  __syncthreads();
  if (dbg_on && tid == 0) dbg[0] = clock64() - t_begin;
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int BN>
static void launch6(const TapConvParams& P, const CUtensorMap& tr, const CUtensorMap& to, int RRA, int grid, size_t smem,
                    cudaStream_t st) {
  const int NI = cdiv(RRA, 32);
  if (BN == 32 && P.Cin <= 32 && RRA <= 192) {   // narrow layers: 64 rows per pass, two register sets in flight
    launch_pdl(tcconv6_kernel<(BN == 32 ? 32 : 64), 3, true>, dim3(grid), dim3(V6_THREADS), smem, st, P, tr, to);
    return;
  }
  if (NI <= 5) launch_pdl(tcconv6_kernel<BN, 5, false>, dim3(grid), dim3(V6_THREADS), smem, st, P, tr, to);
  else if (NI <= 6) launch_pdl(tcconv6_kernel<BN, 6, false>, dim3(grid), dim3(V6_THREADS), smem, st, P, tr, to);
  else launch_pdl(tcconv6_kernel<BN, 10, false>, dim3(grid), dim3(V6_THREADS), smem, st, P, tr, to);
}
template <int BN>
static void attrs6() {
  AGPT_CUDA(cudaFuncSetAttribute(tcconv6_kernel<BN, 5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn6));
  AGPT_CUDA(cudaFuncSetAttribute(tcconv6_kernel<BN, 6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn6));
  AGPT_CUDA(cudaFuncSetAttribute(tcconv6_kernel<BN, 10, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn6));
  if (BN == 32)
    AGPT_CUDA(cudaFuncSetAttribute(tcconv6_kernel<32, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn6));
}

}  // namespace

static bool tcconv6_try(TapConvParams P, int BN, cudaStream_t st) {
  int lo = P.tap_off[0], hi = P.tap_off[0];
  for (int t = 1; t < P.ntaps; ++t) { lo = std::min(lo, P.tap_off[t]); hi = std::max(hi, P.tap_off[t]); }
  P.lo_al = lo;
  const int RRA = round_up(TC_ROWS + (hi - lo), 8);
  P.R = RRA;
  P.tc_bn = BN;
  const int NI = cdiv(RRA, 32);
  if (NI > 10) return false;
  // epilogue through tensor maps when the output (and residual) rows are affine in the row index
  static int allow_tma = -1;
  if (allow_tma < 0) { const char* e = getenv("AGPT_TC_TMA"); allow_tma = (e && e[0] == '0') ? 0 : 1; }
  const bool epi_ok = P.epi == EPI_BIAS || P.epi == EPI_RES || P.epi == EPI_ACC || P.epi == EPI_RELU || P.epi == EPI_ADDVEC ||
                      P.epi == EPI_TANH || P.epi == EPI_MISH || P.epi == EPI_SILU;
  const bool has_res = (P.epi == EPI_RES || P.epi == EPI_ACC) && P.res != nullptr;
  CUtensorMap tm_res, tm_out;
  memset(&tm_res, 0, sizeof(tm_res));
  memset(&tm_out, 0, sizeof(tm_out));
  static int box_rows = 0;
  if (!box_rows) { const char* e = getenv("AGPT_TC_EPIBOX"); box_rows = (e && atoi(e) == 32) ? 32 : 128; }
  bool tma = allow_tma && epi_ok && P.Wreal == 0;
  if (tma) tma = tma_encode_rows(&tm_out, P.out, P.Cout, P.L, P.G, P.out_pitch, P.out_gstride, box_rows);
  if (tma && has_res) tma = tma_encode_rows(&tm_res, P.res, P.Cout, P.L, P.G, P.res_pitch, P.res_gstride, box_rows);
  if (tma && box_rows == 128) P.tc_flags |= (int)V6_FLAG_WBLK;
  const long fixed = 1024 /*align*/ + (4 * 256 * 4) /*cvs*/ + (4 * RRA * 4 + 2 * TC_ROWS * 4 + 48 * 8 + 64);
  // a weight stage holds tps taps (~32 KB): amortises the per-stage handshake on narrow layers
  const int tps = std::max(1, std::min(P.ntaps, (int)(32768 / (2L * BN * 128))));
  P.tc_tps = tps;
  const long abytes = 2L * RRA * 128, wbytes = (long)tps * 2L * BN * 128;
  const int stages_per_tile = P.tc_chunks_h * cdiv(P.ntaps, tps);
  int NA = 2, NB = 1, NW = 0;
  bool ok = false;
  for (int nb = tma ? 4 : 1; nb >= (tma ? 2 : 1) && !ok; --nb) {
    const long avail = (long)kMaxDyn6 - fixed - (long)nb * V6_EBLK - NA * abytes;
    const int nw = (int)std::min<long>(std::min(MAX_NW6, std::max(2, 2 * stages_per_tile)), avail / wbytes);
    if (nw >= std::min(3, 2 * stages_per_tile) || (nb == (tma ? 2 : 1) && nw >= 2)) { NB = nb; NW = nw; ok = true; }
  }
  if (!ok) return false;      // a single operand buffer cannot overlap transform and MMA -> v5
  // without the TMA epilogue the 4 epilogue warps (LSU loads/stores, one 32-column block in flight) are the
  // bottleneck unless the tile has many taps to hide them behind (profiles/r1c_conv_microbench.txt)
  if (!tma && P.ntaps < 7 && !(P.tc_flags_user & 128)) return false;
  if (P.ntaps == 1) {         // GEMM-like: a third activation buffer when it still leaves 3 weight stages
    const long avail3 = (long)kMaxDyn6 - fixed - (long)NB * V6_EBLK - 3 * abytes;
    if (avail3 >= 3 * wbytes) { NA = 3; NW = (int)std::min<long>(MAX_NW6, avail3 / wbytes); }
  }
  P.tc_na = NA; P.tc_nw = NW; P.tc_nb = NB;
  if (tma) P.tc_flags |= (int)V6_FLAG_TMA;
  Tc6Smem S;
  tc6_layout(S, BN, RRA, NA, NW, NB, tps);
  const size_t smem = (size_t)S.total + 1024;
  if (smem > (size_t)kMaxDyn6) return false;
  const int Wv = P.Wreal > 0 ? P.Wreal + 1 : 0;
  const int Lv = Wv ? (P.L / P.Wreal) * Wv : P.L;
  const int ntiles = cdiv(Lv, TC_ROWS) * cdiv(P.Cout, BN) * P.G;
  int dev = 0;
  AGPT_CUDA(cudaGetDevice(&dev));
  static bool attr_done_dev[64] = {false};
  static int sms_dev[64] = {0};
  if (!attr_done_dev[dev & 63]) {
    attrs6<256>(); attrs6<128>(); attrs6<64>(); attrs6<32>();
    AGPT_CUDA(cudaDeviceGetAttribute(&sms_dev[dev & 63], cudaDevAttrMultiProcessorCount, dev));
    attr_done_dev[dev & 63] = true;
  }
  const int grid = std::min(ntiles, sms_dev[dev & 63]);
  if (BN == 256) launch6<256>(P, tm_res, tm_out, RRA, grid, smem, st);
  else if (BN == 128) launch6<128>(P, tm_res, tm_out, RRA, grid, smem, st);
  else if (BN == 64) launch6<64>(P, tm_res, tm_out, RRA, grid, smem, st);
  else launch6<32>(P, tm_res, tm_out, RRA, grid, smem, st);
  return true;
}

// Persistent schedule: pays when a CTA gets more than one tile (otherwise there is nothing to overlap
// and v5's 8 transform+epilogue warps are at least as good).  Returns false -> caller uses v5.
bool tcconv6_launch(TapConvParams P, cudaStream_t st, bool force) {
  if (!P.w_h || P.strips > 0) return false;     // strip-tiled wide images run on tcconv5
  int dev = 0;
  AGPT_CUDA(cudaGetDevice(&dev));
  static int sms_dev[64] = {0};
  if (!sms_dev[dev & 63]) AGPT_CUDA(cudaDeviceGetAttribute(&sms_dev[dev & 63], cudaDevAttrMultiProcessorCount, dev));
  const int sms = sms_dev[dev & 63];
  const HTile c = pick_h_tile(P, sms, false);
  if (!force && c.ntiles <= sms) return false;
  if (c.bn != P.tc_bn) {
    TapConvParams Q = P;
    Q.w_h = c.w;
    if (tcconv6_try(Q, c.bn, st)) return true;
    const long ntiles = c.ntiles / cdiv(P.Cout, c.bn) * cdiv(P.Cout, P.tc_bn);
    if (!force && ntiles <= sms) return false;
  }
  return tcconv6_try(P, P.tc_bn, st);
}

}  // namespace agpt
