// tcconv v7: the persistent tap-GEMM fed by PRODUCER-SIDE fp16 hi/lo OPERAND PLANES -- no transform warps, no LSU
// work on the activation side.
//
// The A operand arrives as two fp16 planes hi/lo [G][L][C] written by the PRODUCING layer's epilogue (the consumer's
// prologue -- leaky-relu on the HiFi-GAN path -- already applied), so the kernel is a pure TMA -> tcgen05 pipeline:
//
//   warps 0-3   epilogue: TMEM -> regs -> (+ TMA-loaded residual) -> fp32 block -> TMA store / reduce-add,
//               and/or prologue(out) -> fp16 hi/lo -> plane blocks -> two TMA stores        (whole-block boxes)
//   warp  4     MMA issuer (elected lane), two TMEM accumulators, multi-tap weight stages    (as tcconv6)
//   warp  5     weight producer (cp.async.bulk ring)                                         (as tcconv6)
//   warp  6     activation producer: per chunk TWO tensor-map loads (hi, lo planes) of the [RRA rows x 64 ch]
//               K-major SWIZZLE_128B tile at row coordinate q0 + lo.  Rows < 0 or >= L and channels >= C are
//               zero-filled by the TMA engine = the conv's zero padding (the planes hold prologue(x), and
//               prologue(0) = 0 for leaky-relu / SiLU / identity).
//
// A conv tap is still a row-shifted UMMA descriptor of the same tile: the TMA engine writes SWIZZLE_128B tiles with
// the same absolute-address XOR pattern (16-byte chunk ^ (row & 7), 1 KB-aligned tile) that sw128() produces.
// Why (profiles/r1e_findings.md): in tcconv6 the 8 transform warps' LDG/STS traffic and the epilogue's LDS/STS share
// one memory-instruction queue -- the epilogue needed ~3.4 k cycles per [128 x 32] block and bound every k = 3 layer
// and the narrow stages; with the planes the queue serves the epilogue alone.
#include <cuda_fp16.h>
#include "tapconv.cuh"
#include "tapconv_epi.cuh"
#include "tc_common.cuh"
#include "tc_h16.cuh"
#include "tc_tma.cuh"
#include "models.h"

namespace agpt {

// fp16 plane [G][L][C] as a 3-D tensor {C, L, G}; box = {box_c channels, box_rows, 1}
inline bool tma_encode_plane(CUtensorMap* map, const __half* base, int C, long L, int G, long pitch, long gstride,
                             int box_c, int box_rows, bool swizzle128) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess) return false;
    fn = reinterpret_cast<EncodeFn>(p);
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch % 8) != 0 || (gstride % 8) != 0) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)(G > 0 ? G : 1)};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)(G > 1 ? gstride : pitch * L) * 2};
  const cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

namespace {

constexpr int V7_THREADS = 256;               // 4 epilogue + MMA + W producer + A producer (+ 1 idle) warps
constexpr int MAX_NA7 = 4, MAX_NW7 = 6;
constexpr int kMaxDyn7 = 227 * 1024 - 512;
constexpr int V7_EBLK = TC_ROWS * 128;        // fp32 epilogue block [128 rows][32 cols]
constexpr int V7_PBLK = TC_ROWS * 64;         // one fp16 plane block [128 rows][32 cols] (no swizzle)

struct Tc7Smem {
  uint32_t a_hi[MAX_NA7], a_lo[MAX_NA7], w[MAX_NW7], stg, pl, cvs, bars, tmem_slot, total;
};
__host__ __device__ inline void tc7_layout(Tc7Smem& s, int BN, int RRA, int NA, int NW, int NB, int tps, int planes) {
  uint32_t o = 0;
  for (int i = 0; i < MAX_NA7; ++i) { s.a_hi[i] = o; if (i < NA) o += RRA * 128; }
  for (int i = 0; i < MAX_NA7; ++i) { s.a_lo[i] = o; if (i < NA) o += RRA * 128; }
  for (int i = 0; i < MAX_NW7; ++i) { s.w[i] = o; if (i < NW) o += tps * 2 * BN * 128; }
  s.stg = o; o += NB * V7_EBLK;                       // fp32 blocks (residual in / result out), SWIZZLE_128B
  s.pl = o; o += planes ? 2 * 2 * V7_PBLK : 0;        // 2 buffers x (hi, lo) plane blocks
  s.cvs = o; o += 4 * 256 * 4;
  o = (o + 15) & ~15u;
  s.bars = o; o += 48 * 8;
  s.tmem_slot = o; o += 16;
  s.total = o;
}

struct TileId7 { int g, q0, ct; };
__device__ __forceinline__ TileId7 tile_of7(int t, int nct, int nrt) {
  TileId7 r;
  r.ct = t % nct;
  const int u = t / nct;
  r.q0 = (u % nrt) * TC_ROWS;
  r.g = u / nrt;
  return r;
}

__device__ __forceinline__ float pro_scalar(int pro, float slope, float v) {
  if (pro == PRO_LRELU) return lrelu(v, slope);
  if (pro == PRO_SILU) return siluf_(v);
  return v;
}

// accumulate 32 more fp32 TMEM columns of this lane into rg (the A_hi x W_lo range of a stacked accumulator)
__device__ __forceinline__ void tc7_ld32_add(uint32_t taddr, uint32_t* rg) {
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


template <int BN>
__global__ void __launch_bounds__(V7_THREADS, 1)
tcconv7_kernel(const __grid_constant__ TapConvParams P, const __grid_constant__ PlaneIO Q,
               const __grid_constant__ CUtensorMap tm_ahi, const __grid_constant__ CUtensorMap tm_alo,
               const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ CUtensorMap tm_out,
               const __grid_constant__ CUtensorMap tm_phi, const __grid_constant__ CUtensorMap tm_plo) {
  extern __shared__ uint8_t smem_raw_[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~(uintptr_t)1023);
  const int RRA = P.R, NA = P.tc_na, NW = P.tc_nw, NB = P.tc_nb, tps = P.tc_tps;
  const bool planes = Q.out_hi != nullptr;
  __shared__ Tc7Smem S;
  if (threadIdx.x == 0) tc7_layout(S, BN, RRA, NA, NW, NB, tps, planes ? 1 : 0);
  __syncthreads();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S.bars);
  uint64_t* a_full = bars + 0;             // [MAX_NA7]
  uint64_t* a_empty = bars + MAX_NA7;      // [MAX_NA7]
  uint64_t* w_full = bars + 2 * MAX_NA7;   // [MAX_NW7]
  uint64_t* w_empty = w_full + MAX_NW7;    // [MAX_NW7]
  uint64_t* acc_full = w_empty + MAX_NW7;  // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint64_t* e_full = acc_empty + 2;        // [8] residual block buffers
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + S.tmem_slot);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunks = P.tc_chunks_h, ntaps = P.ntaps, iters_per_tile = nchunks * ntaps;
  const int lo = P.lo_al;
  const int nct = (P.Cout + BN - 1) / BN, nrt = (P.L + TC_ROWS - 1) / TC_ROWS;
  const int ntiles = nct * nrt * P.G;
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  // two accumulators of stacked weight parts ([A_hi W_hi | A_hi W_lo] by one MMA of width 2 BN, + A_lo W_hi: tcconv6.cu)
  constexpr uint32_t ACCW = 2 * BN;
  constexpr uint32_t TMEM_COLS = (2 * ACCW <= 32) ? 32 : (2 * ACCW <= 64 ? 64 : (2 * ACCW <= 128 ? 128 : (2 * ACCW <= 256 ? 256 : 512)));   // power of two
  const bool stk = !(P.tc_flags & 4);

  if (tid == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
    for (int i = 0; i < 8; ++i) mbar_init(&e_full[i], 1);
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

  if (warp == 6) {
    // =========================== activation producer (TMA) ===========================
    if (lane == 0) {
      tma_prefetch_desc(&tm_ahi);
      tma_prefetch_desc(&tm_alo);
      const uint32_t tile_bytes = (uint32_t)RRA * 128u;      // one part; boxes are [RRA rows][64 ch] fp16
      int k = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const TileId7 T = tile_of7((int)blockIdx.x + tl * (int)gridDim.x, nct, nrt);
        for (int c = 0; c < nchunks; ++c, ++k) {
          const int buf = k % NA, n = k / NA;
          if (n >= 1) mbar_wait(&a_empty[buf], (uint32_t)((n - 1) & 1));
          mbar_arrive_expect_tx(&a_full[buf], 2u * tile_bytes);
          // (RRA <= 256: one box per part; the host splits taller halos into two boxes)
          tma_load_3d(smem + S.a_hi[buf], &tm_ahi, c * H_KCH, T.q0 + lo, T.g, &a_full[buf]);
          tma_load_3d(smem + S.a_lo[buf], &tm_alo, c * H_KCH, T.q0 + lo, T.g, &a_full[buf]);
        }
      }
    }
  } else if (warp == 4) {
    // =========================== MMA issuer (as tcconv6) ===========================
    const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
    const uint64_t DC = make_desc(0);
    const uint32_t a16 = smem_u32(smem + S.a_hi[0]) >> 4;
    const uint32_t abuf16 = (uint32_t)(RRA * 128) >> 4;
    const uint32_t alo16 = (uint32_t)NA * abuf16;
    const uint32_t w16 = smem_u32(smem + S.w[0]) >> 4;
    const uint32_t wtap16 = (uint32_t)(2 * BN * 128) >> 4;
    const uint32_t wstage16 = (uint32_t)tps * wtap16;
    constexpr uint32_t wlo16 = (uint32_t)(BN * 128) >> 4;
    int gc = 0, it = 0;
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int acc = tl & 1, na = tl >> 1;
      if (na >= 1) { mbar_wait(&acc_empty[acc], (uint32_t)((na - 1) & 1)); tc_fence_after(); }
      const uint32_t tmem_d = tmem_base + (uint32_t)acc * ACCW;
      uint32_t nz = 0;
      for (int c = 0; c < nchunks; ++c, ++gc) {
        const int buf = gc % NA;
        const int kv = min(H_KCH, P.Cin - c * H_KCH);
        const int ksteps = (kv + 15) >> 4;
        mbar_wait(&a_full[buf], (uint32_t)((gc / NA) & 1));
        tc_fence_after();
        const uint64_t dA = DC + (uint64_t)(a16 + (uint32_t)buf * abuf16);
        for (int t0 = 0; t0 < ntaps; t0 += tps, ++it) {
          const int s = it % NW;
          const int t1 = min(ntaps, t0 + tps);
          mbar_wait(&w_full[s], (uint32_t)((it / NW) & 1));
          tc_fence_after();
          const uint64_t dW = DC + (uint64_t)(w16 + (uint32_t)s * wstage16);
          if (elect_one()) {
            for (int t = t0; t < t1; ++t) {
              const uint64_t dah = dA + (uint64_t)((uint32_t)(P.tap_off[t] - lo) * 8u);
              const uint64_t dal = dah + alo16;
              const uint64_t dwh = dW + (uint64_t)((uint32_t)(t - t0) * wtap16);
              const uint64_t dwl = dwh + wlo16;
              for (int k = 0; k < ksteps; ++k) {
                const uint64_t ko = (uint64_t)(2 * k);
                if (stk) {
                  umma_f16(tmem_d, dah + ko, dwh + ko, idesc2, nz);        // [hi x hi | hi x lo]
                  nz = 1u;
                  umma_f16(tmem_d, dal + ko, dwh + ko, idesc, 1u);         // += lo x hi
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
  } else if (warp == 5) {
    // =========================== weight producer (as tcconv6) ===========================
    if (lane == 0) {
      const uint32_t tapbytes = 2u * BN * 128u;
      int it = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const TileId7 T = tile_of7((int)blockIdx.x + tl * (int)gridDim.x, nct, nrt);
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(P.w_h) + (size_t)T.ct * (size_t)iters_per_tile * tapbytes;
        for (int c = 0; c < nchunks; ++c) {
          for (int t0 = 0; t0 < ntaps; t0 += tps, ++it) {
            const int s = it % NW, n = it / NW;
            const uint32_t bytes = (uint32_t)(min(ntaps, t0 + tps) - t0) * tapbytes;
            if (n >= 1) mbar_wait(&w_empty[s], (uint32_t)((n - 1) & 1));
            mbar_arrive_expect_tx(&w_full[s], bytes);
            bulk_g2s(smem + S.w[s], wsrc + (size_t)(c * ntaps + t0) * tapbytes, bytes, &w_full[s]);
          }
        }
      }
    }
  } else if (warp < 4) {
    // =========================== epilogue warps (whole-block TMA boxes) ===========================
    const int quad = warp;
    const int LA = NB - 2;
    const bool has_res = (P.epi == EPI_RES || P.epi == EPI_ACC) && P.res != nullptr;
    const bool red_add = (P.epi == EPI_ACC) && P.accumulate;
    const bool f32_out = Q.store_f32 != 0;
    const bool leader = tid == 0;
    float* cvs = reinterpret_cast<float*>(smem + S.cvs) + quad * 256;
    const float dsc = P.tc_descale;
    constexpr int nblk = BN / 32;
    const int total_blk = my_tiles * nblk;
    auto issue_load = [&](int m) {
      const int tl = m / nblk, b = m - tl * nblk;
      const TileId7 T = tile_of7((int)blockIdx.x + tl * (int)gridDim.x, nct, nrt);
      const int bi = m % NB;
      mbar_arrive_expect_tx(&e_full[bi], 16384u);
      tma_load_3d(smem + S.stg + bi * V7_EBLK, &tm_res, T.ct * BN + 32 * b, T.q0, T.g, &e_full[bi]);
    };
    if (leader) {
      if (f32_out) tma_prefetch_desc(&tm_out);
      if (planes) { tma_prefetch_desc(&tm_phi); tma_prefetch_desc(&tm_plo); }
      if (has_res) {
        tma_prefetch_desc(&tm_res);
        for (int m = 0; m < LA && m < total_blk; ++m) issue_load(m);
      }
    }
    int j = 0;
    for (int tl = 0; tl < my_tiles; ++tl) {
      const TileId7 T = tile_of7((int)blockIdx.x + tl * (int)gridDim.x, nct, nrt);
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
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 32, ++j) {
        const int bi = NB > 0 ? j % NB : 0;
        // Group accounting: every block commits ONE bulk group (fp32 store and/or the two plane stores), so
        // "wait_group.read 1" = everything of block j-2 has left shared memory: its fp32 buffer (== the buffer of
        // block j+LA) and its plane buffer (== the plane buffer of block j, two plane buffers alternate).
        if (leader) {
          tma_wait_group_read<1>();
          if (has_res && j + LA < total_blk) issue_load(j + LA);
        }
        named_bar_sync(2, 128);                  // plane buffer (j & 1) and fp32 buffer bi are free for everybody
        if (!acc_ready) {
          mbar_wait(&acc_full[acc], (uint32_t)((tl >> 1) & 1));
          tc_fence_after();
          acc_ready = true;
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
        if (stk) tc7_ld32_add(taddr + (uint32_t)BN, rg);
        if (cb + 32 >= BN) {
          tc_fence_before();
          mbar_arrive(&acc_empty[acc]);
        }
        if (has_res) mbar_wait(&e_full[bi], (uint32_t)((j / NB) & 1));
        const uint32_t buf_sh = smem_u32(smem + S.stg + bi * V7_EBLK + quad * 4096);
        const uint32_t cvs_sh = smem_u32(cvs + cb);
        const uint32_t pl_sh = smem_u32(smem + S.pl + (j & 1) * 2 * V7_PBLK + quad * 2048);   // hi block, lo at + V7_PBLK
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float4 rr[4], cc[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int qd = 4 * hh + i;
            rr[i] = has_res ? lds128(buf_sh + sw128(lane, qd)) : make_float4(0.f, 0.f, 0.f, 0.f);
            cc[i] = lds128(cvs_sh + 16u * (uint32_t)qd);
          }
          uint32_t hi[8], lw[8];
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
            if (f32_out) sts128(buf_sh + sw128(lane, qd), v);
            if (planes) {      // NOTE: a layer that accumulates (red_add) cannot emit planes of the final sum
              const float p0 = pro_scalar(Q.out_pro, Q.out_slope, v.x), p1 = pro_scalar(Q.out_pro, Q.out_slope, v.y);
              const float p2 = pro_scalar(Q.out_pro, Q.out_slope, v.z), p3 = pro_scalar(Q.out_pro, Q.out_slope, v.w);
              hi[2 * i] = split2(p0, p1, lw[2 * i]);
              hi[2 * i + 1] = split2(p2, p3, lw[2 * i + 1]);
            }
          }
          if (planes) {        // this lane's row: 16 fp16 of this half = 32 bytes per plane (row pitch 64 B, no swizzle)
            const uint32_t ro = pl_sh + (uint32_t)lane * 64u + (uint32_t)hh * 32u;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ro), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ro + 16u), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ro + V7_PBLK), "r"(lw[0]), "r"(lw[1]), "r"(lw[2]), "r"(lw[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ro + V7_PBLK + 16u), "r"(lw[4]), "r"(lw[5]), "r"(lw[6]), "r"(lw[7]) : "memory");
          }
        }
        fence_proxy_async();
        named_bar_sync(2, 128);
        if (leader) {
          if (f32_out) {
            const uint8_t* blk = smem + S.stg + bi * V7_EBLK;
            if (red_add) tma_reduce_add_3d(&tm_out, co0 + cb, T.q0, T.g, blk);
            else tma_store_3d(&tm_out, co0 + cb, T.q0, T.g, blk);
          }
          if (planes) {
            const uint8_t* pb = smem + S.pl + (j & 1) * 2 * V7_PBLK;
            tma_store_3d(&tm_phi, co0 + cb, T.q0, T.g, pb);
            tma_store_3d(&tm_plo, co0 + cb, T.q0, T.g, pb + V7_PBLK);
          }
          tma_commit_group();
        }
      }
    }
    if (leader) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace

// fp32 [G][L][C] -> prologue -> fp16 hi/lo planes: the entry of a network (mel input) and the test harness
__global__ void make_planes_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, long n,
                                   int pro, float slope) {
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i + 1 >= n + 1) return;
  const float a = pro_scalar(pro, slope, x[i]);
  const float b = (i + 1 < n) ? pro_scalar(pro, slope, x[i + 1]) : 0.f;
  uint32_t l;
  const uint32_t h = split2(a, b, l);
  if (i + 1 < n) {
    *reinterpret_cast<uint32_t*>(hi + i) = h;
    *reinterpret_cast<uint32_t*>(lo + i) = l;
  } else {
    hi[i] = __ushort_as_half((unsigned short)(h & 0xffff));
    lo[i] = __ushort_as_half((unsigned short)(l & 0xffff));
  }
}

void make_planes(const float* x, __half* hi, __half* lo, long n, int pro, float slope, cudaStream_t st) {
  const long pairs = (n + 1) / 2;
  make_planes_kernel<<<(unsigned)cdivl(pairs, 256), 256, 0, st>>>(x, hi, lo, n, pro, slope);
  count_launch(1);
  AGPT_CUDA(cudaGetLastError());
}

// Launch (1-D layers, bias / residual / accumulate / relu-type epilogues).  Returns false when the layer does not
// qualify; the caller then uses tcconv6 on the fp32 tensor.
bool tcconv7_launch(TapConvParams P, const PlaneIO& Q, cudaStream_t st) {
  P.tc_flags = tc_env_flags() | P.tc_flags_user;
  if (!P.w_h || P.Wreal != 0 || !Q.in_hi || !Q.in_lo) return false;
  const bool epi_ok = P.epi == EPI_BIAS || P.epi == EPI_RES || P.epi == EPI_ACC || P.epi == EPI_RELU || P.epi == EPI_ADDVEC ||
                      P.epi == EPI_TANH || P.epi == EPI_MISH || P.epi == EPI_SILU;
  if (!epi_ok) return false;
  int lo = P.tap_off[0], hi = P.tap_off[0];
  for (int t = 1; t < P.ntaps; ++t) { lo = std::min(lo, P.tap_off[t]); hi = std::max(hi, P.tap_off[t]); }
  P.lo_al = lo;
  const int RRA = round_up(TC_ROWS + (hi - lo), 8);
  if (RRA > 256) return false;                       // one tensor-map box per part (boxDim <= 256)
  P.R = RRA;
  int dev = 0, sms = 148;
  AGPT_CUDA(cudaGetDevice(&dev));
  AGPT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // tile width: a narrower image of a native-128 layer when 128-wide tiles would leave SMs idle (the UNet's GEMMs on
  // 1 560 rows: 640 columns -> 130 tiles of 64 instead of 65 of 128)
  static int pick7 = -1;
  if (pick7 < 0) { const char* e = getenv("AGPT_TC7_PICK"); pick7 = (e && e[0] == '0') ? 0 : 1; }
  if (pick7 && P.strips == 0) {
    const HTile c = pick_h_tile(P, sms, true, false);
    if (c.bn != P.tc_bn && c.w && (c.bn == 64 || c.bn == 96)) { P.w_h = c.w; P.tc_bn = c.bn; }
  }
  const int BN = P.tc_bn;
  const bool planes = Q.out_hi != nullptr;
  const bool has_res = (P.epi == EPI_RES || P.epi == EPI_ACC) && P.res != nullptr;
  if (planes && P.epi == EPI_ACC && P.accumulate) return false;
  CUtensorMap tm_ahi, tm_alo, tm_res, tm_out, tm_phi, tm_plo;
  memset(&tm_res, 0, sizeof(tm_res)); memset(&tm_out, 0, sizeof(tm_out));
  memset(&tm_phi, 0, sizeof(tm_phi)); memset(&tm_plo, 0, sizeof(tm_plo));
  bool ok = tma_encode_plane(&tm_ahi, Q.in_hi, P.Cin, P.L, P.G, Q.in_pitch, Q.in_gstride, H_KCH, RRA, true) &&
            tma_encode_plane(&tm_alo, Q.in_lo, P.Cin, P.L, P.G, Q.in_pitch, Q.in_gstride, H_KCH, RRA, true);
  if (ok && Q.store_f32) ok = tma_encode_rows(&tm_out, P.out, P.Cout, P.L, P.G, P.out_pitch, P.out_gstride, 128);
  if (ok && has_res) ok = tma_encode_rows(&tm_res, P.res, P.Cout, P.L, P.G, P.res_pitch, P.res_gstride, 128);
  if (ok && planes)
    ok = tma_encode_plane(&tm_phi, Q.out_hi, P.Cout, P.L, P.G, Q.outp_pitch, Q.outp_gstride, 32, 128, false) &&
         tma_encode_plane(&tm_plo, Q.out_lo, P.Cout, P.L, P.G, Q.outp_pitch, Q.outp_gstride, 32, 128, false);
  if (!ok) return false;
  const int tps = std::max(1, std::min(P.ntaps, (int)(32768 / (2L * BN * 128))));
  P.tc_tps = tps;
  const long abytes = 2L * RRA * 128, wbytes = (long)tps * 2L * BN * 128;
  const long fixed = 1024 + 4 * 256 * 4 + 48 * 8 + 64 + (planes ? 4L * V7_PBLK : 0);
  // no transform registers / raw buffers any more: the budget goes to a deeper operand and epilogue ring
  // fp32 block buffers only where the epilogue moves fp32 data (a planes-only layer without residual needs none:
  // its 64 KB go to the weight / operand rings)
  const bool need_f32 = Q.store_f32 || has_res;
  const int stages_per_tile = P.tc_chunks_h * cdiv(P.ntaps, tps);
  int NB = need_f32 ? 4 : 0, NA = 2, NW = 0;
  for (;; --NB) {
    const long avail = (long)kMaxDyn7 - fixed - (long)NB * V7_EBLK - NA * abytes;
    NW = (int)std::min<long>(MAX_NW7, avail / wbytes);
    if (NW >= std::min(3, std::max(2, stages_per_tile)) || NB <= (need_f32 ? 2 : 0)) break;
  }
  if (NW < 2) return false;
  while (NA < MAX_NA7 && (long)kMaxDyn7 - fixed - (long)NB * V7_EBLK - (NA + 1) * abytes >= std::max(NW, 3) * wbytes) ++NA;
  NW = (int)std::min<long>(MAX_NW7, ((long)kMaxDyn7 - fixed - (long)NB * V7_EBLK - NA * abytes) / wbytes);
  P.tc_na = NA; P.tc_nw = NW; P.tc_nb = NB;
  Tc7Smem S;
  tc7_layout(S, BN, RRA, NA, NW, NB, tps, planes ? 1 : 0);
  const size_t smem = (size_t)S.total + 1024;
  if (smem > (size_t)kMaxDyn7) return false;
  const int ntiles = cdiv(P.L, TC_ROWS) * cdiv(P.Cout, BN) * P.G;
  static bool attr_done = false;
  if (!attr_done) {
    AGPT_CUDA(cudaFuncSetAttribute(tcconv7_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn7));
    AGPT_CUDA(cudaFuncSetAttribute(tcconv7_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn7));
    AGPT_CUDA(cudaFuncSetAttribute(tcconv7_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn7));
    AGPT_CUDA(cudaFuncSetAttribute(tcconv7_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn7));
    attr_done = true;
  }
  const int grid = std::min(ntiles, sms);
  if (BN == 128) launch_pdl(tcconv7_kernel<128>, dim3(grid), dim3(V7_THREADS), smem, st, P, Q, tm_ahi, tm_alo, tm_res, tm_out, tm_phi, tm_plo);
  else if (BN == 96) launch_pdl(tcconv7_kernel<96>, dim3(grid), dim3(V7_THREADS), smem, st, P, Q, tm_ahi, tm_alo, tm_res, tm_out, tm_phi, tm_plo);
  else if (BN == 64) launch_pdl(tcconv7_kernel<64>, dim3(grid), dim3(V7_THREADS), smem, st, P, Q, tm_ahi, tm_alo, tm_res, tm_out, tm_phi, tm_plo);
  else if (BN == 32) launch_pdl(tcconv7_kernel<32>, dim3(grid), dim3(V7_THREADS), smem, st, P, Q, tm_ahi, tm_alo, tm_res, tm_out, tm_phi, tm_plo);
  else return false;
  return true;
}

}  // namespace agpt
