// Linear / 1x1-conv / im2col-conv GEMM on the 5th-gen tensor cores.
//
//   out[M,N] = epilogue( A[M,K] (bf16, row-major, lda) * W[N,K]^T (bf16, row-major, ldw) )
//
// This is the `nn.Linear` / `nn.Conv2d` arithmetic of the reference forward
// (qkv/proj/fc1/fc2 in lib/models/mixformer_vit/mixformer.py:45-47,123; value/output/ffn
// projections of the fusion encoder, deformable_encoder_lnspecific.py:127-155; folded conv+BN
// of the corner head, lib/models/mixformer_cvt/head.py:7-20) with bias, activation, residual
// and positional-table adds fused into the epilogue.
//
// Structure (one CTA per SM, persistent over output tiles, 320 threads):
//   warp 8      : TMA producer  - cp.async.bulk.tensor 2D loads of the A (128 x 64) and W (BN x 64)
//                 K-slices into a multi-stage 128B-swizzled smem ring (mbarrier full/empty).
//   warp 9      : MMA issuer    - one elected thread issues tcgen05.mma (M=128, N=BN, K=16) x 4
//                 per stage into a double-buffered fp32 accumulator in TMEM; tcgen05.commit
//                 releases smem stages and publishes finished accumulators.
//   warps 0..7  : epilogue      - tcgen05.ld the accumulator (warp w may touch TMEM lanes
//                 32*(w%4)..+31 = 32 output rows; the two warps sharing a quadrant split the
//                 column chunks), bias/act in registers, then a per-warp smem transpose so that
//                 residual loads and output stores are row-contiguous (coalesced) in HBM.
// The accumulator double buffer lets the epilogue of tile i overlap the MMAs of tile i+1.
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "../../include/mmt_b200.h"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

// Developer experiments (epilogue-isolation switches, per-CTA cycle counters, A/B environment switches) exist only in
// builds made with -DMMT_GEMM_DEV (`python multi-modal-tracking_b200/build.py --dev` -> libmmt_b200_dev.so).  The shipped
// library compiles them out: kDev is a compile-time false, the kernel carries no debug branch and the ABI entry points
// read no environment variable.
#ifdef MMT_GEMM_DEV
constexpr bool kDev = true;
#else
constexpr bool kDev = false;
#endif

namespace mmt {
using namespace ptx;

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_SMEM_BUDGET = 192 * 1024;  // operand ring; the rest holds the epilogue staging
constexpr int GEMM_STAGING_WORDS = 32 * 32;   // per epilogue warp: 32 rows x 128 B, XOR-swizzled 16-byte vectors

struct GemmEpi {
  const float* bias;      // [N] or nullptr
  const float* resid;     // fp32 [M, ldr] or nullptr (added after the activation)
  const float* rowadd;    // fp32 [rowadd_period, N] or nullptr: += rowadd[row % period][n]
  void* out;              // bf16 or fp32 [M, ldo]
  int ldr;
  int rowadd_period;
  int ldo;
  int act;                // MMT_ACT_*
  int out_fp32;
  int vec_ok;             // all vector-store alignment preconditions hold
  int prefetch;           // L2-prefetch the next tile's A rows (off unless MMT_GEMM_PREFETCH=1; A/B measurements)
  int tma_store;          // bf16 output tiles leave through TMA stores (tmC valid): plain GEMM, 16-byte aligned rows
  int tma_f32;            // PAIR kernel, fp32 output + fp32 residual: residual tiles arrive and result tiles leave by TMA
                          // (tmR = residual, tmC = output, 32 x 32 fp32 boxes, 128-byte swizzle)
  // ---- LayerNorm folded into the GEMMs around it (mmt_gemm_bf16_ex):
  // consumer side (bf16 output): A holds RAW residual rows, W / bias carry gamma / beta, and the epilogue finishes the
  // normalisation per row: out = rs * (acc - mu * colsum[n]) + bias[n], mu / rs from the row's partial sums
  const float* ln_stats;  // [ln_slots][ln_stride rows][2] (sum, sum of squares) partials of the A rows (slot-major: the 32
                          // rows of a warp read / write 256 contiguous bytes per slot), or nullptr
  const float* colsum;    // [N] fp32: sum over k of the (bf16-rounded) folded weight row
  int ln_slots;
  int ln_stride;          // rows between two slots of ln_stats
  float ln_inv_k;         // 1 / K
  float ln_eps;
  // producer side (fp32 residual output): also emit a bf16 copy of the output rows and their partial sums
  bf16* xb_out;           // [M, ld_xb] or nullptr
  float* stats_out;       // [N / 128][stats_stride rows][2] or nullptr: slot = (256-column tile, epilogue half)
  int stats_stride;
  int ld_xb;
  int dbg_flags;          // MMT_GEMM_DEV builds only (MMT_GEMM_DBG): 1 = epilogue without global traffic, 2 = every tile
                          // loads the operands of tile 0 (pure L2 hits), ...
  long long* dbg;         // MMT_GEMM_DEV builds only: per-CTA cycle counters, see mmt_dev_gemm_timing
};

// Implicit-GEMM geometry of a Conv2d(k=3, pad=1) on an NHWC map [B, H, W, C]: an M tile is a BW x BH pixel box
// (BW * BH <= 128) of one image; the K loop walks 9 taps x ceil(C / 64) channel chunks, each A slice being ONE 4-D
// TMA box shifted by the tap offset (out-of-image pixels and channels arrive as zeros).  enabled == 0: plain GEMM.
struct GemmConv {
  int enabled;
  int H, W, C;
  int BW, BH;
  int tiles_x, tiles_y;    // boxes per image
  int cchunks;             // ceil(C / 64)
};

// PAIR: two CTAs of a cluster compute one 256 x BN tile with cta_group::2 MMAs - each loads its own 128 A rows and
// HALF of the W rows, so a k-slice costs 32 KB of L2->SM traffic and smem per SM instead of 48 KB (BN = 256):
// 4 pipeline stages of 32 KB (the per-warp epilogue tiles take the rest) and 1/3 less operand traffic for the same MMA work.
template <int BN, bool PAIR = false>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // epilogue shared memory per warp: one 4 KB staging tile; PAIR: two 4 KB tiles (fp32 residual / output tiles of the
  // TMA epilogue, double-buffered; the bf16 epilogue uses 2 KB of it for the packed rows + 512 B for its bias slices)
  // (+ 2 KB: the bf16 shadow tile of the folded-LayerNorm producer, mmt_gemm_bf16_ex)
#ifdef MMT_EXP_EPI8K     // developer experiment: 8 KB per warp -> 5 ring stages (the shadow-tile producer is unusable in this build)
  static constexpr int EPI_TILE_BYTES = PAIR ? 8192 : GEMM_STAGING_WORDS * 4;
#else
  static constexpr int EPI_TILE_BYTES = PAIR ? 8192 + 2048 : GEMM_STAGING_WORDS * 4;
#endif
  static constexpr int EPI_BYTES = GEMM_EPI_WARPS * EPI_TILE_BYTES;
  static constexpr int BAR_BYTES = 512;
  // the dynamic shared memory is declared __align__(1024) (checked at kernel start): no alignment slack is reserved
  static constexpr int RING_BUDGET = PAIR ? (227 * 1024 - BAR_BYTES - EPI_BYTES) : GEMM_SMEM_BUDGET;
  static constexpr int STAGES = (RING_BUDGET / STAGE_BYTES) > 8 ? 8 : (RING_BUDGET / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + (PAIR ? 0 : 1024) /*align slack*/ + BAR_BYTES;
  static constexpr int CH = (BN % 32 == 0) ? 32 : 16;  // epilogue column chunk
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                        : (2 * BN <= 256) ? 256 : 512;
};

// ---------------------------------------------------------------------------------------------------------------------
// Lean epilogues of the CTA-pair kernel.  The backbone's four GEMMs (qkv, fc1: bias [+ folded LayerNorm] [+ GELU] -> bf16;
// proj, fc2: bias + fp32 residual in place [+ shadow rows and partial sums]) take these specialised loops: compile-time
// activation / LayerNorm switches, thread == row throughout, nothing of the generic path's bookkeeping (ragged tiles,
// strided stores, positional adds, convolution geometry).  The generic loop in the kernel body cost ~390 warp instructions
// per 32 x 32 chunk of which ~60 were the chunk's arithmetic (ncu source page, profiles/r2_gemm_epilogue.md); the epilogue
// warps' instruction stream is what bounds the K = 768 GEMMs.
struct EpiCtx {
  uint32_t tmem_base;
  uint32_t tfull0, tempty0;      // shared addresses of tmem_full[0] / tmem_empty[0] (slot 1 = + 8)
  uint32_t rbar0;                // this warp's two residual-tile barriers
  uint4* stg4;                   // this warp's epilogue region (Cfg::EPI_TILE_BYTES)
  int worker, n_workers, num_tiles, n_tiles, M, N, warp, lane, cta_rank;
  int mt_mul, mt_add;            // 256-row block of scheduler tile t: (t / n_tiles) * mt_mul + mt_add (cluster of two pairs: 2, pair index)
};

template <int BN, int ACT, bool LN>
__device__ __forceinline__ void epi_pair_bf16(const EpiCtx& cx, const GemmEpi& ep, const CUtensorMap* tmC) {
  constexpr int CH = 32, MAXC = BN / CH / 2;
  const int quad = cx.warp & 3, half = cx.warp >> 2, lane = cx.lane;
  float* bias_s = reinterpret_cast<float*>(cx.stg4) + 32 * 16;      // +2048 B: after staging tile 0
  float* csum_s = bias_s + MAXC * CH;
  float2* ln_stage = reinterpret_cast<float2*>(cx.stg4 + 384);      // +6144 B: [slot][lane]
  auto ln_fetch = [&](int t) {
    const int g = ((t / cx.n_tiles) * cx.mt_mul + cx.mt_add) * 2 * GEMM_BM + cx.cta_rank * GEMM_BM + quad * 32 + lane;
    const float2* sp = reinterpret_cast<const float2*>(ep.ln_stats) + g;
    for (int q = 0; q < ep.ln_slots; ++q) {
      if (g < cx.M) cp_async_8(smem_u32(ln_stage + q * 32 + lane), sp + static_cast<size_t>(q) * ep.ln_stride);
      else ln_stage[q * 32 + lane] = make_float2(0.f, 0.f);
    }
    cp_async_commit();
  };
  if (LN && cx.worker < cx.num_tiles) ln_fetch(cx.worker);
  uint32_t bk = 0;
  int local = 0;
  for (int tile = cx.worker; tile < cx.num_tiles; tile += cx.n_workers, ++local) {
    const int as = local & 1;
    const uint32_t aphase = (local >> 1) & 1u;
    const int n0 = (tile % cx.n_tiles) * BN;
    const int r0 = ((tile / cx.n_tiles) * cx.mt_mul + cx.mt_add) * 2 * GEMM_BM + cx.cta_rank * GEMM_BM + quad * 32;
    __syncwarp();
    if (lane < MAXC * (CH / 4)) {          // this warp's bias (and column-sum) slices: chunk slot k = lane / 8
      const int nbk = n0 + (half + 2 * (lane >> 3)) * CH;
      reinterpret_cast<float4*>(bias_s)[lane] = __ldg(reinterpret_cast<const float4*>(ep.bias + nbk) + (lane & 7));
      if (LN) reinterpret_cast<float4*>(csum_s)[lane] = __ldg(reinterpret_cast<const float4*>(ep.colsum + nbk) + (lane & 7));
    }
    __syncwarp();
    float mu = 0.f, rs = 1.f;
    if (LN) {
      cp_async_wait_all();
      float s1 = 0.f, s2 = 0.f;
      for (int q = 0; q < ep.ln_slots; ++q) {       // fixed slot order: deterministic
        const float2 t = ln_stage[q * 32 + lane];
        s1 += t.x; s2 += t.y;
      }
      mu = s1 * ep.ln_inv_k;
      rs = rsqrtf(fmaxf(fmaf(-mu, mu, s2 * ep.ln_inv_k), 0.f) + ep.ln_eps);
      if (tile + cx.n_workers < cx.num_tiles) ln_fetch(tile + cx.n_workers);
    }
    const float nmu = -mu;
    mbar_wait(cx.tfull0 + 8u * as, aphase);
    tc_fence_after();
    const uint32_t t_row = cx.tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * BN);
#pragma unroll 1
    for (int kc = 0; kc < MAXC; ++kc) {
      const int c = half + 2 * kc;
      uint32_t v[32];
      tmem_ld_32x32(t_row + c * CH, v);
      tmem_ld_wait();
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // packed fp32 arithmetic (fma.rn.f32x2 / add.f32x2): the same operations per element, half the issue slots
        const float4 b = reinterpret_cast<const float4*>(bias_s + kc * CH)[j];      // broadcast reads
        float2 xa = make_float2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]));
        float2 xb = make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        if (LN) {
          const float4 cs = reinterpret_cast<const float4*>(csum_s + kc * CH)[j];
          const float2 rs2 = make_float2(rs, rs), nmu2 = make_float2(nmu, nmu);
          xa = __ffma2_rn(rs2, __ffma2_rn(nmu2, make_float2(cs.x, cs.y), xa), make_float2(b.x, b.y));
          xb = __ffma2_rn(rs2, __ffma2_rn(nmu2, make_float2(cs.z, cs.w), xb), make_float2(b.z, b.w));
        } else {
          xa = __fadd2_rn(xa, make_float2(b.x, b.y));
          xb = __fadd2_rn(xb, make_float2(b.z, b.w));
        }
        if (ACT == MMT_ACT_GELU) { xa = gelu_fast2(xa); xb = gelu_fast2(xb); }
        else if (ACT == MMT_ACT_RELU) { xa.x = fmaxf(xa.x, 0.f); xa.y = fmaxf(xa.y, 0.f); xb.x = fmaxf(xb.x, 0.f); xb.y = fmaxf(xb.y, 0.f); }
        w[2 * j] = pack_bf16x2(xa.x, xa.y);
        w[2 * j + 1] = pack_bf16x2(xb.x, xb.y);
      }
      // two staging tiles per warp: only the store of TWO chunks ago must have read its tile
      uint4* stg = cx.stg4 + 256 * (bk++ & 1u);
      if (lane == 0) bulk_wait_read_1();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j)       // 64-byte rows, 16-byte chunk j of row `lane` at j ^ ((lane >> 1) & 3): TMA SWIZZLE_64B
        stg[lane * 4 + (j ^ ((lane >> 1) & 3))] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
      fence_proxy_async_shared();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmC, smem_u32(stg), n0 + c * CH, r0);       // rows beyond M are clipped by the TMA
        bulk_commit_group();
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster((cx.tempty0 + 8u * as) & kPeerBitMask);     // the leader's barrier counts both CTAs
  }
  if (lane == 0) bulk_wait_all();
}

template <int BN, bool LNOUT>
__device__ __forceinline__ void epi_pair_f32(const EpiCtx& cx, const GemmEpi& ep, const CUtensorMap* tmC,
                                             const CUtensorMap* tmR, const CUtensorMap* tmX) {
  constexpr int CH = 32, MAXC = BN / CH / 2;
  const int quad = cx.warp & 3, half = cx.warp >> 2, lane = cx.lane;
  uint4* xtile = cx.stg4 + 512;        // bf16 shadow tile (+8192 B)
  uint32_t rk = 0;
  int local = 0;
  for (int tile = cx.worker; tile < cx.num_tiles; tile += cx.n_workers, ++local) {
    const int as = local & 1;
    const uint32_t aphase = (local >> 1) & 1u;
    const int n0 = (tile % cx.n_tiles) * BN;
    const int r0 = ((tile / cx.n_tiles) * cx.mt_mul + cx.mt_add) * 2 * GEMM_BM + cx.cta_rank * GEMM_BM + quad * 32;
    auto issue_resid = [&](int c, uint32_t k) {     // TMA load of the residual tile of chunk c into buffer k & 1
      if (lane == 0) {
        bulk_wait_read_all();                       // the store that last used this buffer has read it
        const uint32_t bar = cx.rbar0 + 8u * (k & 1u);
        mbar_expect_tx(bar, 4096);
        tma_load_2d(smem_u32(cx.stg4) + 4096u * (k & 1u), tmR, bar, n0 + c * CH, r0);
      }
    };
    issue_resid(half, rk);
    mbar_wait(cx.tfull0 + 8u * as, aphase);
    tc_fence_after();
    const uint32_t t_row = cx.tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * BN);
    float st_s1 = 0.f, st_s2 = 0.f;
#pragma unroll 1
    for (int kc = 0; kc < MAXC; ++kc) {
      const int c = half + 2 * kc;
      const int nb = n0 + c * CH;
      uint32_t v[32];
      tmem_ld_32x32(t_row + c * CH, v);
      const uint32_t k = rk++;
      if (kc + 1 < MAXC) issue_resid(c + 2, k + 1);
      uint4* tile_s = cx.stg4 + 256 * (k & 1u);
      float4 bq[8];                                  // the chunk's 32 bias values: 8 warp-uniform (broadcast) L1 loads
#pragma unroll
      for (int j = 0; j < 8; ++j) bq[j] = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + nb) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      tmem_ld_wait();
      mbar_wait(cx.rbar0 + 8u * (k & 1u), (k >> 1) & 1u);
      uint32_t xbp[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int slot = lane * 8 + (j ^ (lane & 7));      // 128-byte rows, SWIZZLE_128B
        const uint4 rw = tile_s[slot];
        const float4 b = bq[j];
        // (accumulator + bias) + residual, two lanes per instruction (add.f32x2: same rounding as the scalar adds)
        const float2 ya = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])),
                                                make_float2(b.x, b.y)),
                                     make_float2(__uint_as_float(rw.x), __uint_as_float(rw.y)));
        const float2 yb = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])),
                                                make_float2(b.z, b.w)),
                                     make_float2(__uint_as_float(rw.z), __uint_as_float(rw.w)));
        const float y0 = ya.x, y1 = ya.y, y2 = yb.x, y3 = yb.y;
        tile_s[slot] = make_uint4(__float_as_uint(y0), __float_as_uint(y1), __float_as_uint(y2), __float_as_uint(y3));
        if (LNOUT) {
          st_s1 += (y0 + y1) + (y2 + y3);
          st_s2 = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, st_s2))));
          xbp[2 * j] = pack_bf16x2(y0, y1);
          xbp[2 * j + 1] = pack_bf16x2(y2, y3);
        }
      }
      if (LNOUT) {
        if (lane == 0) bulk_wait_read_all();         // the previous chunk's shadow-tile store has read it
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          xtile[lane * 4 + (j ^ ((lane >> 1) & 3))] = make_uint4(xbp[4 * j], xbp[4 * j + 1], xbp[4 * j + 2], xbp[4 * j + 3]);
      }
      fence_proxy_async_shared();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmC, smem_u32(tile_s), nb, r0);
        if (LNOUT) tma_store_2d(tmX, smem_u32(xtile), nb, r0);
        bulk_commit_group();
      }
    }
    if (LNOUT && r0 + lane < cx.M) {
      // slot = (256-column tile, epilogue half): every (slot, row) is written by exactly one thread of the grid
      const int slot = (n0 / BN) * (BN / 128) + half;
      reinterpret_cast<float2*>(ep.stats_out)[static_cast<size_t>(slot) * ep.stats_stride + r0 + lane] = make_float2(st_s1, st_s2);
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster((cx.tempty0 + 8u * as) & kPeerBitMask);
  }
  if (lane == 0) bulk_wait_all();
}

// 320 threads are allocated as 384 (warps come in groups of four), so a thread may hold 65 536 / 384 = 170 -> 168 registers:
// that is what __launch_bounds__(320, 1) makes ptxas target; a higher cap (__maxnreg__) compiles but the launch is refused
// (cudaErrorLaunchOutOfResources).  The few spilled words are the epilogue's one-tile-ahead prefetch registers (written and
// read once per tile, L1-resident).
// CL4 (PAIR only): a cluster of FOUR CTAs = two CTA pairs that work on two 256-row blocks of the SAME 256-column tile.  Every
// CTA still loads its own 128 A rows, but only a QUARTER of the W tile, multicast to the CTA of the other pair that needs the
// same half: a k-slice costs the L2 96 KB per two tiles instead of 128 KB.  An EXPERIMENT (off by default, see
// g_cluster4_enabled): the backbone GEMMs deliver operands at ~10 TB/s x 128 FLOP/B = 1.28 PFLOP/s (qkv 1.29, fc1 1.25), which
// looked like the L2 read rate being the bound - the measurement says otherwise.  Stage reuse is gated on BOTH pairs (empty
// barriers count two commits, multicast to all four CTAs); accumulator barriers stay pair-local.
template <int BN, bool PAIR, bool CL4 = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                         const __grid_constant__ CUtensorMap tmX, int M, int N, int K, GemmEpi ep, GemmConv cv) {
  static_assert(PAIR || !CL4, "clusters of four are two CTA pairs");
  using Cfg = GemmCfg<BN, PAIR>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "shared memory budget");
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;         // rank in the cluster (0..1, or 0..3 with CL4)
  const uint32_t cta_rank = crank & 1u;                         // rank in the CTA pair: 0 = leader
  const int cpair = CL4 ? static_cast<int>(crank >> 1) : 0;     // which pair of the cluster
  constexpr int CSZ = CL4 ? 4 : 2;
  const int worker = PAIR ? (blockIdx.x / CSZ) : blockIdx.x;    // tile-scheduler slot (a cluster is one worker)
  const int n_workers = PAIR ? (gridDim.x / CSZ) : gridDim.x;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if (PAIR && (smem_u32(smem_raw) & 1023u) != 0) __trap();     // 128B-swizzle tiles need 1024-byte alignment
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = staging_base + Cfg::EPI_BYTES;
  // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * Cfg::STAGES + 4);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_smem - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Role -> warp id.  The SM's warp arbiter prefers the highest warp id among eligible warps, so the two
  // single-thread critical roles (TMA producer, MMA issuer) sit ABOVE the eight epilogue warps: with the roles the
  // other way round the epilogue's instruction stream delayed every MMA / TMA issue and cost 17-27 % of the GEMM
  // rate (measured with MMT_GEMM_DBG=1, profiles/r1_gemm_bound.md).
  constexpr int W_PRODUCER = GEMM_EPI_WARPS, W_MMA = GEMM_EPI_WARPS + 1;
  const int n_tiles = (N + BN - 1) / BN;
  const int tiles_per_img = cv.tiles_x * cv.tiles_y;
  constexpr int TILE_M = PAIR ? 2 * GEMM_BM : GEMM_BM;
  const int m_blocks = cv.enabled ? (M / (cv.H * cv.W)) * tiles_per_img : (M + TILE_M - 1) / TILE_M;
  const int m_tiles = CL4 ? (m_blocks + 1) / 2 : m_blocks;      // CL4: a scheduler tile = two 256-row blocks x one column tile
  const int num_tiles = n_tiles * m_tiles;
  constexpr int MT_MUL = CL4 ? 2 : 1;
  const int num_kb = cv.enabled ? 9 * cv.cchunks : (K + GEMM_BK - 1) / GEMM_BK;
  const int dbg_flags = kDev ? ep.dbg_flags : 0;          // compile-time 0 in the shipped build
  long long* const dbg = kDev ? ep.dbg : nullptr;

  if (warp == W_PRODUCER && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CL4 ? 2 : 1);       // CL4: the other pair's CTA writes into this stage too
    }
    if (PAIR)
      for (int i = 0; i < 2 * GEMM_EPI_WARPS; ++i) mbar_init(bar_base + 256u + 8u * i, 1);   // residual tiles landed
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), PAIR ? 2 * GEMM_EPI_WARPS : GEMM_EPI_WARPS);   // PAIR: both CTAs' epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) {
    if (PAIR) { tmem_alloc_pair(tmem_ptr_smem, Cfg::TMEM_COLS); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();    // barriers of BOTH CTAs are initialised before any remote arrive / multicast commit
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  // everything above touched only this CTA's shared / tensor memory: it may overlap the previous kernel of the stream
  pdl_wait();
  pdl_launch_dependents();

  if (warp == W_PRODUCER) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = worker; tile < num_tiles; tile += n_workers) {
        const int mt = (tile / n_tiles) * MT_MUL + cpair;
        const int m0 = mt * TILE_M + static_cast<int>(cta_rank) * GEMM_BM;
        const int n0 = (tile % n_tiles) * BN;
        if (CL4) {
          // this CTA: its 128 A rows (rows beyond M arrive as zeros: the odd last block of a cluster tile) and ONE QUARTER of
          // the W tile, multicast to the CTA of the other pair that holds the same half (ranks h and h + 2); per pair and
          // stage the leader's barrier still counts 2 x 16 KB of A + 4 x 8 KB of W
          const int nq0 = n0 + static_cast<int>(cta_rank) * (BN / 2) + cpair * (BN / 4);
          const uint16_t mc_mask = static_cast<uint16_t>((1u << cta_rank) | (1u << (cta_rank + 2)));
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);          // BOTH pairs have consumed this stage
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
            const uint32_t a_dst = smem_base + stage * Cfg::STAGE_BYTES;
            tma_load_2d_pair(a_dst, &tmA, full_bar(stage), kb * GEMM_BK, m0);
            tma_load_2d_pair_mc(a_dst + Cfg::A_BYTES + cpair * (Cfg::B_BYTES / 2), &tmB, full_bar(stage), kb * GEMM_BK, nq0,
                                mc_mask);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
          }
        } else if (PAIR) {
          // each CTA: its 128 A rows and its half of the W rows; all four boxes are counted on the leader's barrier
          const int nb0 = n0 + static_cast<int>(cta_rank) * (BN / 2);
          const int next_tile = tile + n_workers;
          const int pm0 = (next_tile / n_tiles) * TILE_M + static_cast<int>(cta_rank) * GEMM_BM;
          const bool pf = ep.prefetch && next_tile < num_tiles && pm0 != m0;
          for (int kb = 0; kb < num_kb; ++kb) {
            if (pf) tma_prefetch_2d(&tmA, kb * GEMM_BK, pm0);     // next tile's A rows -> L2 (first touch is DRAM)
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
            const uint32_t a_dst = smem_base + stage * Cfg::STAGE_BYTES;
            const bool same = (dbg_flags & 2) != 0;
            tma_load_2d_pair(a_dst, &tmA, full_bar(stage), kb * GEMM_BK, same ? static_cast<int>(cta_rank) * GEMM_BM : m0);
            tma_load_2d_pair(a_dst + Cfg::A_BYTES, &tmB, full_bar(stage), kb * GEMM_BK,
                             same ? static_cast<int>(cta_rank) * (BN / 2) : nb0);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
          }
        } else if (!cv.enabled) {
          const int next_tile = tile + n_workers;
          const int pm0 = (next_tile / n_tiles) * TILE_M;
          const bool pf = ep.prefetch && next_tile < num_tiles && pm0 != m0;
          for (int kb = 0; kb < num_kb; ++kb) {
            if (pf) tma_prefetch_2d(&tmA, kb * GEMM_BK, pm0);     // next tile's A rows -> L2 (first touch is DRAM)
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
            const uint32_t a_dst = smem_base + stage * Cfg::STAGE_BYTES;
            tma_load_2d(a_dst, &tmA, full_bar(stage), kb * GEMM_BK, m0);
            tma_load_2d(a_dst + Cfg::A_BYTES, &tmB, full_bar(stage), kb * GEMM_BK, n0);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
          }
        } else {
          const int img = mt / tiles_per_img, bt = mt % tiles_per_img;
          const int x0 = (bt % cv.tiles_x) * cv.BW, y0 = (bt / cv.tiles_x) * cv.BH;
          const uint32_t a_bytes = static_cast<uint32_t>(cv.BW * cv.BH) * GEMM_BK * 2;
          for (int tap = 0; tap < 9; ++tap) {
            for (int cc = 0; cc < cv.cchunks; ++cc) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              mbar_expect_tx(full_bar(stage), a_bytes + Cfg::B_BYTES);
              const uint32_t a_dst = smem_base + stage * Cfg::STAGE_BYTES;
              tma_load_4d(a_dst, &tmA, full_bar(stage), cc * GEMM_BK, x0 + tap % 3 - 1, y0 + tap / 3 - 1, img);
              tma_load_2d(a_dst + Cfg::A_BYTES, &tmB, full_bar(stage), tap * cv.C + cc * GEMM_BK, n0);
              if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0 && cta_rank == 0) {     // PAIR: only the leader CTA issues (for both SMs)
      constexpr uint32_t idesc = make_idesc_bf16_f32(TILE_M, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      long long t_wait_acc = 0, t_wait_full = 0, t_total = dbg ? clock64() : 0;
      for (int tile = worker; tile < num_tiles; tile += n_workers, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1u;
        long long t0 = dbg ? clock64() : 0;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        if (dbg) t_wait_acc += clock64() - t0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          t0 = dbg ? clock64() : 0;
          mbar_wait(full_bar(stage), phase);
          if (dbg) t_wait_full += clock64() - t0;
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
          const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128B swizzle atom: +2 in 16-byte units
            if (PAIR) mma_bf16_ss_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else mma_bf16_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if (CL4) mma_commit_pair(empty_bar(stage), 0xF);  // one of the two commits that free the slot in all four CTAs
          else if (PAIR) mma_commit_pair(empty_bar(stage)); // frees the slot in both CTAs
          else mma_commit(empty_bar(stage));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        if (PAIR) mma_commit_pair(tfull_bar(as), static_cast<uint16_t>(3u << (2 * cpair)));   // this pair's two CTAs: their
        else mma_commit(tfull_bar(as));                                                         // epilogues read their own lanes
      }
      if (dbg) {
        dbg[blockIdx.x * 4 + 0] = clock64() - t_total;
        dbg[blockIdx.x * 4 + 1] = t_wait_acc;     // MMA thread stalled on the epilogue (accumulator not drained)
        dbg[blockIdx.x * 4 + 2] = t_wait_full;    // MMA thread stalled on TMA data
        dbg[blockIdx.x * 4 + 3] = local;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 0..7)
    // Each warp owns 32 output rows (its TMEM lane quadrant) and walks them in CH-column chunks:
    //   1. tcgen05.ld the raw fp32 accumulators (thread == row) and write them to a per-warp, XOR-swizzled
    //      fp32 staging tile with 16-byte shared stores (conflict-free);
    //   2. read the tile back TRANSPOSED - a lane now owns a fixed group of 4 (fp32 out) or 8 (bf16 out)
    //      consecutive columns of several rows - so bias is ONE register vector per lane and chunk, and every
    //      HBM access (residual read, positional-table read, output store) is a 16-byte vector with a warp
    //      instruction covering whole row segments;
    //   3. bias, activation, positional add, residual add and the store happen in that layout.
    // The bias / residual vectors of the NEXT chunk are requested before the current chunk is processed, so their
    // latency hides behind the TMEM load and the math of the current one.
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int half = warp >> 2;         // which of the two warps sharing the quadrant
    uint4* stg4 = reinterpret_cast<uint4*>(smem_raw + (staging_base - smem_u32(smem_raw))) +
                  warp * (Cfg::EPI_TILE_BYTES / 16);
    const uint32_t rbar0 = bar_base + 256u + 16u * warp;          // this warp's two residual-tile barriers (PAIR)
    bool lean_done = false;
    if constexpr (PAIR) {
      // the backbone's shapes: every tile complete, plain GEMM, TMA epilogues available -> specialised loops (above)
      if (!cv.enabled && dbg_flags == 0 && (N % BN) == 0 && ep.vec_ok && !ep.rowadd) {
        const EpiCtx cx{tmem_base, tfull_bar(0), tempty_bar(0), rbar0, stg4, worker, n_workers, num_tiles, n_tiles, M, N,
                        warp, lane, static_cast<int>(cta_rank), MT_MUL, cpair};
        const bool ln = ep.ln_stats != nullptr;
        if (!ep.out_fp32 && ep.tma_store && ep.bias && (ep.act != MMT_ACT_RELU || !ln)) {
          if (ep.act == MMT_ACT_GELU) { if (ln) epi_pair_bf16<BN, MMT_ACT_GELU, true>(cx, ep, &tmC); else epi_pair_bf16<BN, MMT_ACT_GELU, false>(cx, ep, &tmC); }
          else if (ep.act == MMT_ACT_RELU) epi_pair_bf16<BN, MMT_ACT_RELU, false>(cx, ep, &tmC);
          else { if (ln) epi_pair_bf16<BN, MMT_ACT_NONE, true>(cx, ep, &tmC); else epi_pair_bf16<BN, MMT_ACT_NONE, false>(cx, ep, &tmC); }
          lean_done = true;
        } else if (ep.out_fp32 && ep.tma_f32 && ep.act == MMT_ACT_NONE) {
          if (ep.xb_out) epi_pair_f32<BN, true>(cx, ep, &tmC, &tmR, &tmX); else epi_pair_f32<BN, false>(cx, ep, &tmC, &tmR, &tmX);
          lean_done = true;
        }
      }
    }
    uint32_t rk = 0;                                               // running chunk counter of the TMA fp32 epilogue
    uint32_t bk = 0;                                               // running chunk counter of the TMA bf16 epilogue
    constexpr int CH = Cfg::CH;
    constexpr int NCH = BN / CH;
    constexpr int VPR = CH / 4;                 // 16-byte fp32 vectors per staged row (8 or 4)
    constexpr int KEYDIV = 8 / VPR;             // rows sharing a swizzle key
    constexpr int IT_F = VPR;                   // fp32 out: 32/VPR rows per instruction, VPR instructions
    constexpr int LPR_H = CH / 8;               // bf16 out: lanes per row (8 columns each)
    constexpr int IT_H = LPR_H;
    auto key_of = [](int r) { return (r / KEYDIV) & (VPR - 1); };
    // folded LayerNorm, consumer side.  CTA-pair kernel: the partial sums of this thread's row are fetched ONE TILE AHEAD with
    // cp.async into a per-warp staging area (2 KB at +6144 of the warp's epilogue region) and reduced when the tile starts, so
    // their L2 latency never sits on the epilogue's critical path.  (Parking them in registers instead made ptxas spill them
    // right after the loads - the spill stores then waited for the loads: 11 % of the kernel's stall samples,
    // profiles/r2_gemm_epilogue.md.)  Single-CTA kernel (small problems): plain loads at the start of the tile.
    const bool ln_active = !ep.out_fp32 && ep.vec_ok && ep.bias != nullptr && ep.ln_stats != nullptr && !cv.enabled;
    constexpr int LN_MAX_SLOTS = 8;
    float2* ln_stage = reinterpret_cast<float2*>(stg4 + 384);        // [slot][lane], CTA-pair kernel only
    auto ln_row_of = [&](int t) { return (t / n_tiles) * TILE_M + static_cast<int>(cta_rank) * GEMM_BM + (warp & 3) * 32 + lane; };
    auto ln_fetch = [&](int t) {
      if (!PAIR) return;
      const int g = ln_row_of(t);
      const float2* sp = reinterpret_cast<const float2*>(ep.ln_stats) + g;
      for (int q = 0; q < ep.ln_slots; ++q) {
        if (g < M) cp_async_8(smem_u32(ln_stage + q * 32 + lane), sp + static_cast<size_t>(q) * ep.ln_stride);
        else ln_stage[q * 32 + lane] = make_float2(0.f, 0.f);
      }
      cp_async_commit();
    };
    if (ln_active && worker < num_tiles && !lean_done) ln_fetch(worker);
    // (Measured and dropped: requesting a warp's residual tiles into L2 one tile ahead with cp.async.bulk.prefetch made proj
    // 50 -> 56 us and fc2 120 -> 130 us at M = 28 928 - like the A-operand prefetch of round 1, profiles/r2_gemm_epilogue.md.)
    int local = 0;
    if (CL4 && !lean_done) __trap();      // the cluster-of-four launch is only made for the lean epilogues' shapes
    for (int tile = (lean_done || CL4) ? num_tiles : worker; tile < num_tiles; tile += n_workers, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1u;
      const int mt = tile / n_tiles;
      const int n0 = (tile % n_tiles) * BN;
      // global output row of tile-local row i (or -1): plain GEMM rows are contiguous; conv rows are the pixels
      // of the tile's BW x BH box that fall inside the image
      int cimg = 0, cx0 = 0, cy0 = 0;
      if (cv.enabled) {
        cimg = mt / tiles_per_img;
        const int bt = mt % tiles_per_img;
        cx0 = (bt % cv.tiles_x) * cv.BW;
        cy0 = (bt / cv.tiles_x) * cv.BH;
      }
      auto grow_of = [&](int i) -> int {
        if (!cv.enabled) {
          const int g = mt * TILE_M + static_cast<int>(cta_rank) * GEMM_BM + i;
          return g < M ? g : -1;
        }
        const int by = i / cv.BW, bx = i - by * cv.BW;
        const int y = cy0 + by, x = cx0 + bx;
        return (by < cv.BH && y < cv.H && x < cv.W) ? (cimg * cv.H + y) * cv.W + x : -1;
      };
      const int lrow0 = quad * 32;
      const int rr_f = lane / VPR, cc_f = lane % VPR;          // fp32-out lane -> (row in group, vector)
      const int rr_h = lane / LPR_H, cc_h = lane % LPR_H;      // bf16-out lane -> (row in group, 8-column group)
      // per-chunk prefetch registers
      float4 rv[IT_F];          // residual vectors (fp32 out only)
      float4 bv[2];             // bias: [0] = this lane's 4 columns (fp32 out) / first 4 of its 8 (bf16), [1] = last 4
#pragma unroll
      for (int it = 0; it < IT_F; ++it) rv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      bv[0] = bv[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      auto prefetch = [&](int c) {
        const int nb = n0 + c * CH;
        const bool full = ep.vec_ok && (nb + CH <= N);
        if (!full) return;
        if (ep.bias) {
          if (ep.out_fp32) bv[0] = __ldg(reinterpret_cast<const float4*>(ep.bias + nb) + cc_f);
        }
        if (ep.resid && !(dbg_flags & 4) && !(PAIR && ep.tma_f32 && !cv.enabled)) {
#pragma unroll
          for (int it = 0; it < IT_F; ++it) {
            const int grow = grow_of(lrow0 + it * (32 / VPR) + rr_f);
            if (grow >= 0)
              rv[it] = __ldcs(reinterpret_cast<const float4*>(ep.resid + static_cast<size_t>(grow) * ep.ldr + nb) + cc_f);
          }
        }
      };
      prefetch(half);
      // bf16 outputs: bias/activation happen BEFORE the transpose (thread == row) so that the staging tile holds
      // packed bf16 - half the shared-memory traffic of an fp32 tile, and shared-memory wavefronts (epilogue staging
      // + TMA fills) are what bounds this kernel (profiles/r1_gemm_bound.md).  thread == row needs every bias value
      // of a chunk, so the warp's bias slices (<= 4 chunks x CH floats) are parked in the upper part of its staging
      // area once per tile and read back as broadcast vectors.
      constexpr int MAXC = (NCH + 1) / 2;           // chunks per warp
      // (per WARP, never shared between warps: the epilogue warps of a CTA drift apart by up to a tile)
      float* bias_s = reinterpret_cast<float*>(stg4) + 32 * 16;   // after the 32 x 64 B bf16 staging rows
      const bool f32_tma = PAIR && CH == 32 && ep.tma_f32 && !cv.enabled;
      const bool bias_smem = !ep.out_fp32 && ep.vec_ok && ep.bias != nullptr;
      // folded LayerNorm (consumer side): column sums of the folded weight beside the bias slices, and this thread's row
      // statistics from the partial sums its producer left (fixed summation order -> deterministic)
      const bool ln_in = ln_active;
      float* csum_s = bias_s + MAXC * CH;
      float ln_mu = 0.f, ln_rs = 1.f;
      if (bias_smem) {
        __syncwarp();
        if (lane < MAXC * (CH / 4)) {
          const int k = lane / (CH / 4), vq = lane % (CH / 4);
          const int nbk = n0 + (half + 2 * k) * CH;
          float4 bvv = make_float4(0.f, 0.f, 0.f, 0.f), cvv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (half + 2 * k < NCH && nbk + CH <= N) {
            bvv = __ldg(reinterpret_cast<const float4*>(ep.bias + nbk) + vq);
            if (ln_in) cvv = __ldg(reinterpret_cast<const float4*>(ep.colsum + nbk) + vq);
          }
          reinterpret_cast<float4*>(bias_s)[lane] = bvv;
          if (ln_in) reinterpret_cast<float4*>(csum_s)[lane] = cvv;
        }
        __syncwarp();
      }
      // folded LayerNorm (producer side): running partial sums of this thread's output row over the chunks it owns
      float st_s1 = 0.f, st_s2 = 0.f;
      if (ln_in) {
        float s1 = 0.f, s2 = 0.f;
        if (PAIR) {
          cp_async_wait_all();                                 // this thread's own copies (issued one tile ago)
          for (int q = 0; q < ep.ln_slots; ++q) {              // fixed slot order: deterministic
            const float2 t = ln_stage[q * 32 + lane];
            s1 += t.x; s2 += t.y;
          }
          if (tile + n_workers < num_tiles) ln_fetch(tile + n_workers);      // consumed at the start of the next tile
        } else {
          const int g = ln_row_of(tile);
          if (g < M) {
            const float2* sp = reinterpret_cast<const float2*>(ep.ln_stats) + g;
            for (int q = 0; q < ep.ln_slots; ++q) {
              const float2 t = __ldg(sp + static_cast<size_t>(q) * ep.ln_stride);
              s1 += t.x; s2 += t.y;
            }
          }
        }
        ln_mu = s1 * ep.ln_inv_k;
        ln_rs = rsqrtf(fmaxf(fmaf(-ln_mu, ln_mu, s2 * ep.ln_inv_k), 0.f) + ep.ln_eps);
        if (dbg_flags & 64) { ln_mu = 0.f; ln_rs = 1.f; }
      }

      const int grow_w = mt * TILE_M + static_cast<int>(cta_rank) * GEMM_BM + lrow0;   // first global row of this warp
      auto issue_resid = [&](int c, uint32_t k) {     // TMA load of the residual tile of chunk c into buffer k & 1
        if (lane == 0) {
          bulk_wait_read_all();                       // the store that last used this buffer has read it
          const uint32_t bar = rbar0 + 8u * (k & 1u);
          mbar_expect_tx(bar, 4096);
          tma_load_2d(smem_u32(stg4) + 4096u * (k & 1u), &tmR, bar, n0 + c * CH, grow_w);
        }
      };
      if (f32_tma && half < NCH && n0 + half * CH + CH <= N) issue_resid(half, rk);
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * BN);
#pragma unroll 1
      for (int c = half; c < NCH; c += 2) {
        uint32_t v[32];
        if (CH == 32) tmem_ld_32x32(t_row + c * CH, v);
        else tmem_ld_32x16(t_row + c * CH, v);
        float4 rcur[IT_F];
#pragma unroll
        for (int it = 0; it < IT_F; ++it) rcur[it] = rv[it];
        const float4 b0 = bv[0], b1 = bv[1];
        if (c + 2 < NCH) prefetch(c + 2);
        tmem_ld_wait();
        const int nb = n0 + c * CH;
        if (nb >= N || (dbg_flags & 1)) continue;
        const bool full = ep.vec_ok && (nb + CH <= N);
        if (full && f32_tma) {
          // ---- fp32 output + fp32 residual, CTA-pair kernel: thread == row throughout.  The residual tile was fetched
          // by TMA into this warp's buffer (128-byte rows, 128B swizzle); the row is updated in place and the tile
          // leaves by one TMA store - no per-lane global access, full 128-byte lines, loads one chunk ahead.
          const uint32_t k = rk++;
          if (c + 2 < NCH && n0 + (c + 2) * CH + CH <= N) issue_resid(c + 2, k + 1);
          uint4* tile = stg4 + 256 * (k & 1u);                       // 4096 B per buffer
          // bias: both buffers are in use, so the chunk's 32 values come as 8 warp-uniform (broadcast) L1 loads
          float4 bq[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            bq[j] = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + nb) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          mbar_wait(rbar0 + 8u * (k & 1u), (k >> 1) & 1u);
          const bool ln_out = ep.xb_out != nullptr;     // bf16 copy + partial sums for the LayerNorm folded downstream
          uint32_t xbp[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int slot = lane * 8 + (j ^ (lane & 7));
            const uint4 rw = tile[slot];
            const float4 b = bq[j];
            float x[4] = {__uint_as_float(v[4 * j]) + b.x, __uint_as_float(v[4 * j + 1]) + b.y,
                          __uint_as_float(v[4 * j + 2]) + b.z, __uint_as_float(v[4 * j + 3]) + b.w};
            if (ep.act == MMT_ACT_GELU) {
#pragma unroll
              for (int q = 0; q < 4; ++q) x[q] = gelu_fast(x[q]);
            } else if (ep.act == MMT_ACT_RELU) {
#pragma unroll
              for (int q = 0; q < 4; ++q) x[q] = fmaxf(x[q], 0.f);
            }
            const float y0 = x[0] + __uint_as_float(rw.x), y1 = x[1] + __uint_as_float(rw.y);
            const float y2 = x[2] + __uint_as_float(rw.z), y3 = x[3] + __uint_as_float(rw.w);
            uint4 o;
            o.x = __float_as_uint(y0); o.y = __float_as_uint(y1); o.z = __float_as_uint(y2); o.w = __float_as_uint(y3);
            tile[slot] = o;
            if (ln_out) {
              st_s1 += (y0 + y1) + (y2 + y3);
              st_s2 = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, st_s2))));
              xbp[2 * j] = pack_bf16x2(y0, y1);
              xbp[2 * j + 1] = pack_bf16x2(y2, y3);
            }
          }
          uint4* xtile = stg4 + 512;                     // bf16 shadow tile: 32 rows x 64 B, 64-byte swizzle (after the 2 x 4 KB)
          if (ln_out) {
            if (lane == 0) bulk_wait_read_all();         // the previous chunk's shadow-tile store has read it
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j)
              xtile[lane * 4 + (j ^ ((lane >> 1) & 3))] = make_uint4(xbp[4 * j], xbp[4 * j + 1], xbp[4 * j + 2], xbp[4 * j + 3]);
          }
          fence_proxy_async_shared();
          __syncwarp();
          if (lane == 0) {
            if (dbg_flags & 16) tma_store_2d_hint(&tmC, smem_u32(tile), nb, grow_w, l2_policy_evict_last());
            else tma_store_2d(&tmC, smem_u32(tile), nb, grow_w);
            if (ln_out && !(dbg_flags & 256)) tma_store_2d(&tmX, smem_u32(xtile), nb, grow_w);
            bulk_commit_group();
          }
          continue;
        }
        if (full && ep.out_fp32) {
          // ---- fp32 output: stage raw accumulators (row `lane`, vector j at slot j ^ key(lane)), finish transposed
#pragma unroll
          for (int j = 0; j < VPR; ++j)
            stg4[lane * VPR + (j ^ key_of(lane))] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
          {
            float* obase = reinterpret_cast<float*>(ep.out) + nb;
#pragma unroll
            for (int it = 0; it < IT_F; ++it) {
              const int r = it * (32 / VPR) + rr_f;
              const int grow = grow_of(lrow0 + r);
              const uint4 w = stg4[r * VPR + (cc_f ^ key_of(r))];
              float x[4] = {__uint_as_float(w.x) + b0.x, __uint_as_float(w.y) + b0.y, __uint_as_float(w.z) + b0.z,
                            __uint_as_float(w.w) + b0.w};
              if (ep.act == MMT_ACT_GELU) {
#pragma unroll
                for (int k = 0; k < 4; ++k) x[k] = gelu_fast(x[k]);
              } else if (ep.act == MMT_ACT_RELU) {
#pragma unroll
                for (int k = 0; k < 4; ++k) x[k] = fmaxf(x[k], 0.f);
              }
              if (grow >= 0 && !(dbg_flags & 4)) {
                if (ep.rowadd) {
                  const float4 p = __ldg(reinterpret_cast<const float4*>(
                                             ep.rowadd + static_cast<size_t>(grow % ep.rowadd_period) * N + nb) + cc_f);
                  x[0] += p.x; x[1] += p.y; x[2] += p.z; x[3] += p.w;
                }
                float4 o;
                o.x = x[0] + rcur[it].x; o.y = x[1] + rcur[it].y; o.z = x[2] + rcur[it].z; o.w = x[3] + rcur[it].w;
                reinterpret_cast<float4*>(obase + static_cast<size_t>(grow) * ep.ldo)[cc_f] = o;
              }
            }
          }
          __syncwarp();
        } else if (full) {
          // ---- bf16 output: bias + activation (+ positional add) with thread == row, pack, stage bf16, store transposed
          const int kc = (c - half) >> 1;                      // this warp's chunk slot -> its bias slice
          float f[CH];
#pragma unroll
          for (int j = 0; j < CH; ++j) f[j] = __uint_as_float(v[j]);
          if (ln_in) {
            const float nmu = -ln_mu;
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
              const float4 b = reinterpret_cast<const float4*>(bias_s + kc * CH)[j >> 2];   // broadcast reads
              const float4 c = (dbg_flags & 128) ? make_float4(0.f, 0.f, 0.f, 0.f)
                                                 : reinterpret_cast<const float4*>(csum_s + kc * CH)[j >> 2];
              f[j] = fmaf(ln_rs, fmaf(nmu, c.x, f[j]), b.x);
              f[j + 1] = fmaf(ln_rs, fmaf(nmu, c.y, f[j + 1]), b.y);
              f[j + 2] = fmaf(ln_rs, fmaf(nmu, c.z, f[j + 2]), b.z);
              f[j + 3] = fmaf(ln_rs, fmaf(nmu, c.w, f[j + 3]), b.w);
            }
          } else if (ep.bias) {
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
              const float4 b = reinterpret_cast<const float4*>(bias_s + kc * CH)[j >> 2];   // broadcast read
              f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
            }
          }
          if (ep.act == MMT_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < CH; ++j) f[j] = gelu_fast(f[j]);
          } else if (ep.act == MMT_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < CH; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (ep.rowadd) {
            const int grow = grow_of(lrow0 + lane);
            if (grow >= 0) {
              const float4* pr = reinterpret_cast<const float4*>(
                  ep.rowadd + static_cast<size_t>(grow % ep.rowadd_period) * N + nb);
#pragma unroll
              for (int j = 0; j < CH; j += 4) {
                const float4 p = __ldg(pr + (j >> 2));
                f[j] += p.x; f[j + 1] += p.y; f[j + 2] += p.z; f[j + 3] += p.w;
              }
            }
          }
          // staging rows are CH * 2 bytes = LPR_H 16-byte vectors; vector j of row `lane` at slot j ^ key
          // (for CH = 32 this is exactly the TMA SWIZZLE_64B pattern: 16-byte chunk ^ ((row >> 1) & 3))
          constexpr int KD_H = 8 / LPR_H;
          if (dbg_flags & 8) {
            // experiment: no shared-memory staging at all - thread == row, CH bf16 = CH/8 16-byte stores per thread
            const int grow = grow_of(lrow0 + lane);
            if (grow >= 0) {
              uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(ep.out) + static_cast<size_t>(grow) * ep.ldo + nb);
#pragma unroll
              for (int j = 0; j < LPR_H; ++j) {
                uint4 w;
                w.x = pack_bf16x2(f[8 * j], f[8 * j + 1]); w.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
                w.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]); w.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
                o[j] = w;
              }
            }
            continue;
          }
          const bool tma_out = CH == 32 && ep.tma_store && !cv.enabled;
          uint4* stg = stg4;
          if (tma_out) {
            // the bulk store that last used this staging tile must have read it.  CTA-pair kernel: two tiles per warp (the
            // second one in the area of the fp32 epilogue's buffers), so only the store of TWO chunks ago is waited for and
            // packing chunk k overlaps the store of chunk k - 1; single-CTA kernel: one tile
            if (PAIR && !(dbg_flags & 2048)) {
              stg = stg4 + 256 * (bk++ & 1u);              // 4096 B apart (bias / column-sum slices sit at +2048 .. +3072)
              if (lane == 0) bulk_wait_read_1();
            } else if (lane == 0) {
              bulk_wait_read_all();
            }
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < LPR_H; ++j) {
            uint4 w;
            w.x = pack_bf16x2(f[8 * j], f[8 * j + 1]); w.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
            w.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]); w.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
            stg[lane * LPR_H + (j ^ ((lane / KD_H) & (LPR_H - 1)))] = w;
          }
          if (tma_out) {
            // one bulk tensor store of the warp's 32 x 32 tile: no per-lane STG, rows beyond M are clipped by the TMA
            fence_proxy_async_shared();
            __syncwarp();
            if (lane == 0) {
              const int r0 = mt * TILE_M + static_cast<int>(cta_rank) * GEMM_BM + lrow0;
              if (dbg_flags & 32) tma_store_2d_hint(&tmC, smem_u32(stg), nb, r0, l2_policy_evict_first());
              else tma_store_2d(&tmC, smem_u32(stg), nb, r0);
              bulk_commit_group();
            }
            continue;
          }
          __syncwarp();
          bf16* obase = reinterpret_cast<bf16*>(ep.out) + nb;
#pragma unroll
          for (int it = 0; it < IT_H; ++it) {
            const int r = it * (32 / LPR_H) + rr_h;
            const int grow = grow_of(lrow0 + r);
            const uint4 w = stg4[r * LPR_H + (cc_h ^ ((r / KD_H) & (LPR_H - 1)))];
            if (grow >= 0 && !(dbg_flags & 4)) reinterpret_cast<uint4*>(obase + static_cast<size_t>(grow) * ep.ldo)[cc_h] = w;
          }
          __syncwarp();
        } else if (grow_of(lrow0 + lane) >= 0) {
          // ragged / unaligned tail: scalar, bounds-checked, thread == row
          const int row = grow_of(lrow0 + lane);
          const float* resid_row = ep.resid ? ep.resid + static_cast<size_t>(row) * ep.ldr : nullptr;
          const float* rowadd_row =
              ep.rowadd ? ep.rowadd + static_cast<size_t>(row % ep.rowadd_period) * N : nullptr;
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int n = nb + j;
            if (n >= N) continue;
            float x = __uint_as_float(v[j]);
            if (ep.bias) x += ep.bias[n];
            if (ep.act == MMT_ACT_GELU) x = gelu_fast(x);
            else if (ep.act == MMT_ACT_RELU) x = fmaxf(x, 0.f);
            if (rowadd_row) x += rowadd_row[n];
            if (resid_row) x += resid_row[n];
            if (ep.out_fp32) reinterpret_cast<float*>(ep.out)[static_cast<size_t>(row) * ep.ldo + n] = x;
            else reinterpret_cast<bf16*>(ep.out)[static_cast<size_t>(row) * ep.ldo + n] = __float2bfloat16_rn(x);
          }
        }
      }
      if (f32_tma && ep.stats_out != nullptr && grow_w + lane < M && !(dbg_flags & 512)) {
        // slot = (256-column tile, epilogue half): every (slot, row) is written by exactly one thread of the grid, the 32
        // rows of the warp as one 256-byte segment
        const int slot = (n0 / BN) * (BN / 128) + half;
        reinterpret_cast<float2*>(ep.stats_out)[static_cast<size_t>(slot) * ep.stats_stride + grow_w + lane] =
            make_float2(st_s1, st_s2);
      }
      // all tcgen05.ld of this warp have completed (wait::ld above): release the accumulator
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(tempty_bar(as) & kPeerBitMask);   // the leader's barrier counts both CTAs
        else mbar_arrive(tempty_bar(as));
      }
    }
    if (lane == 0) bulk_wait_all();      // outstanding bulk stores of this warp (smem must outlive their reads)
  }

  tc_fence_before();
  if (PAIR) {
    cluster_sync_all();            // no CTA leaves (or frees TMEM) while its peer may still touch it
    if (warp == W_MMA) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  } else {
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// Encoded tensor maps are cached per (kind, base pointer, shape, stride, box): the engine calls the same GEMMs on the same
// workspaces every frame, and cuTensorMapEncodeTiled costs about a microsecond of host time per map (3-4 maps per launch).
struct TmapKey {
  const void* ptr;
  long long a, b;      // packed (rows, cols) and (ld, box / kind)
  bool operator==(const TmapKey& o) const { return ptr == o.ptr && a == o.a && b == o.b; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    h ^= static_cast<size_t>(k.a) * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= static_cast<size_t>(k.b) * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    return h;
  }
};
static std::mutex g_tmap_mutex;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;

template <typename F>
static int cached_tmap(CUtensorMap* tm, int kind, const void* ptr, long long d0, long long d1, long long d2, long long d3,
                       F&& encode) {
  const TmapKey key{ptr, (d0 << 32) | (d1 & 0xffffffffll), (d2 << 32) | ((d3 & 0xffffffll) << 8) | kind};
  {
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) { *tm = it->second; return MMT_OK; }
  }
  const int rc = encode();
  if (rc == MMT_OK) {
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();      // bounded: workspaces are few, this is a safety net
    g_tmap_cache.emplace(key, *tm);
  }
  return rc;
}

// 2-D bf16 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows, 64].
// bf16 output [rows, cols] (leading dimension ld): 32 x 32 boxes, 64-byte swizzle (the epilogue's staging layout)
static int make_tmap_out_raw(CUtensorMap* tm, const void* ptr, int rows, int cols, int ld) {
  auto fn = get_encode_fn();
  if (!fn) return MMT_ERR_UNSUPPORTED;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MMT_OK : MMT_ERR_BAD_ARG;
}

static int make_tmap_out(CUtensorMap* tm, const void* ptr, int rows, int cols, int ld) {
  return cached_tmap(tm, 1, ptr, rows, cols, ld, 0, [&] { return make_tmap_out_raw(tm, ptr, rows, cols, ld); });
}

// fp32 [rows, cols] (leading dimension ld): 32 x 32 boxes = 128-byte rows, 128-byte swizzle (residual in / result out)
static int make_tmap_f32_tile_raw(CUtensorMap* tm, const void* ptr, int rows, int cols, int ld) {
  auto fn = get_encode_fn();
  if (!fn) return MMT_ERR_UNSUPPORTED;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MMT_OK : MMT_ERR_BAD_ARG;
}

static int make_tmap_f32_tile(CUtensorMap* tm, const void* ptr, int rows, int cols, int ld) {
  return cached_tmap(tm, 2, ptr, rows, cols, ld, 0, [&] { return make_tmap_f32_tile_raw(tm, ptr, rows, cols, ld); });
}

static int make_tmap_2d_raw(CUtensorMap* tm, const void* ptr, int rows, int cols, int ld, int box_rows) {
  auto fn = get_encode_fn();
  if (!fn) return MMT_ERR_UNSUPPORTED;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(GEMM_BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MMT_OK : MMT_ERR_BAD_ARG;
}

static int make_tmap_2d(CUtensorMap* tm, const void* ptr, int rows, int cols, int ld, int box_rows) {
  return cached_tmap(tm, 3, ptr, rows, cols, ld, box_rows, [&] { return make_tmap_2d_raw(tm, ptr, rows, cols, ld, box_rows); });
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_num_sms;
}

// CTA-pair launch (cluster of 2, cta_group::2 MMAs): plain GEMMs with N a multiple of 256 and enough 256-row tiles.
// CL4: clusters of four (two pairs sharing the W tile by TMA multicast) for the lean epilogues' shapes.
// Measured (profiles/r2_gemm_epilogue.md): correct, but NOT faster - qkv 82 -> 85.5 us, fc1 113.5 -> 118.5 us, fc2 121-126 -> 142 us
// at M = 28 928, neutral in the step: clusters of four leave SMs out (GPCs of 16-20 SMs) and every SM still ingests the same
// 32 KB per k-slice, so the bound is per SM, not the L2's read rate.  Off by default; mmt_config_cluster4(1) selects it.
int g_cluster4_enabled = 0;
// SM budget of the one-wave tile-narrowing rule (0 = the whole GPU).  When two independent chains of small GEMMs run on two
// streams (the two modality backbones at one sequence), a grid that covers every SM serialises them - every CTA of this kernel
// owns its SM's shared memory; with half the SMs as budget each launch picks the narrowest tile that fits HALF a wave and
// the two chains run side by side.  mmt_config_small_gemm_sms.
int g_small_gemm_sms = 0;

template <bool CL4>
static int launch_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmC, const CUtensorMap& tmR, const CUtensorMap& tmX,
                            const void* W, int ldw, int M, int N, int K, const GemmEpi& ep, cudaStream_t stream) {
  constexpr int BN = 256;
  constexpr int CSZ = CL4 ? 4 : 2;
  using Cfg = GemmCfg<BN, true>;
  CUtensorMap tmB;
  int rc = make_tmap_2d(&tmB, W, N, K, ldw, CL4 ? BN / 4 : BN / 2);      // CL4: every CTA loads a quarter of the W tile
  if (rc) return rc;
  auto kernel = gemm_bf16_tcgen05_kernel<BN, true, CL4>;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CSZ;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static bool attr_set = false;
  static int max_clusters = 0;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    // how many clusters fit at once (148 SMs are 8 GPCs of 16-20: clusters of four leave a few SMs out)
    cfg.gridDim = dim3(CSZ * (num_sms() / CSZ));
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg) != cudaSuccess || max_clusters <= 0) {
      cudaGetLastError();
      max_clusters = CL4 ? 0 : num_sms() / 2;
    }
    attr_set = true;
  }
  if (CL4 && max_clusters <= 0) return MMT_ERR_UNSUPPORTED;
  const int blocks256 = cdiv(M, 2 * GEMM_BM);
  const int tiles = (CL4 ? cdiv(blocks256, 2) : blocks256) * cdiv(N, BN);
  int clusters = max_clusters < num_sms() / CSZ ? max_clusters : num_sms() / CSZ;
  if (tiles < clusters) clusters = tiles;
  cfg.gridDim = dim3(CSZ * clusters);
  cfg.numAttrs = g_pdl_enabled ? 2 : 1;
  GemmConv cv = {};
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, tmC, tmR, tmX, M, N, K, ep, cv);
  if (e != cudaSuccess) return (int)e;
  MMT_RETURN_LAST_ERROR();
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmC, const void* W, int ldw, int M, int N, int K,
                       const GemmEpi& ep, const GemmConv& cv, int max_ctas, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap tmB;
  int rc = make_tmap_2d(&tmB, W, N, K, ldw, BN);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int m_tiles = cv.enabled ? (M / (cv.H * cv.W)) * cv.tiles_x * cv.tiles_y : cdiv(M, GEMM_BM);
  const int tiles = m_tiles * cdiv(N, BN);
  int grid = tiles < num_sms() ? tiles : num_sms();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  cudaError_t e = launch_pdl(gemm_bf16_tcgen05_kernel<BN, false>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, tmA,
                             tmB, tmC, tmC, tmC, M, N, K, ep, cv);
  if (e != cudaSuccess) return (int)e;
  MMT_RETURN_LAST_ERROR();
}

static int pick_bn(int N) {
  // widest tile whose padding waste is <= 1/8 of a tile, preferring fewer tiles
  const int cands[8] = {256, 192, 128, 96, 64, 48, 32, 16};
  int best = 16;
  long best_cost = -1;
  for (int i = 0; i < 8; ++i) {
    const int bn = cands[i];
    const long padded = static_cast<long>(cdiv(N, bn)) * bn;
    // cost = padded MMA columns, with a small penalty for narrow tiles (less operand reuse)
    const long cost = padded * 16 + (256 - bn);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

// A/B switches of the developer build (MMT_GEMM_DEV): environment variables read once.  The shipped build has none:
// every function below is a compile-time constant there.
static bool dev_switch(const char* name, bool dflt) {
  if (!kDev) return dflt;
  const char* e = getenv(name);
  return e ? (e[0] != '0') : dflt;
}
static bool prefetch_enabled() {      // L2 prefetch of the next tile's A rows: measured -4 % (profiles/r1_gemm_bound.md)
  static const bool v = dev_switch("MMT_GEMM_PREFETCH", false);
  return v;
}
long long* g_gemm_dbg = nullptr;      // MMT_GEMM_DEV builds: per-CTA cycle counters, see mmt_dev_gemm_timing (gemm_dev.cu)
static bool pair_enabled() {          // MMT_GEMM_PAIR=0: single-CTA kernel everywhere
  static const bool v = dev_switch("MMT_GEMM_PAIR", true);
  return v;
}
static bool tmastore_enabled() {      // MMT_GEMM_TMASTORE=0: per-lane store epilogue
  static const bool v = dev_switch("MMT_GEMM_TMASTORE", true);
  return v;
}
static bool narrow_enabled() {        // MMT_GEMM_NARROW=0: no tile narrowing for small grids
  static const bool v = dev_switch("MMT_GEMM_NARROW", true);
  return v;
}

// *ln_produced (optional) is set to 1 when the launch writes ep.xb_out / ep.stats_out itself (CTA-pair kernel with the TMA
// fp32 epilogue); otherwise the caller completes them with the row-statistics kernel.
static int dispatch_gemm(const CUtensorMap& tmA, const void* W, int ldw, int M, int N, int K, GemmEpi ep,
                         const GemmConv& cv, int max_ctas, cudaStream_t s, int* ln_produced = nullptr) {
  const int g_tmastore_mode = tmastore_enabled() ? 1 : 0;
  if (ln_produced) *ln_produced = 0;
  // bf16 output of a plain GEMM with 16-byte aligned rows: tiles leave through TMA stores
  CUtensorMap tmC = tmA;
  ep.tma_store = 0;
  ep.tma_f32 = 0;
  if (g_tmastore_mode && !cv.enabled && !ep.out_fp32 && ep.vec_ok && !ep.rowadd) {
    if (make_tmap_out(&tmC, ep.out, M, N, ep.ldo) == MMT_OK) ep.tma_store = 1;
    else tmC = tmA;
  }
  if (!cv.enabled && max_ctas <= 0 && (N % 256) == 0 && cdiv(M, 2 * GEMM_BM) * (N / 256) >= num_sms() / 2 &&
      pair_enabled()) {
    CUtensorMap tmR = tmA;
    ep.tma_f32 = 0;
    if (g_tmastore_mode && ep.out_fp32 && ep.resid && ep.vec_ok && !ep.rowadd && (ep.ldo % 4) == 0 && (ep.ldr % 4) == 0 &&
        make_tmap_f32_tile(&tmC, ep.out, M, N, ep.ldo) == MMT_OK &&
        make_tmap_f32_tile(&tmR, ep.resid, M, N, ep.ldr) == MMT_OK)
      ep.tma_f32 = 1;
    CUtensorMap tmX = tmA;
    if (ep.tma_f32 && ep.xb_out && make_tmap_out(&tmX, ep.xb_out, M, N, ep.ld_xb) != MMT_OK) ep.tma_f32 = 0;
    if (!ep.tma_f32) { ep.xb_out = nullptr; ep.stats_out = nullptr; }
    else if (ln_produced && ep.xb_out) *ln_produced = 1;
    // clusters of four for the shapes of the lean epilogues (the conditions of the kernel's own dispatch), with at least two
    // waves of cluster tiles
    const bool lean_bf16 = !ep.out_fp32 && ep.tma_store && ep.bias && (ep.act != MMT_ACT_RELU || !ep.ln_stats);
    const bool lean_f32 = ep.out_fp32 && ep.tma_f32 && ep.act == MMT_ACT_NONE;
    const bool dev_flags = kDev && ep.dbg_flags != 0;
    if (g_cluster4_enabled && ep.vec_ok && !ep.rowadd && (lean_bf16 || lean_f32) && !dev_flags &&
        cdiv(cdiv(M, 2 * GEMM_BM), 2) * (N / 256) >= 2 * (num_sms() / 4)) {
      const int rc4 = launch_gemm_pair<true>(tmA, tmC, tmR, tmX, W, ldw, M, N, K, ep, s);
      if (rc4 != MMT_ERR_UNSUPPORTED) return rc4;
    }
    return launch_gemm_pair<false>(tmA, tmC, tmR, tmX, W, ldw, M, N, K, ep, s);
  }
  ep.xb_out = nullptr;
  ep.stats_out = nullptr;
  int bn = pick_bn(N);
  // Small problems (bs = 1 latency: M = 452 rows per modality) leave most SMs idle with wide tiles - 4 x 3 tiles for
  // fc2 - and then the serial K loop of one CTA is the launch time.  Narrow the tile until the grid covers the machine
  // (never below 64 columns; only for plain GEMMs whose N the narrower tile divides).  MMT_GEMM_NARROW=0 disables (A/B).
  if (narrow_enabled() && max_ctas <= 0) {
    // (implicit-GEMM convolutions included: the head at one sequence is 3 pixel boxes x 7 tiles of 192 columns walking 108
    // k-slices each - 35 us - against 84 CTAs of 48 columns)
    const int m_tiles = cv.enabled ? (M / (cv.H * cv.W)) * cv.tiles_x * cv.tiles_y : cdiv(M, GEMM_BM);
    // ... but never past ONE wave: a narrower tile that needs a second round of CTAs doubles the launch time (fc1 at one
    // sequence: 192 tiles of 64 columns on 148 SMs took two rounds; 96 tiles of 128 columns take one)
    const int budget = (g_small_gemm_sms > 0 && g_small_gemm_sms < num_sms()) ? g_small_gemm_sms : num_sms();
    while (bn > 64 && (bn % 2) == 0 && (N % (bn / 2)) == 0 && m_tiles * cdiv(N, bn / 2) <= budget) bn /= 2;
  }
  switch (bn) {
    case 256: return launch_gemm<256>(tmA, tmC, W, ldw, M, N, K, ep, cv, max_ctas, s);
    case 192: return launch_gemm<192>(tmA, tmC, W, ldw, M, N, K, ep, cv, max_ctas, s);
    case 128: return launch_gemm<128>(tmA, tmC, W, ldw, M, N, K, ep, cv, max_ctas, s);
    case 96: return launch_gemm<96>(tmA, tmC, W, ldw, M, N, K, ep, cv, max_ctas, s);
    case 64: return launch_gemm<64>(tmA, tmC, W, ldw, M, N, K, ep, cv, max_ctas, s);
    case 48: return launch_gemm<48>(tmA, tmC, W, ldw, M, N, K, ep, cv, max_ctas, s);
    case 32: return launch_gemm<32>(tmA, tmC, W, ldw, M, N, K, ep, cv, max_ctas, s);
    default: return launch_gemm<16>(tmA, tmC, W, ldw, M, N, K, ep, cv, max_ctas, s);
  }
}

// NHWC map [B, H, W, C] (row stride ld elements) as a 4-D tensor; box = 64 channels x BW x BH pixels of one image.
static int make_tmap_nhwc_raw(CUtensorMap* tm, const void* ptr, int B, int H, int W, int C, int ld, int BW, int BH) {
  auto fn = get_encode_fn();
  if (!fn) return MMT_ERR_UNSUPPORTED;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(W) * ld * 2,
                        static_cast<cuuint64_t>(H) * W * ld * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(GEMM_BK), static_cast<cuuint32_t>(BW), static_cast<cuuint32_t>(BH), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MMT_OK : MMT_ERR_BAD_ARG;
}

static int make_tmap_nhwc(CUtensorMap* tm, const void* ptr, int B, int H, int W, int C, int ld, int BW, int BH) {
  // key: (B, H | W << 16), (C | ld << 20 ...) - all dimensions are small positive ints
  return cached_tmap(tm, 4, ptr, (static_cast<long long>(B) << 12) | BW, (static_cast<long long>(H) << 16) | W,
                     (static_cast<long long>(C) << 12) | BH, ld,
                     [&] { return make_tmap_nhwc_raw(tm, ptr, B, H, W, C, ld, BW, BH); });
}

// pixel box of <= 128 pixels that wastes the fewest MMA rows on an H x W map
static void pick_conv_box(int H, int W, int* BW, int* BH) {
  double best = -1.0;
  for (int bw = 1; bw <= W && bw <= 128; ++bw) {
    const int bh = 128 / bw < H ? 128 / bw : H;
    if (bh < 1) continue;
    const double eff = static_cast<double>(H) * W / (static_cast<double>(cdiv(W, bw)) * cdiv(H, bh) * 128.0);
    if (eff > best + 1e-9) { best = eff; *BW = bw; *BH = bh; }
  }
}

}  // namespace mmt

namespace mmt {
int launch_rowstats_cast(const float* x, int rows, int C, void* xb, int ld_xb, float* stats, int slots, int slot_stride,
                         cudaStream_t s);

static void init_epi(GemmEpi& ep) {
  std::memset(&ep, 0, sizeof(ep));
  ep.ln_inv_k = 0.f;
  ep.ln_eps = 0.f;
  if (kDev) {
    ep.dbg = g_gemm_dbg;
    static const int flags = [] { const char* e = getenv("MMT_GEMM_DBG"); return e ? atoi(e) : 0; }();
    ep.dbg_flags = flags;
  }
}
}  // namespace mmt

extern "C" int mmt_gemm_bf16_ex(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                                int act, const float* resid, int ldr, const float* rowadd, int rowadd_period, void* out,
                                int ldo, int out_fp32, int max_ctas, const float* ln_stats, int ln_slots, int ln_stride,
                                float ln_eps, const float* colsum, void* xb_out, int ld_xb, float* stats_out,
                                int stats_stride, void* stream) {
  using namespace mmt;
  MMT_CHECK_ARG(A && W && out && M > 0 && N > 0 && K > 0);
  MMT_CHECK_ARG(lda >= K && ldw >= K && ldo >= N);
  MMT_CHECK_ARG((lda % 8) == 0 && (ldw % 8) == 0);  // TMA: 16-byte global strides
  MMT_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0);
  MMT_CHECK_ARG(!rowadd || rowadd_period > 0);
  MMT_CHECK_ARG(!resid || (ldr >= N && out_fp32));  // the residual stream is fp32
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  GemmEpi ep;
  init_epi(ep);
  ep.bias = bias; ep.resid = resid; ep.rowadd = rowadd; ep.out = out;
  ep.ldr = ldr; ep.rowadd_period = rowadd_period; ep.ldo = ldo; ep.act = act; ep.out_fp32 = out_fp32;
  ep.prefetch = mmt::prefetch_enabled() ? 1 : 0;
  // fast path preconditions: 16-byte vector loads of bias / rowadd / resid and 16-byte vector stores
  ep.vec_ok = al16(out) && (out_fp32 ? (ldo % 4 == 0) : (ldo % 8 == 0)) && (!bias || al16(bias)) &&
              (!rowadd || ((N % 4 == 0) && al16(rowadd))) && (!resid || (al16(resid) && ldr % 4 == 0));
  if (ln_stats) {
    // folded LayerNorm, consumer side: bf16 output, per-column (bias, colsum) vectors, every column chunk complete
    MMT_CHECK_ARG(colsum && bias && !out_fp32 && !rowadd && ep.vec_ok && (N % 32) == 0);
    MMT_CHECK_ARG(ln_slots > 0 && ln_slots <= 8 && ln_stride >= M && (reinterpret_cast<uintptr_t>(ln_stats) & 7) == 0 &&
                  al16(colsum));
    ep.ln_stats = ln_stats; ep.colsum = colsum; ep.ln_slots = ln_slots; ep.ln_stride = ln_stride;
    ep.ln_inv_k = 1.0f / static_cast<float>(K);
    ep.ln_eps = ln_eps;
  }
  if (xb_out || stats_out) {
    // producer side: fp32 output rows (N a multiple of 128: one statistics slot per 128 columns) + bf16 copy + partial sums
    MMT_CHECK_ARG(xb_out && stats_out && out_fp32 && (N % 128) == 0 && ld_xb >= N && (ld_xb % 8) == 0 && al16(xb_out) &&
                  (reinterpret_cast<uintptr_t>(stats_out) & 7) == 0 && ldo == N && stats_stride >= M);
    ep.xb_out = static_cast<bf16*>(xb_out); ep.stats_out = stats_out; ep.ld_xb = ld_xb; ep.stats_stride = stats_stride;
  }
  CUtensorMap tmA;
  int rc = make_tmap_2d(&tmA, A, M, K, lda, GEMM_BM);
  if (rc) return rc;
  GemmConv cv = {};
  int produced = 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  rc = dispatch_gemm(tmA, W, ldw, M, N, K, ep, cv, max_ctas, s, &produced);
  if (rc) return rc;
  if (xb_out && !produced)      // launch shapes without the fused form: same outputs from the row-statistics kernel
    return launch_rowstats_cast(static_cast<const float*>(out), M, N, xb_out, ld_xb, stats_out, N / 128, stats_stride, s);
  return MMT_OK;
}

extern "C" int mmt_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                             int act, const float* resid, int ldr, const float* rowadd, int rowadd_period, void* out,
                             int ldo, int out_fp32, int max_ctas, void* stream) {
  return mmt_gemm_bf16_ex(A, lda, W, ldw, M, N, K, bias, act, resid, ldr, rowadd, rowadd_period, out, ldo, out_fp32,
                          max_ctas, nullptr, 0, 0, 0.f, nullptr, nullptr, 0, nullptr, 0, stream);
}

extern "C" int mmt_conv3x3_bf16(const void* in, int ld_in, int B, int H, int W, int C, const void* Wt, int ldw, int N,
                                const float* bias, int act, void* out, int ldo, int out_fp32, void* stream) {
  using namespace mmt;
  MMT_CHECK_ARG(in && Wt && out && B > 0 && H > 0 && W > 0 && C > 0 && N > 0);
  MMT_CHECK_ARG(ld_in >= C && (ld_in % 8) == 0 && ldw >= 9 * C && (ldw % 8) == 0 && ldo >= N);
  MMT_CHECK_ARG((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(Wt) & 15) == 0);
  GemmConv cv = {};
  cv.enabled = 1; cv.H = H; cv.W = W; cv.C = C;
  pick_conv_box(H, W, &cv.BW, &cv.BH);
  cv.tiles_x = cdiv(W, cv.BW); cv.tiles_y = cdiv(H, cv.BH); cv.cchunks = cdiv(C, GEMM_BK);
  GemmEpi ep;
  init_epi(ep);
  ep.dbg = nullptr;
  ep.dbg_flags = 0;
  ep.bias = bias; ep.out = out;
  ep.ldo = ldo; ep.act = act; ep.out_fp32 = out_fp32;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  ep.vec_ok = al16(out) && (out_fp32 ? (ldo % 4 == 0) : (ldo % 8 == 0)) && (!bias || al16(bias));
  CUtensorMap tmA;
  int rc = make_tmap_nhwc(&tmA, in, B, H, W, C, ld_in, cv.BW, cv.BH);
  if (rc) return rc;
  return dispatch_gemm(tmA, Wt, ldw, B * H * W, N, 9 * C, ep, cv, 0, reinterpret_cast<cudaStream_t>(stream));
}
