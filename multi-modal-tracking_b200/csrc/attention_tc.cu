// Asymmetric mixed attention on the 5th-gen tensor cores (bf16 mode of mmt_mixattn_fwd).
//
// Reference arithmetic: Attention.forward lib/models/mixformer_vit/mixformer.py:51-77 and the cross-modal
// variants lib/models/mixformer_vit_rgbt/asymmetric_shared.py:55-104 / asymmetric_shared_ce.py:146-200:
//   out[q, h*64:(h+1)*64] = softmax_k( q . k * scale ) v     over the key rows listed for the query tile
// (the template/search asymmetry is WHICH key rows a query tile lists - there is no mask tensor).
//
// A work item = one query tile (<= 128 rows) of one head; at most 2 persistent CTAs are resident per SM (70 KB smem, 256
// TMEM columns each) so that the softmax of one overlaps the loads / MMAs / epilogue of the other.  Keys are walked in super-blocks
// of 128 (two 64-row TMA boxes); per-block fixed costs (barrier round trips, TMEM load latency, proxy fence)
// dominate this small problem, so blocks are as large as TMEM allows with two CTAs per SM.
//   warp 0     TMA producer : Q tile (2 x [64 x 64] boxes) and a ring of 2 x 2 K / V boxes, straight out of the
//              packed qkv buffer (cp.async.bulk.tensor, 128B swizzle).
//   warp 1     MMA issuer   : S = Q K^T  (tcgen05.mma M128 N128 K16 x4, both operands K-major smem) into TMEM;
//              O += P V with A = P read from TENSOR MEMORY and B = the V boxes in smem, MN-major (the probabilities
//              never touch shared memory: operand reads from smem, not the MMA rate, bound an SS-mode PV here).
//   warps 2-9  softmax      : two threads per query row (one 64-key box of the super-block each).  Online softmax
//              in ONE pass over the keys with a LAZY running maximum: p = exp2((s - m) * scale * log2 e) is formed
//              against the current m; m (and with it O and the row sum, by exp2((m_old - m_new) ...)) is only moved
//              when a block's maximum exceeds it by more than 2^8 - p stays <= 256, exactly representable headroom
//              for bf16 P and the fp32 accumulators, and the O / row-sum ratio is unchanged.  P is written as packed
//              bf16 into TMEM with tcgen05.st (the A operand of the PV MMA); the rare O rescale is a TMEM
//              load-scale-store by the same threads between two PV MMAs.  Epilogue: O * (1 / row sum) -> bf16 -> out.
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "../../include/mmt_b200.h"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>

namespace mmt {
using namespace ptx;

struct AttnTileTC {      // same 16-int record as AttnTile (attention.cu); q_rows <= 128 here
  int q_row0, q_rows;
  int out_row0;
  int nseg;
  int k_row0[3], k_len[3], k_buf[3];
  int pad[3];
};

// Developer builds only (tools/build_variant.py -> csrc/libmmt_b200_<name>.so, loaded by MMT_B200_DEV_LIB=<name>; the shipped
// library is compiled with 0): timing experiments of profiles/r2_attention.md.  Bits 1-8 make the RESULTS WRONG by construction.
//   1  no row-maximum exchange between the two threads of a row     2  no exponentials (a multiply instead)
//   4  no global stores in the epilogue                            8  one CTA per SM instead of two
//  16  phase cycle stamps of softmax warp 4 lane 0 (results stay correct; read back with mmt_dev_attn_stamps, tools/attn_stamps.py)
#ifndef MMT_ATTN_EXP
#define MMT_ATTN_EXP 0
#endif
#if MMT_ATTN_EXP & 16
__device__ unsigned long long g_attn_stamp[512][12];     // per-CTA phase cycle sums of softmax warp 4, lane 0
#define ATTN_STAMP(i, a, b) do { if (stamping) acc[i] += (unsigned long long)((b) - (a)); } while (0)
#define ATTN_CLK() clock64()
#else
#define ATTN_STAMP(i, a, b) do { (void)(a); (void)(b); } while (0)
#define ATTN_CLK() 0ll
#endif
constexpr int ATC_HD = 64;
constexpr int ATC_KB = 64;                          // keys per block
#ifndef MMT_ATTN_STAGES
#define MMT_ATTN_STAGES 4      // K/V ring depth in 64-key blocks (even); 6: 84.8 vs 82.8 us (profiles/r2_attention.md)
#endif
constexpr int ATC_STAGES = MMT_ATTN_STAGES;
constexpr int ATC_BLK_BYTES = ATC_KB * ATC_HD * 2;  // 8 KB
constexpr int ATC_Q_BYTES = 128 * ATC_HD * 2;       // 16 KB
constexpr int ATC_THREADS = 320;
constexpr uint32_t ATC_TMEM_COLS = 256;             // O: [0,64)  S: [64,192)  P: [192,256) (128 keys, bf16x2 per column)

// kind::f16 instruction descriptor with B taken MN-major (bit 16): V blocks are [key][d] with d contiguous.
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32_bmn(int m, int n) {
  return make_idesc_bf16_f32(m, n) | (1u << 16);
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for x <= 0 on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f, |f| <= 0.5, cubic for 2^f
// (relative error 6e-4, far below the bf16 rounding of P), exponent patched in with an integer add.  Every fourth
// probability takes this path so that the MUFU unit - the bound of the softmax - gets 25% less work.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;              // 1.5 * 2^23: integer part of x lands in the low mantissa bits
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.0555041f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// Two 2^x at once with packed fp32 instructions (fma.rn.f32x2 / add.f32x2: one issue slot for two lanes of work).
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);
  const float2 t = __fadd2_rn(x, magic);
  const float2 tm = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(tm, make_float2(-1.f, -1.f), x);
  float2 p = __ffma2_rn(f, make_float2(0.0555041f, 0.0555041f), make_float2(0.2402265f, 0.2402265f));
  p = __ffma2_rn(p, f, make_float2(0.6931472f, 0.6931472f));
  p = __ffma2_rn(p, f, make_float2(1.0f, 1.0f));
  p.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  p.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return p;
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent schedule: at most two CTAs per SM walk the work items (query tile, head) with a fixed stride, so the per-CTA
// fixed costs - TMEM allocation, barrier initialisation, the first TMA round trip, the epilogue / teardown tail: ~8 k of
// the ~22 k cycles a search tile cost in the one-CTA-per-item form this kernel replaced (profiles/r1_attention_phases.md;
// that form was deleted in round 2) - are paid once per CTA or hidden: the producer warp runs ahead into the next item (Q is
// double-buffered, the K/V ring simply continues), the MMA warp starts the next item's S = Q K^T while the softmax
// warps are still in the previous item's epilogue.  Barrier phases run on counters that continue across items:
//   q_full/q_empty[2]  Q buffer n & 1 of the CTA's n-th item (q_empty: committed after the item's last S MMA)
//   s_full/s_empty, p_full/p_empty   one phase per 128-key super-block, counted over all items
//   o_full/o_empty     one phase per item (o_empty: the eight softmax warps have read O in the epilogue - the next
//                      item's first P V MMA overwrites it)
// Items are ordered tile-major with the head fastest; the host sorts the tile table by key count (heavy first), so a
// stride walk gives every CTA the same mix of 4-block search tiles and 1-block template tiles.
constexpr int ATP_STG_BYTES = 8 * 2048;   // one 32 x 32 bf16 staging tile per softmax warp (epilogue TMA stores)
constexpr int ATP_SMEM = 2 * ATC_Q_BYTES + ATC_STAGES * ATC_BLK_BYTES + ATP_STG_BYTES + 1024 + 256 + 3072;

__global__ void __launch_bounds__(ATC_THREADS, 2)
attn_tc_persist_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                       const __grid_constant__ CUtensorMap tmo, int C,
                       const AttnTileTC* __restrict__ tiles, int n_items, int heads, bf16* __restrict__ out, int ldo,
                       float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;                                  // two Q buffers
  const uint32_t ring_smem = q_smem + 2 * ATC_Q_BYTES;
  const uint32_t stg_smem = ring_smem + ATC_STAGES * ATC_BLK_BYTES;  // 64 KB past the 1024-aligned base
  const uint32_t bar_base = stg_smem + ATP_STG_BYTES;
  auto kv_full = [&](int s) { return bar_base + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 8u * (ATC_STAGES + s); };
  auto q_full = [&](int b) { return bar_base + 8u * (2 * ATC_STAGES + b); };
  auto q_empty = [&](int b) { return bar_base + 8u * (2 * ATC_STAGES + 2 + b); };
  const uint32_t s_full = bar_base + 8u * (2 * ATC_STAGES + 4);
  const uint32_t s_empty = bar_base + 8u * (2 * ATC_STAGES + 5);
  const uint32_t p_full = bar_base + 8u * (2 * ATC_STAGES + 6);
  const uint32_t p_empty = bar_base + 8u * (2 * ATC_STAGES + 7);
  const uint32_t o_full = bar_base + 8u * (2 * ATC_STAGES + 8);
  const uint32_t o_empty = bar_base + 8u * (2 * ATC_STAGES + 9);
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * ATC_STAGES + 10);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_smem - smem_u32(smem_raw)));
  // [2 block parities][2 halves][128] row maxima + [2 halves][128] row sums of the epilogue
  float* smax = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm0);
    prefetch_tmap(&tm1);
    prefetch_tmap(&tmo);
    for (int s = 0; s < ATC_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(q_full(b), 1); mbar_init(q_empty(b), 1); }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 8);   // one arrive per softmax warp
    mbar_init(p_full, 8);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, ATC_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const uint32_t tmem_o = tmem_base;
  const uint32_t tmem_p = tmem_base + 192u;
  const uint32_t tmem_s = tmem_base + 64u;
  // the prologue above touched only this CTA's shared / tensor memory: it may overlap the previous kernel (the qkv GEMM)
  pdl_wait();
  pdl_launch_dependents();

  // key-box bookkeeping of one tile
  struct Plan {
    int nblk_seg[3];
    int nb, nsb;
  };
  auto make_plan = [](const AttnTileTC& t) {
    Plan p;
    p.nb = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      p.nblk_seg[s] = s < t.nseg ? (t.k_len[s] + ATC_KB - 1) / ATC_KB : 0;
      p.nb += p.nblk_seg[s];
    }
    p.nsb = (p.nb + 1) >> 1;
    return p;
  };
  auto locate = [](const AttnTileTC& t, const Plan& p, int blk, int& row0, int& len, int& buf) {
    const bool ghost = blk >= p.nb;
    if (ghost) blk = p.nb - 1;
    int s = 0;
    if (blk >= p.nblk_seg[0]) { blk -= p.nblk_seg[0]; s = 1; if (blk >= p.nblk_seg[1]) { blk -= p.nblk_seg[1]; s = 2; } }
    row0 = t.k_row0[s] + blk * ATC_KB;
    len = ghost ? 0 : min(ATC_KB, t.k_len[s] - blk * ATC_KB);
    buf = t.k_buf[s];
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        const AttnTileTC t = tiles[item / heads];
        const int h = item % heads;
        const Plan pl = make_plan(t);
        const int qb = n & 1;
        mbar_wait(q_empty(qb), ((n >> 1) & 1u) ^ 1u);          // the S MMAs of item n-2 have read this buffer
        mbar_expect_tx(q_full(qb), ATC_Q_BYTES);
        tma_load_2d(q_smem + qb * ATC_Q_BYTES, &tm0, q_full(qb), h * ATC_HD, t.q_row0);
        tma_load_2d(q_smem + qb * ATC_Q_BYTES + ATC_Q_BYTES / 2, &tm0, q_full(qb), h * ATC_HD, t.q_row0 + 64);
        auto load_blk = [&](int blk, int col) {
          int row0, len, buf;
          locate(t, pl, blk, row0, len, buf);
          mbar_wait(kv_empty(stage), phase ^ 1u);
          mbar_expect_tx(kv_full(stage), ATC_BLK_BYTES);
          tma_load_2d(ring_smem + stage * ATC_BLK_BYTES, buf ? &tm1 : &tm0, kv_full(stage), col, row0);
          if (++stage == ATC_STAGES) { stage = 0; phase ^= 1u; }
        };
        for (int sb = 0; sb <= pl.nsb; ++sb) {                                  // K_sb, then V_(sb-1)
          if (sb < pl.nsb) { load_blk(2 * sb, C + h * ATC_HD); load_blk(2 * sb + 1, C + h * ATC_HD); }
          if (sb >= 1) { load_blk(2 * sb - 2, 2 * C + h * ATC_HD); load_blk(2 * sb - 1, 2 * C + h * ATC_HD); }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(128, 2 * ATC_KB);
      constexpr uint32_t idesc_o = make_idesc_bf16_f32_bmn(128, ATC_HD);
      int stage = 0;           // always even here: a super-block is the stage pair (stage, stage + 1)
      uint32_t phase = 0;
      uint32_t g = 0;          // S super-blocks issued so far (all items)
      uint32_t jj = 0;         // P V super-blocks issued so far
      auto wait_pair = [&]() {
        mbar_wait(kv_full(stage), phase);
        mbar_wait(kv_full(stage + 1), phase);
        tc_fence_after();
      };
      auto release_pair = [&]() {
        mma_commit(kv_empty(stage));
        mma_commit(kv_empty(stage + 1));
        stage += 2;
        if (stage == ATC_STAGES) { stage = 0; phase ^= 1u; }
      };
      int n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        const AttnTileTC t = tiles[item / heads];
        const Plan pl = make_plan(t);
        const int qb = n & 1;
        mbar_wait(q_full(qb), (n >> 1) & 1u);
        tc_fence_after();
        const uint64_t qdesc = make_kmajor_sw128_desc(q_smem + qb * ATC_Q_BYTES);
        for (int sb = 0; sb <= pl.nsb; ++sb) {
          if (sb < pl.nsb) {
            mbar_wait(s_empty, (g & 1u) ^ 1u);
            wait_pair();
            const uint64_t kdesc = make_kmajor_sw128_desc(ring_smem + stage * ATC_BLK_BYTES);   // 128 key rows
#pragma unroll
            for (int k = 0; k < ATC_HD / 16; ++k)
              mma_bf16_ss(tmem_s, qdesc + 2u * k, kdesc + 2u * k, idesc_s, k ? 1u : 0u);
            mma_commit(s_full);
            if (sb == pl.nsb - 1) mma_commit(q_empty(qb));       // last read of this item's Q
            release_pair();
            ++g;
          }
          if (sb >= 1) {
            const int j = sb - 1;
            mbar_wait(p_full, jj & 1u);
            if (j == 0) mbar_wait(o_empty, (n & 1u) ^ 1u);       // the previous item's epilogue has read O
            wait_pair();
            tc_fence_after();
            const uint64_t vdesc = make_kmajor_sw128_desc(ring_smem + stage * ATC_BLK_BYTES);
#pragma unroll
            for (int k = 0; k < 2 * ATC_KB / 16; ++k)   // 16 keys per MMA: P advances 8 TMEM columns, V 16 rows = 2 KB
              mma_bf16_ts(tmem_o, tmem_p + 8u * k, vdesc + 128u * k, idesc_o, (j | k) ? 1u : 0u);
            mma_commit(p_empty);
            release_pair();
            ++jj;
          }
        }
        mma_commit(o_full);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps (two threads per query row)
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;               // which 32 of a block's 64 keys this thread handles
    const int r = quad * 32 + lane;                 // row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_addr = tmem_s + lane_off + 64u * half;
    const uint32_t o_addr = tmem_o + lane_off + 32u * half;
    uint32_t g = 0;               // super-blocks processed so far (all items)
    int n = 0;
#if MMT_ATTN_EXP & 16
    const bool stamping = warp == 4 && lane == 0;
    unsigned long long acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long tk0 = clock64();
#endif
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      const AttnTileTC t = tiles[item / heads];
      const int h = item % heads;
      const Plan pl = make_plan(t);
      float m_row = -INFINITY;      // lazy running maximum (raw score units)
      float mc = 0.f;               // m_row * scale_log2e
      float l_row = 0.f;            // this thread's part of the row sum (partially valid column groups)
      float2 l2 = make_float2(0.f, 0.f);   // ... and of the fully valid ones, as two packed partial sums
      auto row_max = [&](const uint32_t (&v)[32], int lim0, float m) {
        if (lim0 >= 32) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) m = fmaxf(m, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < lim0) m = fmaxf(m, __uint_as_float(v[j]));
        }
        return m;
      };
      auto probs = [&](uint32_t (&v)[32], int lim0) {
        uint32_t* pk = v;
        if (lim0 >= 32) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            // packed fp32 arithmetic (the softmax warps are issue-bound: profiles/r2_attention.md); every fourth PAIR of
            // exponentials on the FMA pipe instead of the MUFU unit
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])),
                                        make_float2(scale_log2e, scale_log2e), make_float2(-mc, -mc));
#if MMT_ATTN_EXP & 2
            const float2 pp = __fmul2_rn(x, make_float2(0.001f, 0.001f));
#else
#ifndef MMT_ATTN_POLY_SET
#define MMT_ATTN_POLY_SET 0x8888      // bit i set: pair i of the 16 goes to the FMA-pipe cubic instead of the MUFU unit
#endif
            const float2 pp = ((MMT_ATTN_POLY_SET >> i) & 1) ? ex2_poly2(x) : make_float2(ex2_approx(x.x), ex2_approx(x.y));
#endif
            l2 = __fadd2_rn(l2, pp);
            pk[i] = pack_bf16x2(pp.x, pp.y);
          }
        } else if (lim0 <= 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), scale_log2e, -mc));
            float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), scale_log2e, -mc));
            if (2 * i >= lim0) p0 = 0.f;
            if (2 * i + 1 >= lim0) p1 = 0.f;
            l_row += p0 + p1;
            pk[i] = pack_bf16x2(p0, p1);
          }
        }
      };
      const bool warp_live = quad * 32 < t.q_rows;
      for (int j = 0; j < pl.nsb; ++j, ++g) {
        int row0, len, buf;
        locate(t, pl, 2 * j + half, row0, len, buf);
        const long long t0 = ATTN_CLK();
        mbar_wait(s_full, g & 1u);
        tc_fence_after();
        const long long t1 = ATTN_CLK();
        // tensor-memory reads are the scarcest resource of this kernel at head dimension 64 (64 KB of S per 128 x 128 block
        // against 512 MMA cycles): warps without a valid row and column groups without a valid key are not read at all
        if (!warp_live) {          // no valid row in this TMEM lane quadrant (both warps of the pair): keep the barrier protocol only;
          __syncwarp();            // P / O rows of dead lanes hold stale finite-or-not values that no valid row ever sees
          if (lane == 0) mbar_arrive(s_empty);
          mbar_wait(p_empty, (g & 1u) ^ 1u);
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full);
          continue;
        }
        uint32_t v0[32], v1[32];
        const bool two = len > 32;                 // warp-uniform: the second column group holds at least one valid key
        tmem_ld_32x32(s_addr, v0);
        if (two) tmem_ld_32x32(s_addr + 32, v1);
        tmem_ld_wait();
        const long long t2 = ATTN_CLK();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty);
        float bm = row_max(v0, len, -INFINITY);
        if (two) bm = row_max(v1, len - 32, bm);
        float* ex = smax + (g & 1u) * 256;
        ex[half * 128 + r] = bm;
        // only the two warps that share this TMEM lane quadrant (the two threads of a row live in warps q and q + 4) meet:
        // four independent 64-thread barriers instead of one over all eight softmax warps
#if !(MMT_ATTN_EXP & 1)
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        bm = fmaxf(bm, ex[(half ^ 1) * 128 + r]);
#endif
        const long long t3 = ATTN_CLK();
        float factor = 1.f;
        bool moved = false;
        if ((bm - m_row) * scale_log2e > 8.f) {
          factor = ex2_approx((m_row - bm) * scale_log2e);   // 0 on the first block
          m_row = bm;
          mc = m_row * scale_log2e;
          l_row *= factor;
          l2 = __fmul2_rn(l2, make_float2(factor, factor));
          moved = warp_live;
        }
        probs(v0, warp_live ? len : 0);
        const long long t4 = ATTN_CLK();
        mbar_wait(p_empty, (g & 1u) ^ 1u);          // the PV MMAs of the previous super-block (any item) have completed
        tc_fence_after();
        const long long t5 = ATTN_CLK();
        if (j > 0 && __any_sync(0xffffffffu, moved)) {
          uint32_t o[32];
          tmem_ld_32x32(o_addr, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
          tmem_st_32x32(o_addr, o);
        }
        tmem_st_32x16(tmem_p + lane_off + 32u * half, v0);
        probs(v1, warp_live ? len - 32 : 0);
        tmem_st_32x16(tmem_p + lane_off + 32u * half + 16u, v1);
        tmem_st_wait();
        const long long t6 = ATTN_CLK();
        ATTN_STAMP(0, t0, t1); ATTN_STAMP(1, t1, t2); ATTN_STAMP(2, t2, t3); ATTN_STAMP(3, t3, t4);
        ATTN_STAMP(4, t4, t5); ATTN_STAMP(5, t5, t6); ATTN_STAMP(8, 0, 1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      // epilogue: O / L -> bf16 -> out (this thread: 32 of the 64 head channels)
      const long long t7 = ATTN_CLK();
      mbar_wait(o_full, n & 1u);
      tc_fence_after();
      const long long t8 = ATTN_CLK();
      uint32_t o[32];
      tmem_ld_32x32(o_addr, o);
      float* exl = smax + 512;                      // dedicated row-sum exchange buffer
      l_row += l2.x + l2.y;
      exl[half * 128 + r] = l_row;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      l_row += exl[(half ^ 1) * 128 + r];
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);          // O may be overwritten by the next item's first P V
      // A warp whose 32 rows are all valid leaves through a 64B-swizzled staging tile and ONE TMA store of a 32 x 32 box
      // (per-thread 16-byte stores of the thread == row layout fill half a 32-byte sector each: measured at 20 % of the
      // launch, profiles/r2_attention.md); the warp that straddles the end of a short tile stores its valid rows directly.
      const bool warp_full = quad * 32 + 32 <= t.q_rows;
      if ((warp_full || r < t.q_rows) && !(MMT_ATTN_EXP & 4)) {
        const float inv = 1.f / l_row;
        uint32_t w[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) w[c] = pack_bf16x2(__uint_as_float(o[2 * c]) * inv, __uint_as_float(o[2 * c + 1]) * inv);
        if (warp_full) {
          uint4* stg = reinterpret_cast<uint4*>(smem_raw + (stg_smem - smem_u32(smem_raw))) + (warp - 2) * 128;
          if (lane == 0) bulk_wait_read_all();       // the previous item's store has read this warp's tile
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j)     // 64-byte rows, 16-byte chunk j of row `lane` at j ^ ((lane >> 1) & 3): SWIZZLE_64B
            stg[lane * 4 + (j ^ ((lane >> 1) & 3))] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmo, smem_u32(stg), h * ATC_HD + 32 * half, t.out_row0 + quad * 32);
            bulk_commit_group();
          }
        } else {
          uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(t.out_row0 + r) * ldo + h * ATC_HD + 32 * half);
#pragma unroll
          for (int c = 0; c < 4; ++c) dst[c] = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        }
      }
      const long long t9 = ATTN_CLK();
      ATTN_STAMP(6, t7, t8); ATTN_STAMP(7, t8, t9); ATTN_STAMP(9, 0, 1);
    }
    if (lane == 0) bulk_wait_all();
#if MMT_ATTN_EXP & 16
    if (stamping && blockIdx.x < 512) {
      acc[10] = (unsigned long long)(clock64() - tk0);
      for (int i = 0; i < 12; ++i) g_attn_stamp[blockIdx.x][i] = acc[i];
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ATC_TMEM_COLS);
}


static PFN_cuTensorMapEncodeTiled_v12000 attn_encode_fn() {
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

// qkv [rows, ld] bf16 viewed as a 2-D tensor; box = 64 columns (one head slice) x 64 rows, 128B swizzle.
static int make_qkv_tmap(CUtensorMap* tm, const void* ptr, int rows, int cols, int ld) {
  auto fn = attn_encode_fn();
  if (!fn) return MMT_ERR_UNSUPPORTED;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, 64};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MMT_OK : MMT_ERR_BAD_ARG;
}

// bf16 output [rows, cols]: 32-row boxes of the epilogue's per-warp staging tile (32 columns / 64-byte swizzle in the
// half-row form, 64 columns / 128-byte swizzle in the row form)
static int make_out_tmap(CUtensorMap* tm, const void* ptr, int rows, int cols, int ld, int box_cols) {
  auto fn = attn_encode_fn();
  if (!fn) return MMT_ERR_UNSUPPORTED;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MMT_OK : MMT_ERR_BAD_ARG;
}

int launch_attn_tc(const void* qkv0, int rows0, const void* qkv1, int rows1, int ld, int C, int heads,
                   const int* tiles_dev, int n_tiles, void* out, int ldo, float scale, cudaStream_t stream) {
  CUtensorMap tm0, tm1, tmo;
  int rc = make_qkv_tmap(&tm0, qkv0, rows0, 3 * C, ld);
  if (rc) return rc;
  rc = make_qkv_tmap(&tm1, qkv1, rows1, 3 * C, ld);
  if (rc) return rc;
  // output rows are addressed through the tile table: the map's row bound is nominal (only groups of 32 valid rows are
  // stored through it, so its clipping is never relied upon)
  rc = make_out_tmap(&tmo, out, 1 << 30, C, ldo, 32);
  if (rc) return rc;
  static bool attr_p = false;
  static int n_sm = 0;
  if (!attr_p) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATP_SMEM);
    if (e != cudaSuccess) return (int)e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    attr_p = true;
  }
  const int n_items = n_tiles * heads;
#if MMT_ATTN_EXP & 8
  const int grid = n_items < n_sm ? n_items : n_sm;              // one CTA per SM
#else
  const int grid = n_items < 2 * n_sm ? n_items : 2 * n_sm;
#endif
  cudaError_t e = launch_pdl(attn_tc_persist_kernel, dim3(grid), dim3(ATC_THREADS), ATP_SMEM, stream, tm0, tm1, tmo, C,
                             reinterpret_cast<const AttnTileTC*>(tiles_dev), n_items, heads, reinterpret_cast<bf16*>(out), ldo,
                             scale * 1.4426950408889634f);
  if (e != cudaSuccess) return (int)e;
  e = cudaGetLastError();
  return e == cudaSuccess ? MMT_OK : (int)e;
}

}  // namespace mmt

#if MMT_ATTN_EXP & 16
extern "C" int mmt_dev_attn_stamps(unsigned long long* host_out) {     // [512][12], developer builds only
  return (int)cudaMemcpyFromSymbol(host_out, mmt::g_attn_stamp, sizeof(mmt::g_attn_stamp));
}
#endif
