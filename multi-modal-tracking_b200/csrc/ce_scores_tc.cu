// Candidate-elimination scores on the tensor cores (bf16 mode of mmt_ce_scores).
//
// Reference: asymmetric_shared_ce.py:202-205 and :91-92
//   attn_t2s = softmax_over_all_2Ls_keys( [q_mt_V; q_mt_I] . [k_s_V; k_s_I]^T * scale )   [B, H, 2*Lt, 2*Ls]
//   score    = attn_t2s.mean(dim=2).mean(dim=1)                                           [B, 2*Ls]
// i.e. per (sequence, head): row softmax of a [2Lt x 2Ls] score matrix, then COLUMN sums.  This kernel produces the
// per-(sequence, head, 128-row query tile) column sums partial[b][h][qt][2Ls]; ce_score_reduce_kernel (attention.cu)
// then averages rows first and heads second, in a fixed order, like the reference.
//
// One CTA = one query tile (<= 128 template rows of one modality) of one head, 2 CTAs per SM.  Same machinery as
// attention_tc.cu: TMA boxes of Q / K straight from the packed qkv buffer, S = Q K^T (M128 N128 K16 x4) in TMEM,
// two softmax threads per query row.  Two passes over the keys:
//   pass A  online row statistics (max m, sum l) - each thread of a row pair keeps its own, merged once at the end;
//   pass B  S recomputed, p = exp2((s - m) * scale * log2 e) / l rounded to bf16 and written TRANSPOSED into a swizzled
//           smem tile PT[key][row]; the column sums over the 128 rows are one more MMA: D = ONES[128 x rows] . PT^T
//           (every row of D then holds the 128 column sums), read back by 16 lanes per warp and stored.
// The products and row statistics are fp32; only P is bf16 (2^-9 relative per term, averaged over 3072 rows x heads).
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "../../include/mmt_b200.h"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>

namespace mmt {
using namespace ptx;

constexpr int CET_HD = 64;
constexpr int CET_KB = 64;
constexpr int CET_STAGES = 4;
constexpr int CET_BLK_BYTES = CET_KB * CET_HD * 2;   // 8 KB
constexpr int CET_Q_BYTES = 128 * CET_HD * 2;        // 16 KB
constexpr int CET_ONES_BYTES = 128 * 128;            // 16 KB: 128 rows x 128 B of bf16 1.0
constexpr int CET_PT_BYTES = 2 * 128 * 128;          // 32 KB: two K-atoms (rows 0-63 / 64-127) of [128 keys x 64]
constexpr int CET_THREADS = 320;
constexpr int CET_SMEM = CET_Q_BYTES + CET_STAGES * CET_BLK_BYTES + CET_ONES_BYTES + CET_PT_BYTES + 1024 + 256 + 2048;
constexpr uint32_t CET_TMEM_COLS = 256;              // S: [0,128)  column sums: [128,256)

__device__ __forceinline__ float cet_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(CET_THREADS, 2)
ce_scores_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tm, int C, int B,
                    int q_seq_rows, int n_tok, int k_row_off, int Lt, int Ls, int nqt_per_mod, float scale_log2e,
                    float* __restrict__ partial) {
  // tmq / q_seq_rows: template rows (queries); tm / n_tok / k_row_off: search rows (keys) - see mmt_ce_scores_split
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;
  const uint32_t ring_smem = q_smem + CET_Q_BYTES;
  const uint32_t ones_smem = ring_smem + CET_STAGES * CET_BLK_BYTES;
  const uint32_t pt_smem = ones_smem + CET_ONES_BYTES;
  const uint32_t bar_base = pt_smem + CET_PT_BYTES;
  auto k_full = [&](int s) { return bar_base + 8u * s; };
  auto k_empty = [&](int s) { return bar_base + 8u * (CET_STAGES + s); };
  const uint32_t q_full = bar_base + 8u * (2 * CET_STAGES);
  const uint32_t s_full = bar_base + 8u * (2 * CET_STAGES + 1);
  const uint32_t s_empty = bar_base + 8u * (2 * CET_STAGES + 2);
  const uint32_t pt_full = bar_base + 8u * (2 * CET_STAGES + 3);
  const uint32_t cs_full = bar_base + 8u * (2 * CET_STAGES + 4);
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * CET_STAGES + 5);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_smem - smem_u32(smem_raw)));
  float* stat = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - smem_u32(smem_raw)));   // [2 halves][2][128]

  const int nqt = 2 * nqt_per_mod;
  const int b = blockIdx.x / nqt, qt = blockIdx.x % nqt;
  const int qmod = qt / nqt_per_mod, qchunk = qt % nqt_per_mod;
  const int h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_row0 = (qmod * B + b) * q_seq_rows + qchunk * 128;
  const int q_rows = min(128, Lt - qchunk * 128);
  // keys: two segments (search rows of the RGB and of the TIR stream of sequence b), 64-row boxes
  const int nbs = (Ls + CET_KB - 1) / CET_KB;     // boxes per segment
  const int nb = 2 * nbs;
  const int nsb = (nb + 1) >> 1;
  auto locate = [&](int blk, int& row0, int& len, int& col0) {
    const bool ghost = blk >= nb;
    if (ghost) blk = nb - 1;
    const int s = blk >= nbs ? 1 : 0, k = blk - s * nbs;
    row0 = (s * B + b) * n_tok + k_row_off + k * CET_KB;
    len = ghost ? 0 : min(CET_KB, Ls - k * CET_KB);
    col0 = s * Ls + k * CET_KB;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm);
    prefetch_tmap(&tmq);
    for (int s = 0; s < CET_STAGES; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); }
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(s_empty, 8);
    mbar_init(pt_full, 8);
    mbar_init(cs_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, CET_TMEM_COLS);
    tmem_relinquish();
  }
  if (warp >= 2) {   // tile of bf16 ones: A operand of the column-sum MMA (generic-proxy writes -> async proxy)
    uint4* ones = reinterpret_cast<uint4*>(smem_raw + (ones_smem - smem_u32(smem_raw)));
    for (int i = threadIdx.x - 64; i < CET_ONES_BYTES / 16; i += CET_THREADS - 64)
      ones[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_cs = tmem_base + 128u;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: Q, then every K pair twice
    if (lane == 0) {
      mbar_expect_tx(q_full, CET_Q_BYTES);
      tma_load_2d(q_smem, &tmq, q_full, h * CET_HD, q_row0);
      tma_load_2d(q_smem + CET_Q_BYTES / 2, &tmq, q_full, h * CET_HD, q_row0 + 64);
      int stage = 0;
      uint32_t phase = 0;
      for (int pass = 0; pass < 2; ++pass)
        for (int blk = 0; blk < 2 * nsb; ++blk) {
          int row0, len, col0;
          locate(blk, row0, len, col0);
          mbar_wait(k_empty(stage), phase ^ 1u);
          mbar_expect_tx(k_full(stage), CET_BLK_BYTES);
          tma_load_2d(ring_smem + stage * CET_BLK_BYTES, &tm, k_full(stage), C + h * CET_HD, row0);
          if (++stage == CET_STAGES) { stage = 0; phase ^= 1u; }
        }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(128, 2 * CET_KB);
      const uint64_t qdesc = make_kmajor_sw128_desc(q_smem);
      const uint64_t odesc = make_kmajor_sw128_desc(ones_smem);
      int stage = 0;
      uint32_t phase = 0;
      mbar_wait(q_full, 0);
      tc_fence_after();
      auto issue_s = [&](int g) {
        mbar_wait(s_empty, (g & 1u) ^ 1u);
        mbar_wait(k_full(stage), phase);
        mbar_wait(k_full(stage + 1), phase);
        tc_fence_after();
        const uint64_t kdesc = make_kmajor_sw128_desc(ring_smem + stage * CET_BLK_BYTES);
#pragma unroll
        for (int k = 0; k < CET_HD / 16; ++k) mma_bf16_ss(tmem_s, qdesc + 2u * k, kdesc + 2u * k, idesc, k ? 1u : 0u);
        mma_commit(s_full);
        mma_commit(k_empty(stage));
        mma_commit(k_empty(stage + 1));
        stage += 2;
        if (stage == CET_STAGES) { stage = 0; phase ^= 1u; }
      };
      for (int j = 0; j < nsb; ++j) issue_s(j);                      // pass A
      for (int j = 0; j <= nsb; ++j) {                               // pass B, column sums one block behind
        if (j < nsb) issue_s(nsb + j);
        if (j >= 1) {
          mbar_wait(pt_full, (j - 1) & 1u);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 8; ++k) {   // K = 128 query rows: PT atom k/4, 32 B per step; the ones operand is reused
            const uint64_t pd = make_kmajor_sw128_desc(pt_smem + (k >> 2) * (CET_PT_BYTES / 2)) + 2u * (k & 3);
            mma_bf16_ss(tmem_cs, odesc, pd, idesc, k ? 1u : 0u);
          }
          mma_commit(cs_full);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps (two threads per query row)
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_addr = tmem_s + lane_off + 64u * half;
    float m_row = -INFINITY, l_row = 0.f;
    // pass A: online (max, sum) of this thread's half of every super-block
    for (int g = 0; g < nsb; ++g) {
      int row0, len, col0;
      locate(2 * g + half, row0, len, col0);
      mbar_wait(s_full, g & 1u);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(s_addr, v0);
      tmem_ld_32x32(s_addr + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);
      float bm = m_row;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < len) bm = fmaxf(bm, __uint_as_float(v0[j]));
        if (j + 32 < len) bm = fmaxf(bm, __uint_as_float(v1[j]));
      }
      if (bm > m_row) {
        l_row *= cet_ex2((m_row - bm) * scale_log2e);     // 0 when m_row was -inf
        m_row = bm;
      }
      const float mc = m_row * scale_log2e;
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < len) acc += cet_ex2(fmaf(__uint_as_float(v0[j]), scale_log2e, -mc));
        if (j + 32 < len) acc += cet_ex2(fmaf(__uint_as_float(v1[j]), scale_log2e, -mc));
      }
      l_row += acc;
    }
    // merge the two halves of the row: m = max, l = sum of the rescaled parts
    stat[(half * 2 + 0) * 128 + r] = m_row;
    stat[(half * 2 + 1) * 128 + r] = l_row;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    {
      const float mo = stat[((half ^ 1) * 2 + 0) * 128 + r], lo = stat[((half ^ 1) * 2 + 1) * 128 + r];
      const float m = fmaxf(m_row, mo);
      float l = 0.f;
      if (m_row > -INFINITY) l += l_row * cet_ex2((m_row - m) * scale_log2e);
      if (mo > -INFINITY) l += lo * cet_ex2((mo - m) * scale_log2e);
      m_row = m;
      l_row = l;
    }
    const float mc = m_row * scale_log2e;
    const bool row_ok = r < q_rows;                              // rows beyond the tile contribute nothing
    const float inv_l = row_ok ? 1.f / l_row : 0.f;
    uint8_t* pt_gen = smem_raw + (pt_smem - smem_u32(smem_raw));
    // this thread's row inside the PT tile: K-atom r/64, 16-byte chunk (r%64)/8 (swizzled by the key row), element r%8
    const uint32_t pt_row_const = static_cast<uint32_t>(r >> 6) * (CET_PT_BYTES / 2) + static_cast<uint32_t>(r & 7) * 2u;
    const uint32_t kchunk = static_cast<uint32_t>((r & 63) >> 3);
    const int cs_w = warp - 2;                                   // this warp reads column sums [16 cs_w, +16)
    auto flush_colsums = [&](int j) {      // D rows are identical: lane i < 16 stores column 16 cs_w + i of block j
      uint32_t cs[32];
      tmem_ld_32x16(tmem_cs + lane_off + 16u * cs_w, cs);
      tmem_ld_wait();
      const int kk = 16 * cs_w + (lane & 15);                    // key inside the super-block
      int row0, len, col0;
      locate(2 * j + (kk >> 6), row0, len, col0);
      float val = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) val = ((lane & 15) == i) ? __uint_as_float(cs[i]) : val;
      if (lane < 16 && (kk & 63) < len)
        partial[((static_cast<size_t>(b) * gridDim.y + h) * nqt + qt) * (2 * Ls) + col0 + (kk & 63)] = val;
    };
    // pass B: normalised probabilities, transposed into PT; column sums by the tensor core
    for (int j = 0; j < nsb; ++j) {
      const int g = nsb + j;
      int row0, len, col0;
      locate(2 * j + half, row0, len, col0);
      mbar_wait(s_full, g & 1u);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(s_addr, v0);
      tmem_ld_32x32(s_addr + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);
      if (j > 0) {                       // column sums of the previous block are complete: PT and D are free after this
        mbar_wait(cs_full, (j - 1) & 1u);
        tc_fence_after();
        flush_colsums(j - 1);
        tc_fence_before();
      }
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float s = __uint_as_float(i < 32 ? v0[i] : v1[i - 32]);
        const float p = (i < len && row_ok) ? cet_ex2(fmaf(s, scale_log2e, -mc)) * inv_l : 0.f;
        const uint32_t key = static_cast<uint32_t>(64 * half + i);           // row of the PT tile
        const uint32_t off = pt_row_const + key * 128u + ((kchunk ^ (key & 7u)) << 4);
        *reinterpret_cast<bf16*>(pt_gen + off) = __float2bfloat16_rn(p);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(pt_full);
    }
    mbar_wait(cs_full, (nsb - 1) & 1u);
    tc_fence_after();
    flush_colsums(nsb - 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, CET_TMEM_COLS);
}

static PFN_cuTensorMapEncodeTiled_v12000 cet_encode_fn() {
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

// partial: fp32 [B, heads, nqt, 2*Ls] with nqt = 2 * ceil(Lt / 128); returns nqt through *nqt_out
int launch_ce_scores_tc(const void* qbuf, int q_seq_rows, const void* qkv, int n_tok, int k_row_off, int ld, int C, int heads,
                        int B, int Lt, int Ls, float scale, float* partial, int* nqt_out, cudaStream_t stream) {
  auto fn = cet_encode_fn();
  if (!fn) return MMT_ERR_UNSUPPORTED;
  CUtensorMap tm, tmq;
  auto encode = [&](CUtensorMap* m, const void* ptr, int rows) {
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(3 * C), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  if (!encode(&tm, qkv, 2 * B * n_tok) || !encode(&tmq, qbuf, 2 * B * q_seq_rows)) return MMT_ERR_BAD_ARG;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(ce_scores_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CET_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int nqt_per_mod = (Lt + 127) / 128;
  *nqt_out = 2 * nqt_per_mod;
  dim3 grid(B * 2 * nqt_per_mod, heads);
  ce_scores_tc_kernel<<<grid, CET_THREADS, CET_SMEM, stream>>>(tmq, tm, C, B, q_seq_rows, n_tok, k_row_off, Lt, Ls, nqt_per_mod,
                                                              scale * 1.4426950408889634f, partial);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? MMT_OK : (int)e;
}

}  // namespace mmt
