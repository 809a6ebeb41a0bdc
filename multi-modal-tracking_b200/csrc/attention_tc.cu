// Asymmetric mixed attention on the 5th-gen tensor cores (bf16 mode of mmt_mixattn_fwd).
//
// Reference arithmetic: Attention.forward lib/models/mixformer_vit/mixformer.py:51-77 and the cross-modal
// variants lib/models/mixformer_vit_rgbt/asymmetric_shared.py:55-104 / asymmetric_shared_ce.py:146-200:
//   out[q, h*64:(h+1)*64] = softmax_k( q . k * scale ) v     over the key rows listed for the query tile
// (the template/search asymmetry is WHICH key rows a query tile lists - there is no mask tensor).
//
// One CTA = one query tile (<= 128 rows) of one head; 2 CTAs are resident per SM (81 KB smem, 256 TMEM columns)
// so that the softmax of one overlaps the loads / MMAs / epilogue of the other.  Keys are walked in 64-row blocks.
//   warp 0     TMA producer : Q tile (2 x [64 x 64] boxes) and a 4-stage ring of K / V blocks, straight out of the
//              packed qkv buffer (cp.async.bulk.tensor, 128B swizzle).
//   warp 1     MMA issuer   : S = Q K^T  (tcgen05.mma M128 N64 K16 x4, both operands K-major) into a
//              double-buffered TMEM tile; O += P V (A = P from smem, K-major; B = V block, MN-major) into TMEM.
//   warps 2-5  softmax      : thread == query row.  EXACT two-pass softmax: pass 1 reads every S block for the row
//              maximum; pass 2 recomputes S (tensor time is cheap here, the exponentials are the bound), forms
//              p = exp2((s - max) * scale * log2 e), accumulates the row sum in fp32, writes P as bf16 into a
//              double-buffered swizzled smem tile (the A operand of the PV MMA) - so O never needs rescaling.
//              Epilogue: O * (1 / row sum) -> bf16 -> out.
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

constexpr int ATC_HD = 64;
constexpr int ATC_KB = 64;                          // keys per block
constexpr int ATC_STAGES = 4;
constexpr int ATC_BLK_BYTES = ATC_KB * ATC_HD * 2;  // 8 KB
constexpr int ATC_Q_BYTES = 128 * ATC_HD * 2;       // 16 KB
constexpr int ATC_P_BYTES = 128 * ATC_KB * 2;       // 16 KB
constexpr int ATC_THREADS = 192;
constexpr int ATC_SMEM = ATC_Q_BYTES + ATC_STAGES * ATC_BLK_BYTES + 2 * ATC_P_BYTES + 1024 + 256;
constexpr uint32_t ATC_TMEM_COLS = 256;             // O: [0,64)  S0: [64,128)  S1: [128,192)

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

__global__ void __launch_bounds__(ATC_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1, int C,
               const AttnTileTC* __restrict__ tiles, bf16* __restrict__ out, int ldo, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;
  const uint32_t ring_smem = q_smem + ATC_Q_BYTES;
  const uint32_t p_smem = ring_smem + ATC_STAGES * ATC_BLK_BYTES;
  const uint32_t bar_base = p_smem + 2 * ATC_P_BYTES;
  auto kv_full = [&](int s) { return bar_base + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 8u * (ATC_STAGES + s); };
  const uint32_t q_full = bar_base + 8u * (2 * ATC_STAGES);
  auto s_full = [&](int b) { return bar_base + 8u * (2 * ATC_STAGES + 1 + b); };
  auto s_empty = [&](int b) { return bar_base + 8u * (2 * ATC_STAGES + 3 + b); };
  auto p_full = [&](int b) { return bar_base + 8u * (2 * ATC_STAGES + 5 + b); };
  auto p_empty = [&](int b) { return bar_base + 8u * (2 * ATC_STAGES + 7 + b); };
  const uint32_t o_full = bar_base + 8u * (2 * ATC_STAGES + 9);
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * ATC_STAGES + 10);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_smem - smem_u32(smem_raw)));

  const AttnTileTC t = tiles[blockIdx.x];
  const int h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // key blocks of this tile: segment s contributes ceil(k_len[s] / 64) blocks
  int nblk_seg[3];
  int nb = 0;
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    nblk_seg[s] = s < t.nseg ? (t.k_len[s] + ATC_KB - 1) / ATC_KB : 0;
    nb += nblk_seg[s];
  }
  auto locate = [&](int blk, int& row0, int& len, int& buf) {
    int s = 0;
    if (blk >= nblk_seg[0]) { blk -= nblk_seg[0]; s = 1; if (blk >= nblk_seg[1]) { blk -= nblk_seg[1]; s = 2; } }
    row0 = t.k_row0[s] + blk * ATC_KB;
    len = min(ATC_KB, t.k_len[s] - blk * ATC_KB);
    buf = t.k_buf[s];
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm0);
    prefetch_tmap(&tm1);
    for (int s = 0; s < ATC_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    mbar_init(q_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(s_full(b), 1);
      mbar_init(s_empty(b), 4);   // one arrive per softmax warp
      mbar_init(p_full(b), 4);
      mbar_init(p_empty(b), 1);
    }
    mbar_init(o_full, 1);
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
  auto tmem_s = [&](int b) { return tmem_base + 64u + 64u * b; };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, ATC_Q_BYTES);
      tma_load_2d(q_smem, &tm0, q_full, h * ATC_HD, t.q_row0);
      tma_load_2d(q_smem + ATC_Q_BYTES / 2, &tm0, q_full, h * ATC_HD, t.q_row0 + 64);
      int stage = 0;
      uint32_t phase = 0;
      auto load_blk = [&](int blk, int col) {
        int row0, len, buf;
        locate(blk, row0, len, buf);
        mbar_wait(kv_empty(stage), phase ^ 1u);
        mbar_expect_tx(kv_full(stage), ATC_BLK_BYTES);
        tma_load_2d(ring_smem + stage * ATC_BLK_BYTES, buf ? &tm1 : &tm0, kv_full(stage), col, row0);
        if (++stage == ATC_STAGES) { stage = 0; phase ^= 1u; }
      };
      for (int kb = 0; kb < nb; ++kb) load_blk(kb, C + h * ATC_HD);          // pass 1: K only
      for (int kb = 0; kb <= nb; ++kb) {                                     // pass 2: K_kb, then V_(kb-1)
        if (kb < nb) load_blk(kb, C + h * ATC_HD);
        if (kb >= 1) load_blk(kb - 1, 2 * C + h * ATC_HD);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(128, ATC_KB);
      constexpr uint32_t idesc_o = make_idesc_bf16_f32_bmn(128, ATC_HD);
      const uint64_t qdesc = make_kmajor_sw128_desc(q_smem);
      int stage = 0;
      uint32_t phase = 0;
      mbar_wait(q_full, 0);
      tc_fence_after();
      auto issue_s = [&](int g) {   // g = running S-block counter over both passes
        const int b = g & 1;
        mbar_wait(s_empty(b), ((g >> 1) & 1u) ^ 1u);
        mbar_wait(kv_full(stage), phase);
        tc_fence_after();
        const uint64_t kdesc = make_kmajor_sw128_desc(ring_smem + stage * ATC_BLK_BYTES);
#pragma unroll
        for (int k = 0; k < ATC_HD / 16; ++k) mma_bf16_ss(tmem_s(b), qdesc + 2u * k, kdesc + 2u * k, idesc_s, k ? 1u : 0u);
        mma_commit(kv_empty(stage));
        mma_commit(s_full(b));
        if (++stage == ATC_STAGES) { stage = 0; phase ^= 1u; }
      };
      for (int kb = 0; kb < nb; ++kb) issue_s(kb);
      for (int kb = 0; kb <= nb; ++kb) {
        if (kb < nb) issue_s(nb + kb);
        if (kb >= 1) {
          const int j = kb - 1, b = j & 1;
          mbar_wait(p_full(b), (j >> 1) & 1u);
          mbar_wait(kv_full(stage), phase);
          tc_fence_after();
          const uint64_t pdesc = make_kmajor_sw128_desc(p_smem + b * ATC_P_BYTES);
          const uint64_t vdesc = make_kmajor_sw128_desc(ring_smem + stage * ATC_BLK_BYTES);
#pragma unroll
          for (int k = 0; k < ATC_KB / 16; ++k)   // 16 keys per MMA: P advances 32 B inside the atom, V 16 rows = 2 KB
            mma_bf16_ss(tmem_o, pdesc + 2u * k, vdesc + 128u * k, idesc_o, (j | k) ? 1u : 0u);
          mma_commit(kv_empty(stage));
          mma_commit(p_empty(b));
          if (++stage == ATC_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      mma_commit(o_full);
    }
  } else {
    // ------------------------------------------------------------------ softmax warps (thread == query row)
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                 // row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    float m_row = -INFINITY;
    // pass 1: row maximum of the raw scores
    for (int g = 0; g < nb; ++g) {
      const int b = g & 1;
      int row0, len, buf;
      locate(g, row0, len, buf);
      mbar_wait(s_full(b), (g >> 1) & 1u);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(tmem_s(b) + lane_off, v0);
      tmem_ld_32x32(tmem_s(b) + lane_off + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty(b));
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < len) m_row = fmaxf(m_row, __uint_as_float(v0[j]));
        if (j + 32 < len) m_row = fmaxf(m_row, __uint_as_float(v1[j]));
      }
    }
    const float mc = m_row * scale_log2e;
    float l_row = 0.f;
    uint8_t* p_gen = smem_raw + (p_smem - smem_u32(smem_raw));
    // pass 2: probabilities -> smem (A operand of the PV MMA), row sums
    for (int j = 0; j < nb; ++j) {
      const int g = nb + j, b = g & 1, pb = j & 1;
      int row0, len, buf;
      locate(j, row0, len, buf);
      mbar_wait(s_full(b), (g >> 1) & 1u);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(tmem_s(b) + lane_off, v0);
      tmem_ld_32x32(tmem_s(b) + lane_off + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty(b));
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float p0 = ex2_approx(fmaf(__uint_as_float(v0[2 * i]), scale_log2e, -mc));
        float p1 = ex2_approx(fmaf(__uint_as_float(v0[2 * i + 1]), scale_log2e, -mc));
        float p2 = ex2_approx(fmaf(__uint_as_float(v1[2 * i]), scale_log2e, -mc));
        float p3 = ex2_approx(fmaf(__uint_as_float(v1[2 * i + 1]), scale_log2e, -mc));
        if (2 * i >= len) p0 = 0.f;
        if (2 * i + 1 >= len) p1 = 0.f;
        if (2 * i + 32 >= len) p2 = 0.f;
        if (2 * i + 33 >= len) p3 = 0.f;
        // the row sum uses the bf16-rounded probabilities, i.e. exactly what the PV MMA multiplies
        const __nv_bfloat162 q01 = __floats2bfloat162_rn(p0, p1), q23 = __floats2bfloat162_rn(p2, p3);
        const float2 f01 = __bfloat1622float2(q01), f23 = __bfloat1622float2(q23);
        l_row += (f01.x + f01.y) + (f23.x + f23.y);
        pk[i] = *reinterpret_cast<const uint32_t*>(&q01);
        pk[16 + i] = *reinterpret_cast<const uint32_t*>(&q23);
      }
      mbar_wait(p_empty(pb), ((j >> 1) & 1u) ^ 1u);
      uint4* prow = reinterpret_cast<uint4*>(p_gen + pb * ATC_P_BYTES + r * 128);
#pragma unroll
      for (int c = 0; c < 8; ++c)   // 16-byte chunk c (keys 8c..8c+7) at slot c ^ (row & 7): the TMA/UMMA 128B swizzle
        prow[c ^ (r & 7)] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(pb));
    }
    // epilogue: O / l -> bf16 -> out
    mbar_wait(o_full, 0);
    tc_fence_after();
    uint32_t o0[32], o1[32];
    tmem_ld_32x32(tmem_o + lane_off, o0);
    tmem_ld_32x32(tmem_o + lane_off + 32, o1);
    tmem_ld_wait();
    if (r < t.q_rows) {
      const float inv = 1.f / l_row;
      uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(t.out_row0 + r) * ldo + h * ATC_HD);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(o0[8 * c]) * inv, __uint_as_float(o0[8 * c + 1]) * inv);
        w.y = pack_bf16x2(__uint_as_float(o0[8 * c + 2]) * inv, __uint_as_float(o0[8 * c + 3]) * inv);
        w.z = pack_bf16x2(__uint_as_float(o0[8 * c + 4]) * inv, __uint_as_float(o0[8 * c + 5]) * inv);
        w.w = pack_bf16x2(__uint_as_float(o0[8 * c + 6]) * inv, __uint_as_float(o0[8 * c + 7]) * inv);
        dst[c] = w;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(o1[8 * c]) * inv, __uint_as_float(o1[8 * c + 1]) * inv);
        w.y = pack_bf16x2(__uint_as_float(o1[8 * c + 2]) * inv, __uint_as_float(o1[8 * c + 3]) * inv);
        w.z = pack_bf16x2(__uint_as_float(o1[8 * c + 4]) * inv, __uint_as_float(o1[8 * c + 5]) * inv);
        w.w = pack_bf16x2(__uint_as_float(o1[8 * c + 6]) * inv, __uint_as_float(o1[8 * c + 7]) * inv);
        dst[4 + c] = w;
      }
    }
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

int launch_attn_tc(const void* qkv0, int rows0, const void* qkv1, int rows1, int ld, int C, int heads,
                   const int* tiles_dev, int n_tiles, void* out, int ldo, float scale, cudaStream_t stream) {
  CUtensorMap tm0, tm1;
  int rc = make_qkv_tmap(&tm0, qkv0, rows0, 3 * C, ld);
  if (rc) return rc;
  rc = make_qkv_tmap(&tm1, qkv1, rows1, 3 * C, ld);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid(n_tiles, heads);
  attn_tc_kernel<<<grid, ATC_THREADS, ATC_SMEM, stream>>>(tm0, tm1, C, reinterpret_cast<const AttnTileTC*>(tiles_dev),
                                                          reinterpret_cast<bf16*>(out), ldo,
                                                          scale * 1.4426950408889634f);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? MMT_OK : (int)e;
}

}  // namespace mmt
