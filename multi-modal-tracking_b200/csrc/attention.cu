// Asymmetric mixed attention of the MixViT backbone (all variants) over the packed qkv buffer.
//
// Reference: Attention.forward lib/models/mixformer_vit/mixformer.py:51-77 (template rows attend the
// template keys only, search rows attend template + search keys), the cross-modal variant
// lib/models/mixformer_vit_rgbt/asymmetric_shared.py:55-104 / asymmetric_shared_ce.py:146-207 (search
// rows of one modality attend both modalities' template keys + their own search keys) and the cached
// template path forward_test mixformer.py:79-93.  There is no mask tensor in the reference: asymmetry is
// expressed by WHICH key rows a query tile reads.  Here that is a per-tile list of up to three key
// segments (row ranges of a qkv buffer), built once per (variant, batch) on the host.
//
// qkv layout: row-major [rows, 3C], q at column h*64, k at C + h*64, v at 2C + h*64 (nn.Linear(C, 3C)
// output reshaped [B,N,3,H,hd], mixformer.py:56-57).  head_dim is 64 for every shipped model.
//
//   * attn_bf16_kernel : flash-style online softmax, bf16 mma.sync m16n8k16 tensor-core contractions,
//                        cp.async double-buffered K/V blocks, fp32 softmax state.  (The tcgen05/TMEM
//                        version of this kernel is the next step of the build; the GEMMs - 93% of the
//                        FLOPs - already run on tcgen05.)
//   * attn_f32_kernel  : parity mode, fp32 FMA, scores materialised in smem, softmax then P*V in the
//                        reference's operation order.
#include "common.cuh"
#include "../../include/mmt_b200.h"

namespace mmt {

constexpr int HD = 64;

struct AttnTile {        // 16 ints (64 B), mirrored by the record layout built in engine.py
  int q_row0, q_rows;    // query rows [q_row0, q_row0 + q_rows), q_rows <= 64, rows of buffer 0
  int out_row0;          // output row of the first query
  int nseg;
  int k_row0[3], k_len[3], k_buf[3];  // key/value segments: rows of buffer k_buf (0 or 1)
  int pad[3];
};
static_assert(sizeof(AttnTile) == 64, "AttnTile must be 16 int32");

// ---------------------------------------------------------------------------------------------- bf16
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// smem tile of 64 rows x 64 bf16 (128 B rows), 16-byte chunks XOR-swizzled by (row & 7)
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

// Load rows [row0, row0 + nrows) (nrows <= 64; the rest is zero-filled) of a 64-wide head slice.
__device__ __forceinline__ void load_tile_async(uint32_t smem, const bf16* base, int ld, int row0, int nrows) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int id = threadIdx.x + i * 128;
    const int r = id >> 3, ch = id & 7;
    const bool ok = r < nrows;
    const bf16* src = base + static_cast<size_t>(row0 + (ok ? r : 0)) * ld + ch * 8;
    cp_async16(smem + tile_off(r, ch), src, ok);
  }
}

__global__ void __launch_bounds__(128)
attn_bf16_kernel(const bf16* __restrict__ qkv0, const bf16* __restrict__ qkv1, int ld, int C,
                 const AttnTile* __restrict__ tiles, bf16* __restrict__ out, int ldo, float scale_log2e) {
  __shared__ __align__(128) uint8_t sQ[64 * 128];
  __shared__ __align__(128) uint8_t sK[2][64 * 128];
  __shared__ __align__(128) uint8_t sV[2][64 * 128];
  const AttnTile t = tiles[blockIdx.x];
  const int h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sq = static_cast<uint32_t>(__cvta_generic_to_shared(sQ));
  const uint32_t sk[2] = {static_cast<uint32_t>(__cvta_generic_to_shared(sK[0])),
                          static_cast<uint32_t>(__cvta_generic_to_shared(sK[1]))};
  const uint32_t sv[2] = {static_cast<uint32_t>(__cvta_generic_to_shared(sV[0])),
                          static_cast<uint32_t>(__cvta_generic_to_shared(sV[1]))};

  int nblk[3];
  int total_blk = 0;
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    nblk[s] = s < t.nseg ? (t.k_len[s] + 63) >> 6 : 0;
    total_blk += nblk[s];
  }
  auto locate = [&](int blk, int& row0, int& len, const bf16*& base) {
    int s = 0;
    if (blk >= nblk[0]) { blk -= nblk[0]; s = 1; if (blk >= nblk[1]) { blk -= nblk[1]; s = 2; } }
    row0 = t.k_row0[s] + blk * 64;
    len = min(64, t.k_len[s] - blk * 64);
    base = t.k_buf[s] ? qkv1 : qkv0;
  };
  auto issue_kv = [&](int blk, int buf) {
    int row0, len;
    const bf16* base;
    locate(blk, row0, len, base);
    load_tile_async(sk[buf], base + C + h * HD, ld, row0, len);
    load_tile_async(sv[buf], base + 2 * C + h * HD, ld, row0, len);
  };

  load_tile_async(sq, qkv0 + h * HD, ld, t.q_row0, t.q_rows);
  issue_kv(0, 0);
  cp_async_commit();

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[4][4];

  for (int blk = 0; blk < total_blk; ++blk) {
    const int buf = blk & 1;
    if (blk + 1 < total_blk) {
      issue_kv(blk + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (blk == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        ldsm_x4(sq + tile_off(warp * 16 + (lane & 15), 2 * ks + (lane >> 4)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
    int row0, len;
    const bf16* base_unused;
    locate(blk, row0, len, base_unused);

    // S = Q K^T (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-key tiles
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sk[buf] + tile_off(np * 16 + (lane & 7) + ((lane >> 4) << 3), 2 * ks + ((lane >> 3) & 1)), b0, b1, b2, b3);
        mma16816(s[2 * np], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
        mma16816(s[2 * np + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
      }
    }
    // mask keys beyond the segment end (only the last block of a segment is partial)
    if (len < 64) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int j = nt * 8 + 2 * (lane & 3);
        if (j >= len) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
        if (j + 1 >= len) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      }
    }
    // online softmax (rows lane/4 and lane/4 + 8 of this warp's 16)
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      corr[r] = exp2f((m_run[r] - mx[r]) * scale_log2e);
      m_run[r] = mx[r];
    }
    uint32_t pf[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f((s[nt][0] - mx[0]) * scale_log2e), p1 = exp2f((s[nt][1] - mx[0]) * scale_log2e);
      const float p2 = exp2f((s[nt][2] - mx[1]) * scale_log2e), p3 = exp2f((s[nt][3] - mx[1]) * scale_log2e);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[nt][0] = pack_bf16x2(p0, p1);
      pf[nt][1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      o[nt][0] *= corr[0]; o[nt][1] *= corr[0]; o[nt][2] *= corr[1]; o[nt][3] *= corr[1];
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // 16-key steps
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {  // pairs of 8-wide d tiles
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sv[buf] + tile_off(kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), 2 * dp + (lane >> 4)), b0, b1, b2, b3);
        mma16816(o[2 * dp], pf[2 * kk][0], pf[2 * kk][1], pf[2 * kk + 1][0], pf[2 * kk + 1][1], b0, b1);
        mma16816(o[2 * dp + 1], pf[2 * kk][0], pf[2 * kk][1], pf[2 * kk + 1][0], pf[2 * kk + 1][1], b2, b3);
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled two iterations later
  }

  // finalise: row sums across the quad, normalise, store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int r0 = warp * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = h * HD + nt * 8 + 2 * (lane & 3);
    if (r0 < t.q_rows)
      *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(t.out_row0 + r0) * ldo + col) =
          pack_bf16x2(o[nt][0] * inv0, o[nt][1] * inv0);
    if (r1 < t.q_rows)
      *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(t.out_row0 + r1) * ldo + col) =
          pack_bf16x2(o[nt][2] * inv1, o[nt][3] * inv1);
  }
}

// ---------------------------------------------------------------------------------------------- fp32
// grid (tiles, heads, 2): each CTA handles 32 of the tile's (up to) 64 query rows.
constexpr int F32_QT = 32;

__global__ void __launch_bounds__(256)
attn_f32_kernel(const float* __restrict__ qkv0, const float* __restrict__ qkv1, int ld, int C,
                const AttnTile* __restrict__ tiles, float* __restrict__ out, int ldo, float scale, int kt_pad) {
  extern __shared__ float sm[];
  float* sQ = sm;                       // [32][64]
  float* sKV = sQ + F32_QT * HD;        // [64][65]
  float* sS = sKV + 64 * 65;            // [32][kt_pad]
  const AttnTile t = tiles[blockIdx.x];
  const int h = blockIdx.y;
  const int qoff = blockIdx.z * F32_QT;
  if (qoff >= t.q_rows) return;
  const int qn = min(F32_QT, t.q_rows - qoff);
  const int tid = threadIdx.x;

  for (int i = tid; i < F32_QT * HD; i += 256) {
    const int r = i / HD, d = i % HD;
    sQ[i] = r < qn ? qkv0[static_cast<size_t>(t.q_row0 + qoff + r) * ld + h * HD + d] : 0.f;
  }
  int ktot = 0;
  for (int s = 0; s < t.nseg; ++s) ktot += t.k_len[s];

  // phase 1: scores
  int kbase = 0;
  for (int s = 0; s < t.nseg; ++s) {
    const float* base = t.k_buf[s] ? qkv1 : qkv0;
    for (int k0 = 0; k0 < t.k_len[s]; k0 += 64) {
      const int len = min(64, t.k_len[s] - k0);
      __syncthreads();
      for (int i = tid; i < 64 * HD; i += 256) {
        const int r = i / HD, d = i % HD;
        sKV[r * 65 + d] = r < len ? base[static_cast<size_t>(t.k_row0[s] + k0 + r) * ld + C + h * HD + d] : 0.f;
      }
      __syncthreads();
      const int r = tid >> 3;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = (tid & 7) + 8 * i;
        float acc = 0.f;
#pragma unroll 16
        for (int d = 0; d < HD; ++d) acc = fmaf(sQ[r * HD + d], sKV[j * 65 + d], acc);
        if (j < len) sS[r * kt_pad + kbase + k0 + j] = acc * scale;
      }
    }
    kbase += t.k_len[s];
  }
  __syncthreads();
  // phase 2: row softmax (warp per row)
  const int warp = tid >> 5, lane = tid & 31;
  for (int r = warp; r < F32_QT; r += 8) {
    float* row = sS + r * kt_pad;
    float mx = -INFINITY;
    for (int j = lane; j < ktot; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < ktot; j += 32) {
      const float e = expf(row[j] - mx);
      row[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < ktot; j += 32) row[j] *= inv;
  }
  // phase 3: O = P V
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  kbase = 0;
  const int r = tid >> 3;
  for (int s = 0; s < t.nseg; ++s) {
    const float* base = t.k_buf[s] ? qkv1 : qkv0;
    for (int k0 = 0; k0 < t.k_len[s]; k0 += 64) {
      const int len = min(64, t.k_len[s] - k0);
      __syncthreads();
      for (int i = tid; i < 64 * HD; i += 256) {
        const int rr = i / HD, d = i % HD;
        sKV[rr * 65 + d] = rr < len ? base[static_cast<size_t>(t.k_row0[s] + k0 + rr) * ld + 2 * C + h * HD + d] : 0.f;
      }
      __syncthreads();
      for (int j = 0; j < len; ++j) {
        const float p = sS[r * kt_pad + kbase + k0 + j];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(p, sKV[j * 65 + (tid & 7) + 8 * i], acc[i]);
      }
    }
    kbase += t.k_len[s];
  }
  if (r < qn) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      out[static_cast<size_t>(t.out_row0 + qoff + r) * ldo + h * HD + (tid & 7) + 8 * i] = acc[i];
  }
}

// ---------------------------------------------------------------------------------------------- CE scores
// Candidate-elimination score of the search tokens (asymmetric_shared_ce.py:202-205 + :91-92):
//   attn_t2s = softmax_over_all_2Ls_keys( [q_mt_V; q_mt_I] . [k_s_V; k_s_I]^T * scale )   [B, H, 2*Lt, 2*Ls]
//   score    = attn_t2s.mean(dim=2).mean(dim=1)                                           [B, 2*Ls]
// computed in fp32 FMA in both modes (the ranking is what must match the reference), without ever
// materialising attn_t2s in HBM.  Stage 1 (this kernel): per (b, head, 32-row query tile) column sums of the
// softmax rows -> partial[b][h][qt][2Ls].  Stage 2 (ce_score_reduce_kernel): fixed-order reduction (mean over the
// 2*Lt rows first, then over heads, like the reference), so the result is run-to-run deterministic.
template <typename T>
__global__ void __launch_bounds__(256)
ce_score_partial_kernel(const T* __restrict__ qkv, int ld, int C, int B, int n_tok, int Lt, int Ls,
                        float scale, int kt_pad, float* __restrict__ partial) {
  extern __shared__ float sm[];
  float* sQ = sm;                 // [32][64]
  float* sKV = sQ + 32 * HD;      // [64][65]
  float* sS = sKV + 64 * 65;      // [32][kt_pad]
  const int b = blockIdx.x, h = blockIdx.y, qt = blockIdx.z;
  const int nqt = (2 * Lt) / 32;  // Lt is a multiple of 32 (128)
  const int tid = threadIdx.x;
  // query rows: tile qt covers template rows [qt*32, qt*32+32) of the stacked [V templates; I templates]
  const int qg = qt * 32;
  const int qmod = qg >= Lt;                       // 0: RGB templates, 1: TIR templates
  const size_t qrow0 = (static_cast<size_t>(qmod) * B + b) * n_tok + (qg - qmod * Lt);
  for (int i = tid; i < 32 * HD; i += 256) {
    const int r = i / HD, d = i % HD;
    sQ[i] = to_f<T>(qkv[(qrow0 + r) * ld + h * HD + d]);
  }
  const int ktot = 2 * Ls;
  for (int mod = 0; mod < 2; ++mod) {
    const size_t krow0 = (static_cast<size_t>(mod) * B + b) * n_tok + Lt;
    for (int k0 = 0; k0 < Ls; k0 += 64) {
      const int len = min(64, Ls - k0);
      __syncthreads();
      for (int i = tid; i < 64 * HD; i += 256) {
        const int r = i / HD, d = i % HD;
        sKV[r * 65 + d] = r < len ? to_f<T>(qkv[(krow0 + k0 + r) * ld + C + h * HD + d]) : 0.f;
      }
      __syncthreads();
      const int r = tid >> 3;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = (tid & 7) + 8 * i;
        float acc = 0.f;
#pragma unroll 16
        for (int d = 0; d < HD; ++d) acc = fmaf(sQ[r * HD + d], sKV[j * 65 + d], acc);
        if (j < len) sS[r * kt_pad + mod * Ls + k0 + j] = acc * scale;
      }
    }
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int r = warp; r < 32; r += 8) {
    float* row = sS + r * kt_pad;
    float mx = -INFINITY;
    for (int j = lane; j < ktot; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < ktot; j += 32) {
      const float e = expf(row[j] - mx);
      row[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < ktot; j += 32) row[j] *= inv;
  }
  __syncthreads();
  float* dst = partial + ((static_cast<size_t>(b) * gridDim.y + h) * nqt + qt) * ktot;
  for (int j = tid; j < ktot; j += 256) {
    float acc = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) acc += sS[r * kt_pad + j];
    dst[j] = acc;
  }
}

__global__ void ce_score_reduce_kernel(const float* __restrict__ partial, int B, int H, int nqt, int ktot, int nrows,
                                       float* __restrict__ scores) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * ktot) return;
  const int b = i / ktot, j = i % ktot;
  float hs = 0.f;
  for (int h = 0; h < H; ++h) {
    float rs = 0.f;
    for (int q = 0; q < nqt; ++q) rs += partial[((static_cast<size_t>(b) * H + h) * nqt + q) * ktot + j];
    hs += rs / nrows;        // mean(dim=2)
  }
  scores[i] = hs / H;        // mean(dim=1)
}

}  // namespace mmt

using namespace mmt;

extern "C" int mmt_mixattn_fwd(const void* qkv0, const void* qkv1, int ld, int C, int heads, const int* tiles_dev,
                               int n_tiles, int max_keys, void* out, int ldo, float scale, int is_bf16, void* stream) {
  MMT_CHECK_ARG(qkv0 && tiles_dev && out && n_tiles > 0 && heads > 0 && C == heads * HD && ld >= 3 * C && ldo >= C);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const AttnTile* tiles = reinterpret_cast<const AttnTile*>(tiles_dev);
  if (!qkv1) qkv1 = qkv0;
  if (is_bf16) {
    MMT_CHECK_ARG(ld % 8 == 0 && ldo % 2 == 0 && (reinterpret_cast<uintptr_t>(qkv0) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(qkv1) & 15) == 0);
    dim3 grid(n_tiles, heads);
    attn_bf16_kernel<<<grid, 128, 0, s>>>(reinterpret_cast<const bf16*>(qkv0), reinterpret_cast<const bf16*>(qkv1), ld,
                                          C, tiles, reinterpret_cast<bf16*>(out), ldo, scale * 1.4426950408889634f);
  } else {
    MMT_CHECK_ARG(max_keys > 0);
    const int kt_pad = max_keys + 1;
    const size_t smem = (F32_QT * HD + 64 * 65 + static_cast<size_t>(F32_QT) * kt_pad) * sizeof(float);
    MMT_CHECK_ARG(smem <= 220 * 1024);
    cudaError_t e = cudaFuncSetAttribute(attn_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(n_tiles, heads, 2);
    attn_f32_kernel<<<grid, 256, smem, s>>>(reinterpret_cast<const float*>(qkv0), reinterpret_cast<const float*>(qkv1),
                                            ld, C, tiles, reinterpret_cast<float*>(out), ldo, scale, kt_pad);
  }
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_ce_scores(const void* qkv, int ld, int C, int heads, int B, int n_tok, int Lt, int Ls, float scale,
                             float* partial_ws, float* scores, int is_bf16, void* stream) {
  MMT_CHECK_ARG(qkv && partial_ws && scores && B > 0 && heads > 0 && C == heads * HD && ld >= 3 * C);
  MMT_CHECK_ARG(Lt > 0 && Lt % 32 == 0 && Ls > 0 && n_tok == Lt + Ls);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int ktot = 2 * Ls, kt_pad = ktot + 1, nqt = (2 * Lt) / 32;
  const size_t smem = (32 * HD + 64 * 65 + static_cast<size_t>(32) * kt_pad) * sizeof(float);
  MMT_CHECK_ARG(smem <= 220 * 1024);
  dim3 grid(B, heads, nqt);
  cudaError_t e;
  if (is_bf16) {
    e = cudaFuncSetAttribute(ce_score_partial_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ce_score_partial_kernel<bf16><<<grid, 256, smem, s>>>(reinterpret_cast<const bf16*>(qkv), ld, C, B, n_tok, Lt, Ls, scale, kt_pad, partial_ws);
  } else {
    e = cudaFuncSetAttribute(ce_score_partial_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ce_score_partial_kernel<float><<<grid, 256, smem, s>>>(reinterpret_cast<const float*>(qkv), ld, C, B, n_tok, Lt, Ls, scale, kt_pad, partial_ws);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  ce_score_reduce_kernel<<<cdiv(B * ktot, 256), 256, 0, s>>>(partial_ws, B, heads, nqt, ktot, 2 * Lt, scores);
  MMT_RETURN_LAST_ERROR();
}
