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
//   * attn_tc_kernel   : (attention_tc.cu) bf16 mode - tcgen05 / TMEM / TMA kernel, exact two-pass softmax.
//   * attn_f32_kernel  : parity mode, fp32 FMA, scores materialised in smem, softmax then P*V in the
//                        reference's operation order.
#include "common.cuh"
#include "../../include/mmt_b200.h"

namespace mmt {

constexpr int HD = 64;

struct AttnTile {        // 16 ints (64 B), mirrored by the record layout built in engine.py
  int q_row0, q_rows;    // query rows [q_row0, q_row0 + q_rows), q_rows <= 128, rows of buffer 0
  int out_row0;          // output row of the first query
  int nseg;
  int k_row0[3], k_len[3], k_buf[3];  // key/value segments: rows of buffer k_buf (0 or 1)
  int pad[3];
};
static_assert(sizeof(AttnTile) == 64, "AttnTile must be 16 int32");

// ---------------------------------------------------------------------------------------------- fp32
// grid (tiles, heads, 4): each CTA handles 32 of the tile's (up to) 128 query rows.
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
ce_score_partial_kernel(const T* __restrict__ qbuf, int q_seq_rows, const T* __restrict__ qkv, int n_tok, int k_row_off,
                        int ld, int C, int B, int Lt, int Ls, float scale, int kt_pad, float* __restrict__ partial) {
  // qbuf / q_seq_rows: buffer and rows per sequence-modality of the TEMPLATE rows (queries, row 0 of each sequence block);
  // qkv / n_tok / k_row_off: buffer, rows per sequence-modality and first search row of the SEARCH rows (keys).  The full
  // forward passes the same packed buffer twice (q_seq_rows = n_tok, k_row_off = Lt); the cached-template path reads the
  // queries from the template cache and the keys from the search-only buffer (k_row_off = 0).
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
  const size_t qrow0 = (static_cast<size_t>(qmod) * B + b) * q_seq_rows + (qg - qmod * Lt);
  for (int i = tid; i < 32 * HD; i += 256) {
    const int r = i / HD, d = i % HD;
    sQ[i] = to_f<T>(qbuf[(qrow0 + r) * ld + h * HD + d]);
  }
  const int ktot = 2 * Ls;
  for (int mod = 0; mod < 2; ++mod) {
    const size_t krow0 = (static_cast<size_t>(mod) * B + b) * n_tok + k_row_off;
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

namespace mmt {
int launch_attn_tc(const void* qkv0, int rows0, const void* qkv1, int rows1, int ld, int C, int heads,
                   const int* tiles_dev, int n_tiles, void* out, int ldo, float scale, cudaStream_t stream);
}

extern "C" int mmt_mixattn_fwd(const void* qkv0, int rows0, const void* qkv1, int rows1, int ld, int C, int heads,
                               const int* tiles_dev, int n_tiles, int max_keys, void* out, int ldo, float scale,
                               int is_bf16, void* stream) {
  MMT_CHECK_ARG(qkv0 && tiles_dev && out && n_tiles > 0 && heads > 0 && C == heads * HD && ld >= 3 * C && ldo >= C);
  MMT_CHECK_ARG(rows0 > 0);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const AttnTile* tiles = reinterpret_cast<const AttnTile*>(tiles_dev);
  if (!qkv1) { qkv1 = qkv0; rows1 = rows0; }
  MMT_CHECK_ARG(rows1 > 0);
  if (is_bf16) {
    MMT_CHECK_ARG(ld % 8 == 0 && ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(qkv0) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(qkv1) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0);
    return launch_attn_tc(qkv0, rows0, qkv1, rows1, ld, C, heads, tiles_dev, n_tiles, out, ldo, scale, s);
  } else {
    MMT_CHECK_ARG(max_keys > 0);
    const int kt_pad = max_keys + 1;
    const size_t smem = (F32_QT * HD + 64 * 65 + static_cast<size_t>(F32_QT) * kt_pad) * sizeof(float);
    MMT_CHECK_ARG(smem <= 220 * 1024);
    cudaError_t e = cudaFuncSetAttribute(attn_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(n_tiles, heads, 4);
    attn_f32_kernel<<<grid, 256, smem, s>>>(reinterpret_cast<const float*>(qkv0), reinterpret_cast<const float*>(qkv1),
                                            ld, C, tiles, reinterpret_cast<float*>(out), ldo, scale, kt_pad);
  }
  MMT_RETURN_LAST_ERROR();
}

namespace mmt {
int launch_ce_scores_tc(const void* qbuf, int q_seq_rows, const void* qkv, int n_tok, int k_row_off, int ld, int C, int heads,
                        int B, int Lt, int Ls, float scale, float* partial, int* nqt_out, cudaStream_t stream);
}

extern "C" int mmt_ce_scores_split(const void* qbuf, int q_seq_rows, const void* qkv, int n_tok, int k_row_off, int ld, int C,
                                   int heads, int B, int Lt, int Ls, float scale, float* partial_ws, float* scores,
                                   int is_bf16, void* stream) {
  MMT_CHECK_ARG(qbuf && qkv && partial_ws && scores && B > 0 && heads > 0 && C == heads * HD && ld >= 3 * C);
  MMT_CHECK_ARG(Lt > 0 && Lt % 32 == 0 && Ls > 0 && q_seq_rows >= Lt && k_row_off >= 0 && n_tok >= k_row_off + Ls);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int ktot = 2 * Ls, kt_pad = ktot + 1, nqt = (2 * Lt) / 32;
  const size_t smem = (32 * HD + 64 * 65 + static_cast<size_t>(32) * kt_pad) * sizeof(float);
  MMT_CHECK_ARG(smem <= 220 * 1024);
  dim3 grid(B, heads, nqt);
  cudaError_t e;
  if (is_bf16) {
    // bf16 mode: tensor-core kernel (ce_scores_tc.cu); fewer, larger query tiles -> its own nqt
    MMT_CHECK_ARG(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(qbuf) & 15) == 0);
    int nqt_tc = 0;
    const int rc = launch_ce_scores_tc(qbuf, q_seq_rows, qkv, n_tok, k_row_off, ld, C, heads, B, Lt, Ls, scale, partial_ws,
                                       &nqt_tc, s);
    if (rc) return rc;
    ce_score_reduce_kernel<<<cdiv(B * ktot, 256), 256, 0, s>>>(partial_ws, B, heads, nqt_tc, ktot, 2 * Lt, scores);
    MMT_RETURN_LAST_ERROR();
  } else {
    e = cudaFuncSetAttribute(ce_score_partial_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ce_score_partial_kernel<float><<<grid, 256, smem, s>>>(reinterpret_cast<const float*>(qbuf), q_seq_rows,
                                                           reinterpret_cast<const float*>(qkv), n_tok, k_row_off, ld, C, B, Lt,
                                                           Ls, scale, kt_pad, partial_ws);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  ce_score_reduce_kernel<<<cdiv(B * ktot, 256), 256, 0, s>>>(partial_ws, B, heads, nqt, ktot, 2 * Lt, scores);
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_ce_scores(const void* qkv, int ld, int C, int heads, int B, int n_tok, int Lt, int Ls, float scale,
                             float* partial_ws, float* scores, int is_bf16, void* stream) {
  MMT_CHECK_ARG(n_tok == Lt + Ls);
  return mmt_ce_scores_split(qkv, n_tok, qkv, n_tok, Lt, ld, C, heads, B, Lt, Ls, scale, partial_ws, scores, is_bf16, stream);
}
