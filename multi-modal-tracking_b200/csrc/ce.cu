// Candidate elimination (token pruning) of the asymmetric_shared_ce backbone:
// descending sort of the per-modality search-token scores, kept/removed global-index bookkeeping,
// kept-token gather and the final scatter back to the 18x18 grid.
// Reference: candidate_elimination / get_token_from_attn lib/models/mixformer_vit_rgbt/asymmetric_shared_ce.py:22-101,
// _recover_search :427-447, global index init :397-399 (indices are float32 there; kept here as float32 too).
#include "common.cuh"
#include "../../include/mmt_b200.h"

namespace mmt {

constexpr int CE_PAD = 1024;  // >= max search tokens per modality (576 for MixViT-L)

// One CTA per (modality m, sequence b): bitonic sort of (score desc, index asc).
// torch.sort(descending=True) is unstable; for exactly equal scores we order by ascending position,
// which is what a stable sort returns.  Kept tokens stay in score order, like the reference.
__global__ void __launch_bounds__(512)
ce_topk_kernel(const float* __restrict__ scores, int B, int Ls, int keep, const float* __restrict__ gidx_in,
               float* __restrict__ gidx_keep, float* __restrict__ gidx_removed, int* __restrict__ order) {
  __shared__ float key[CE_PAD];
  __shared__ int idx[CE_PAD];
  const int b = blockIdx.x, m = blockIdx.y;
  const float* sc = scores + static_cast<size_t>(b) * 2 * Ls + m * Ls;
  for (int i = threadIdx.x; i < CE_PAD; i += blockDim.x) {
    key[i] = i < Ls ? sc[i] : -INFINITY;
    idx[i] = i < Ls ? i : 0x7fffffff;
  }
  __syncthreads();
  // "a before b" in the final order
  auto before = [](float ka, int ia, float kb, int ib) { return ka > kb || (ka == kb && ia < ib); };
  for (int k = 2; k <= CE_PAD; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < CE_PAD; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const bool up = (i & k) == 0;  // this subsequence sorted in final ("before") order
          const float ka = key[i], kb = key[p];
          const int ia = idx[i], ib = idx[p];
          const bool swap = up ? before(kb, ib, ka, ia) : before(ka, ia, kb, ib);
          if (swap) { key[i] = kb; key[p] = ka; idx[i] = ib; idx[p] = ia; }
        }
      }
      __syncthreads();
    }
  }
  const size_t s = static_cast<size_t>(m) * B + b;  // row in the modality-major [2B, .] tensors
  const float* gi = gidx_in + s * Ls;
  for (int i = threadIdx.x; i < Ls; i += blockDim.x) {
    const int src = idx[i];
    order[s * Ls + i] = src;
    if (i < keep) gidx_keep[s * keep + i] = gi[src];
    else gidx_removed[s * (Ls - keep) + (i - keep)] = gi[src];
  }
}

// x_out[s, r, :] = r < Lt ? x[s, r, :] : x[s, Lt + order[s][r - Lt], :]   (fp32 residual stream rows)
__global__ void ce_gather_kernel(const float* __restrict__ x, int n_tok, int Lt, const int* __restrict__ order, int Ls,
                                 int keep, float* __restrict__ x_out, int C, int nseq) {
  const int nv = C >> 2;
  const int n_out = Lt + keep;
  const size_t total = static_cast<size_t>(nseq) * n_out * nv;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = i % nv;
    const size_t row = i / nv;
    const int s = row / n_out, r = row % n_out;
    const int src = r < Lt ? r : Lt + order[static_cast<size_t>(s) * Ls + (r - Lt)];
    reinterpret_cast<float4*>(x_out)[i] =
        reinterpret_cast<const float4*>(x + (static_cast<size_t>(s) * n_tok + src) * C)[cv];
  }
}

// out[s*Ls0 + int(gidx[s][i]), :] = T(x[s, Lt + i, :]); `out` must have been zero-filled (removed
// positions are zeros in the reference's recovered map).
template <typename T>
__global__ void ce_recover_kernel(const float* __restrict__ x, int n_tok, int Lt, const float* __restrict__ gidx, int Lk,
                                  int Ls0, T* __restrict__ out, int C, int nseq) {
  const int nv = C >> 2;
  const size_t total = static_cast<size_t>(nseq) * Lk * nv;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = i % nv;
    const size_t row = i / nv;
    const int s = row / Lk, k = row % Lk;
    const int pos = static_cast<int>(gidx[static_cast<size_t>(s) * Lk + k]);
    const float4 v = reinterpret_cast<const float4*>(x + (static_cast<size_t>(s) * n_tok + Lt + k) * C)[cv];
    const size_t o = (static_cast<size_t>(s) * Ls0 + pos) * nv + cv;
    if (sizeof(T) == 4) reinterpret_cast<float4*>(out)[o] = v;
    else {
      uint2 p; p.x = pack_bf16x2(v.x, v.y); p.y = pack_bf16x2(v.z, v.w);
      reinterpret_cast<uint2*>(out)[o] = p;
    }
  }
}

static inline int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = 148 * 16;
  return static_cast<int>(g < cap ? (g ? g : 1) : cap);
}

}  // namespace mmt

using namespace mmt;

extern "C" int mmt_ce_topk(const float* scores, int B, int Ls, int keep, const float* gidx_in, float* gidx_keep,
                           float* gidx_removed, int* order, void* stream) {
  MMT_CHECK_ARG(scores && gidx_in && gidx_keep && gidx_removed && order);
  MMT_CHECK_ARG(B > 0 && Ls > 0 && Ls <= CE_PAD && keep > 0 && keep < Ls);
  dim3 grid(B, 2);
  ce_topk_kernel<<<grid, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(scores, B, Ls, keep, gidx_in, gidx_keep,
                                                                            gidx_removed, order);
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_ce_gather_tokens(const float* x, int nseq, int n_tok, int Lt, const int* order, int Ls, int keep,
                                    float* x_out, int C, void* stream) {
  MMT_CHECK_ARG(x && order && x_out && nseq > 0 && n_tok == Lt + Ls && keep > 0 && keep <= Ls && C % 4 == 0);
  const size_t total = static_cast<size_t>(nseq) * (Lt + keep) * (C / 4);
  ce_gather_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, n_tok, Lt, order, Ls,
                                                                                             keep, x_out, C, nseq);
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_ce_recover(const float* x, int nseq, int n_tok, int Lt, const float* gidx, int Lk, int Ls0, void* out,
                              int C, int out_bf16, void* stream) {
  MMT_CHECK_ARG(x && gidx && out && nseq > 0 && n_tok == Lt + Lk && Lk <= Ls0 && C % 4 == 0);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t bytes = static_cast<size_t>(nseq) * Ls0 * C * (out_bf16 ? 2 : 4);
  cudaError_t e = cudaMemsetAsync(out, 0, bytes, s);
  if (e != cudaSuccess) return (int)e;
  const size_t total = static_cast<size_t>(nseq) * Lk * (C / 4);
  if (out_bf16) ce_recover_kernel<bf16><<<grid_for(total, 256), 256, 0, s>>>(x, n_tok, Lt, gidx, Lk, Ls0, reinterpret_cast<bf16*>(out), C, nseq);
  else ce_recover_kernel<float><<<grid_for(total, 256), 256, 0, s>>>(x, n_tok, Lt, gidx, Lk, Ls0, reinterpret_cast<float*>(out), C, nseq);
  MMT_RETURN_LAST_ERROR();
}
