// Multi-scale deformable attention sampling (forward only).
//
// Replaces the reference's native op MultiScaleDeformableAttention.ms_deform_attn_forward
// (lib/models/mixformer_vit_rgbt/deformable_attention/ops/src/cuda/ms_deform_attn_cuda.cu:20-80,
// kernel ms_deformable_im2col_gpu_kernel ms_deform_im2col_cuda.cuh:237-299, bilinear :33-84):
//   out[n, q, m, :] = sum_{l,p} w[n,q,m,l,p] * bilinear(value[n, level l, m, :], loc[n,q,m,l,p])
// with pixel coords h = loc_y*H - 0.5, w = loc_x*W - 0.5, contributions only for -1 < h < H, -1 < w < W,
// zero outside the map.  Unlike the reference there is no batch % im2col_step restriction.
//
// One warp per (n, q, m): the 32 lanes cover the head channels, so every corner fetch is one
// coalesced row segment; sampling parameters are warp-uniform broadcast loads.
#include "common.cuh"
#include "../../include/mmt_b200.h"

namespace mmt {

constexpr int MSDA_MAX_LEVELS = 8;
struct MsdaLevels {
  int h[MSDA_MAX_LEVELS];
  int w[MSDA_MAX_LEVELS];
  int start[MSDA_MAX_LEVELS];
};

template <typename T, int VEC>
struct ChanVec;
template <> struct ChanVec<float, 1> { using type = float; };
template <> struct ChanVec<float, 2> { using type = float2; };
template <> struct ChanVec<bf16, 1> { using type = bf16; };
template <> struct ChanVec<bf16, 2> { using type = __nv_bfloat162; };

__device__ __forceinline__ void vacc(float (&a)[2], float w, float v) { a[0] = fmaf(w, v, a[0]); }
__device__ __forceinline__ void vacc(float (&a)[2], float w, bf16 v) { a[0] = fmaf(w, __bfloat162float(v), a[0]); }
__device__ __forceinline__ void vacc(float (&a)[2], float w, float2 v) {
  a[0] = fmaf(w, v.x, a[0]);
  a[1] = fmaf(w, v.y, a[1]);
}
__device__ __forceinline__ void vacc(float (&a)[2], float w, __nv_bfloat162 v) {
  const float2 f = __bfloat1622float2(v);
  a[0] = fmaf(w, f.x, a[0]);
  a[1] = fmaf(w, f.y, a[1]);
}

// Accumulate attn * bilinear(value map of one level/head at (h_im, w_im)) for this lane's channels.
// vbase points at value[n, level_start, m, 0]; consecutive pixels are `pix_stride` elements apart.
template <typename T, int VEC>
__device__ __forceinline__ void bilinear_acc(const T* __restrict__ vbase, int H, int W, int pix_stride, float h_im,
                                             float w_im, float attn, int d, float (&acc)[2]) {
  using V = typename ChanVec<T, VEC>::type;
  if (!(h_im > -1.f && w_im > -1.f && h_im < H && w_im < W)) return;
  const int h_low = static_cast<int>(floorf(h_im)), w_low = static_cast<int>(floorf(w_im));
  const int h_high = h_low + 1, w_high = w_low + 1;
  const float lh = h_im - h_low, lw = w_im - w_low, hh = 1.f - lh, hw = 1.f - lw;
  float r[2] = {0.f, 0.f};
  if (h_low >= 0 && w_low >= 0)
    vacc(r, hh * hw, *reinterpret_cast<const V*>(vbase + static_cast<size_t>(h_low * W + w_low) * pix_stride + d));
  if (h_low >= 0 && w_high <= W - 1)
    vacc(r, hh * lw, *reinterpret_cast<const V*>(vbase + static_cast<size_t>(h_low * W + w_high) * pix_stride + d));
  if (h_high <= H - 1 && w_low >= 0)
    vacc(r, lh * hw, *reinterpret_cast<const V*>(vbase + static_cast<size_t>(h_high * W + w_low) * pix_stride + d));
  if (h_high <= H - 1 && w_high <= W - 1)
    vacc(r, lh * lw, *reinterpret_cast<const V*>(vbase + static_cast<size_t>(h_high * W + w_high) * pix_stride + d));
  acc[0] = fmaf(attn, r[0], acc[0]);
  acc[1] = fmaf(attn, r[1], acc[1]);
}

template <typename T> __device__ __forceinline__ void store_vec(T* p, const float (&a)[2], int vec);
template <> __device__ __forceinline__ void store_vec<float>(float* p, const float (&a)[2], int vec) {
  if (vec == 2) *reinterpret_cast<float2*>(p) = make_float2(a[0], a[1]);
  else *p = a[0];
}
template <> __device__ __forceinline__ void store_vec<bf16>(bf16* p, const float (&a)[2], int vec) {
  if (vec == 2) *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a[0], a[1]);
  else *p = __float2bfloat16_rn(a[0]);
}

// Generic op, reference tensor layout: value [N,S,M,D], loc [N,Lq,M,L,P,2] fp32, attn [N,Lq,M,L,P] fp32,
// out [N,Lq,M*D].
template <typename T, int VEC>
__global__ void msda_generic_kernel(const T* __restrict__ value, MsdaLevels lv, const float* __restrict__ loc,
                                    const float* __restrict__ attn, T* __restrict__ out, int N, int S, int M, int D,
                                    int L, int Lq, int P) {
  const size_t warp = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const size_t total = static_cast<size_t>(N) * Lq * M;
  if (warp >= total) return;
  const int m = warp % M;
  const size_t nq = warp / M;
  const int n = nq / Lq;
  const float* lp = loc + warp * L * P * 2;
  const float* ap = attn + warp * L * P;
  for (int d = lane * VEC; d < D; d += 32 * VEC) {
    float acc[2] = {0.f, 0.f};
    for (int l = 0; l < L; ++l) {
      const int H = lv.h[l], W = lv.w[l];
      const T* vbase = value + (static_cast<size_t>(n) * S + lv.start[l]) * M * D + m * D;
      for (int p = 0; p < P; ++p) {
        const float lx = __ldg(lp + (l * P + p) * 2), ly = __ldg(lp + (l * P + p) * 2 + 1);
        bilinear_acc<T, VEC>(vbase, H, W, M * D, ly * H - 0.5f, lx * W - 0.5f, __ldg(ap + l * P + p), d, acc);
      }
    }
    store_vec<T>(out + warp * D + d, acc, VEC);
  }
}

// Fused bimodal variant used inside the RGB-T fusion encoder (two levels = two modalities of identical
// H x W): the sampling offsets and attention logits come straight from the projection GEMM
// (offw[b*HW + pos, :] = [M*2*P*2 offsets | M*2*P logits]); reference points, offset normalisation, the
// softmax over the 2*P samples and the sampling are done here.  The RGB query and the TIR query at the
// same position share offsets AND weights (ms_deform_attn_bimodal.py:108-111: cat([x, x], dim=1)), so the
// result is computed once and written to both token rows.
template <typename T>
__global__ void msda_bimodal_kernel(const T* __restrict__ value, const float* __restrict__ offw, int ldo_w,
                                    T* __restrict__ out, int B, int H, int W, int M, int P) {
  constexpr int D = 64;
  const int HW = H * W;
  const size_t warp = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const size_t total = static_cast<size_t>(B) * HW * M;
  if (warp >= total) return;
  const int m = warp % M;
  const size_t bp = warp / M;
  const int pos = bp % HW;
  const int b = bp / HW;
  const int py = pos / W, px = pos % W;
  const float* row = offw + bp * ldo_w;
  const float* offs = row + m * (2 * P * 2);
  const float* logits = row + M * 2 * P * 2 + m * (2 * P);
  // softmax over the 2*P logits (F.softmax(attention_weights, -1), ms_deform_attn_bimodal.py:112)
  float mx = -INFINITY;
  for (int i = 0; i < 2 * P; ++i) mx = fmaxf(mx, __ldg(logits + i));
  float den = 0.f;
  for (int i = 0; i < 2 * P; ++i) den += expf(__ldg(logits + i) - mx);
  const float inv_den = 1.f / den;
  // reference point of this query for both levels: ((x + 0.5)/W, (y + 0.5)/H)  (deformable_encoder.py:167-184)
  const float ref_x = (px + 0.5f) / W, ref_y = (py + 0.5f) / H;
  float acc[2] = {0.f, 0.f};
  const int d = lane * 2;
  for (int l = 0; l < 2; ++l) {
    const T* vbase = value + (static_cast<size_t>(b) * 2 * HW + l * HW) * M * D + m * D;
    for (int p = 0; p < P; ++p) {
      const float ox = __ldg(offs + (l * P + p) * 2), oy = __ldg(offs + (l * P + p) * 2 + 1);
      const float lx = ref_x + ox / W, ly = ref_y + oy / H;
      const float a = expf(__ldg(logits + l * P + p) - mx) * inv_den;
      bilinear_acc<T, 2>(vbase, H, W, M * D, ly * H - 0.5f, lx * W - 0.5f, a, d, acc);
    }
  }
  T* o = out + (static_cast<size_t>(b) * 2 * HW + pos) * M * D + m * D + d;
  store_vec<T>(o, acc, 2);
  store_vec<T>(o + static_cast<size_t>(HW) * M * D, acc, 2);
}

// bf16 fast path of the fused bimodal op for P = 4 (2 levels x 4 points = 8 samples): EIGHT lanes per (b, pos, head).
// Lane j of a group owns sample j's scalar work (logit, offset, pixel coordinates) and channels [8j, 8j + 8) of the
// head: the per-sample scalars travel by width-8 shuffles, every corner fetch is one 16-byte load per lane (a 128-byte
// row segment per group), and a warp covers four heads - a quarter of the instructions of the warp-per-head mapping
// above.  Arithmetic and summation order per channel are the same as in msda_bimodal_kernel (bit-identical results).
__device__ __forceinline__ void acc_bf16x8(float (&r)[8], float w, const uint4& v) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&u[i]);
    const float2 f = __bfloat1622float2(h);
    r[2 * i] = fmaf(w, f.x, r[2 * i]);
    r[2 * i + 1] = fmaf(w, f.y, r[2 * i + 1]);
  }
}

__global__ void __launch_bounds__(256)
msda_bimodal_bf16_p4_kernel(const bf16* __restrict__ value, const float* __restrict__ offw, int ldo_w,
                            bf16* __restrict__ out, int B, int H, int W, int M) {
  constexpr int D = 64, P = 4, NS = 8;
  const int HW = H * W;
  const size_t total = static_cast<size_t>(B) * HW * M;
  size_t grp = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 3;
  const int gl = threadIdx.x & 7;
  const bool live = grp < total;
  if (!live) grp = total - 1;                 // keep the whole warp in the shuffles
  const int m = grp % M;
  const size_t bp = grp / M;
  const int pos = bp % HW;
  const int b = bp / HW;
  const int py = pos / W, px = pos % W;
  const float* row = offw + bp * ldo_w;
  const float logit = __ldg(row + M * 2 * P * 2 + m * NS + gl);
  const float2 off = __ldg(reinterpret_cast<const float2*>(row + m * (NS * 2)) + gl);
  float mx = logit;
#pragma unroll
  for (int o = 1; o < NS; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o, NS));
  const float e = expf(logit - mx);
  float den = 0.f;
#pragma unroll
  for (int i = 0; i < NS; ++i) den += __shfl_sync(0xffffffffu, e, i, NS);     // same order as the sequential sum
  const float a_own = e * (1.f / den);
  const float ref_x = (px + 0.5f) / W, ref_y = (py + 0.5f) / H;
  const float lx = ref_x + off.x / W, ly = ref_y + off.y / H;
  const float h_own = ly * H - 0.5f, w_own = lx * W - 0.5f;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const int pix_stride = M * D;
#pragma unroll
  for (int sidx = 0; sidx < NS; ++sidx) {
    const float h_im = __shfl_sync(0xffffffffu, h_own, sidx, NS);
    const float w_im = __shfl_sync(0xffffffffu, w_own, sidx, NS);
    const float a = __shfl_sync(0xffffffffu, a_own, sidx, NS);
    if (!(h_im > -1.f && w_im > -1.f && h_im < H && w_im < W)) continue;
    const bf16* vbase = value + (static_cast<size_t>(b) * 2 * HW + (sidx / P) * HW) * pix_stride + m * D + gl * 8;
    const int h_low = static_cast<int>(floorf(h_im)), w_low = static_cast<int>(floorf(w_im));
    const int h_high = h_low + 1, w_high = w_low + 1;
    const float lh = h_im - h_low, lw = w_im - w_low, hh = 1.f - lh, hw = 1.f - lw;
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = 0.f;
    auto corner = [&](int y, int x, float wgt) {
      acc_bf16x8(r, wgt, __ldg(reinterpret_cast<const uint4*>(vbase + static_cast<size_t>(y * W + x) * pix_stride)));
    };
    if (h_low >= 0 && w_low >= 0) corner(h_low, w_low, hh * hw);
    if (h_low >= 0 && w_high <= W - 1) corner(h_low, w_high, hh * lw);
    if (h_high <= H - 1 && w_low >= 0) corner(h_high, w_low, lh * hw);
    if (h_high <= H - 1 && w_high <= W - 1) corner(h_high, w_high, lh * lw);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(a, r[i], acc[i]);
  }
  if (!live) return;
  uint4 o;
  o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
  o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
  bf16* op = out + (static_cast<size_t>(b) * 2 * HW + pos) * pix_stride + m * D + gl * 8;
  *reinterpret_cast<uint4*>(op) = o;
  *reinterpret_cast<uint4*>(op + static_cast<size_t>(HW) * pix_stride) = o;
}

}  // namespace mmt

using namespace mmt;

extern "C" int mmt_msda_fwd(const void* value, const int* level_hw_host, const float* sampling_loc,
                            const float* attn_weight, void* out, int N, int S, int M, int D, int L, int Lq, int P,
                            int is_bf16, void* stream) {
  MMT_CHECK_ARG(value && level_hw_host && sampling_loc && attn_weight && out);
  MMT_CHECK_ARG(N > 0 && S > 0 && M > 0 && D > 0 && L > 0 && L <= MSDA_MAX_LEVELS && Lq > 0 && P > 0);
  MsdaLevels lv;
  int acc = 0;
  for (int l = 0; l < L; ++l) {
    lv.h[l] = level_hw_host[2 * l];
    lv.w[l] = level_hw_host[2 * l + 1];
    lv.start[l] = acc;
    acc += lv.h[l] * lv.w[l];
  }
  MMT_CHECK_ARG(acc == S);  // same assertion as ms_deform_attn_bimodal.py:101
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t warps = static_cast<size_t>(N) * Lq * M;
  const int block = 256;
  const int grid = static_cast<int>((warps * 32 + block - 1) / block);
  const bool vec2 = (D % 2 == 0) && ((reinterpret_cast<uintptr_t>(value) & 7) == 0) &&
                    ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
  if (is_bf16) {
    if (vec2) msda_generic_kernel<bf16, 2><<<grid, block, 0, s>>>(reinterpret_cast<const bf16*>(value), lv, sampling_loc, attn_weight, reinterpret_cast<bf16*>(out), N, S, M, D, L, Lq, P);
    else msda_generic_kernel<bf16, 1><<<grid, block, 0, s>>>(reinterpret_cast<const bf16*>(value), lv, sampling_loc, attn_weight, reinterpret_cast<bf16*>(out), N, S, M, D, L, Lq, P);
  } else {
    if (vec2) msda_generic_kernel<float, 2><<<grid, block, 0, s>>>(reinterpret_cast<const float*>(value), lv, sampling_loc, attn_weight, reinterpret_cast<float*>(out), N, S, M, D, L, Lq, P);
    else msda_generic_kernel<float, 1><<<grid, block, 0, s>>>(reinterpret_cast<const float*>(value), lv, sampling_loc, attn_weight, reinterpret_cast<float*>(out), N, S, M, D, L, Lq, P);
  }
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_msda_bimodal_fwd(const void* value, const float* offw, int ld_offw, void* out, int B, int H, int W,
                                    int M, int D, int P, int is_bf16, void* stream) {
  MMT_CHECK_ARG(value && offw && out && B > 0 && H > 0 && W > 0 && M > 0 && P > 0);
  MMT_CHECK_ARG(D == 64 && ld_offw >= M * 2 * P * 3);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t warps = static_cast<size_t>(B) * H * W * M;
  const int block = 256;
  const int grid = static_cast<int>((warps * 32 + block - 1) / block);
  const bool al16 = ((reinterpret_cast<uintptr_t>(value) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                    ((reinterpret_cast<uintptr_t>(offw) & 7) == 0) && (ld_offw % 2 == 0);
  if (is_bf16 && P == 4 && al16) {
    const size_t groups = warps;                 // one 8-lane group per (b, pos, head)
    const int grid8 = static_cast<int>((groups * 8 + block - 1) / block);
    msda_bimodal_bf16_p4_kernel<<<grid8, block, 0, s>>>(reinterpret_cast<const bf16*>(value), offw, ld_offw,
                                                        reinterpret_cast<bf16*>(out), B, H, W, M);
    MMT_RETURN_LAST_ERROR();
  }
  if (is_bf16) msda_bimodal_kernel<bf16><<<grid, block, 0, s>>>(reinterpret_cast<const bf16*>(value), offw, ld_offw, reinterpret_cast<bf16*>(out), B, H, W, M, P);
  else msda_bimodal_kernel<float><<<grid, block, 0, s>>>(reinterpret_cast<const float*>(value), offw, ld_offw, reinterpret_cast<float*>(out), B, H, W, M, P);
  MMT_RETURN_LAST_ERROR();
}
