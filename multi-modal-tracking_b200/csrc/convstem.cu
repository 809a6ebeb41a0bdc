// HBM-bound kernels of the ConvMAE stem (lib/models/mixformer_convmae/mixformer_online.py:17-51, 167-189):
//   * layernorm_act : channel LayerNorm of NHWC rows followed by an optional exact GELU (PatchEmbed: proj -> norm ->
//                     act, :45-50), with the output rows optionally re-mapped into the [template | online template |
//                     search] token order so that patch_embed4 + pos-embed is ONE GEMM over all tokens;
//   * patchify2x2   : NHWC rows -> patch matrix of a Conv2d(C, E, 2, stride 2), k = (ky*2 + kx)*C + c;
//   * dwconv5x5     : depthwise Conv2d(E, E, 5, padding 2, groups E) + bias on NHWC rows (CBlock.attn, :172).
// All 1x1 convolutions of the stem are plain GEMMs on the NHWC rows (gemm_tc.cu / gemm_f32.cu).
#include "common.cuh"
#include "../../include/mmt_b200.h"

namespace mmt {

template <int MAXV, bool GELU>
__global__ void layernorm_act_kernel(const float* __restrict__ x, int rows, int C, float eps,
                                     const float* __restrict__ g, const float* __restrict__ b, float* out_f32,
                                     bf16* out_bf16, int seg_rows, int out_seq_rows, int out_row_off) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(warp) * C);
  const int nv = C >> 2;
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) { v[i] = xr[idx]; s += v[i].x + v[i].y + v[i].z + v[i].w; }
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + bb * bb + c * c + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  const size_t orow = seg_rows > 0 ? static_cast<size_t>(warp / seg_rows) * out_seq_rows + out_row_off + warp % seg_rows
                                   : static_cast<size_t>(warp);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + idx), be = __ldg(reinterpret_cast<const float4*>(b) + idx);
      float4 o;
      o.x = (v[i].x - mean) * rstd * gg.x + be.x;
      o.y = (v[i].y - mean) * rstd * gg.y + be.y;
      o.z = (v[i].z - mean) * rstd * gg.z + be.z;
      o.w = (v[i].w - mean) * rstd * gg.w + be.w;
      if (GELU) { o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w); }
      if (out_f32) reinterpret_cast<float4*>(out_f32 + orow * C)[idx] = o;
      if (out_bf16) {
        uint2 p;
        p.x = pack_bf16x2(o.x, o.y);
        p.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(out_bf16 + orow * C)[idx] = p;
      }
    }
  }
}

template <typename T>
__global__ void patchify2x2_kernel(const float* __restrict__ x, int B, int H, int W, int C, T* __restrict__ out) {
  const int nv = C >> 2;
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * 4 * nv;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = i % nv;
    size_t r = i / nv;
    const int tap = r % 4;
    r /= 4;                                   // output row (b, py, px)
    const int px = r % Wo;
    const int py = (r / Wo) % Ho;
    const int b = r / (static_cast<size_t>(Wo) * Ho);
    const int y = 2 * py + tap / 2, xx = 2 * px + tap % 2;
    const float4 v = reinterpret_cast<const float4*>(x + ((static_cast<size_t>(b) * H + y) * W + xx) * C)[cv];
    if (sizeof(T) == 4) reinterpret_cast<float4*>(out)[i] = v;
    else {
      uint2 p; p.x = pack_bf16x2(v.x, v.y); p.y = pack_bf16x2(v.z, v.w);
      reinterpret_cast<uint2*>(out)[i] = p;
    }
  }
}

// one thread = 4 channels of one output pixel; weights [25, E] fp32 (tap-major), fp32 accumulation
template <typename T>
__global__ void dwconv5x5_kernel(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                 int B, int H, int W, int E, T* __restrict__ out) {
  const int nv = E >> 2;
  const size_t total = static_cast<size_t>(B) * H * W * nv;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = i % nv;
    size_t r = i / nv;
    const int x = r % W;
    const int y = (r / W) % H;
    const int b = r / (static_cast<size_t>(W) * H);
    float4 acc = __ldg(reinterpret_cast<const float4*>(bias) + cv);
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
      const int yy = y + ky - 2;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        const int xx = x + kx - 2;
        if (xx < 0 || xx >= W) continue;
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w + static_cast<size_t>(ky * 5 + kx) * E) + cv);
        const T* p = in + ((static_cast<size_t>(b) * H + yy) * W + xx) * E + cv * 4;
        float4 v;
        if (sizeof(T) == 4) v = *reinterpret_cast<const float4*>(p);
        else {
          const uint2 u = *reinterpret_cast<const uint2*>(p);
          const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
          const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
          v = make_float4(a.x, a.y, c.x, c.y);
        }
        acc.x = fmaf(ww.x, v.x, acc.x); acc.y = fmaf(ww.y, v.y, acc.y);
        acc.z = fmaf(ww.z, v.z, acc.z); acc.w = fmaf(ww.w, v.w, acc.w);
      }
    }
    if (sizeof(T) == 4) reinterpret_cast<float4*>(out)[i] = acc;
    else {
      uint2 p; p.x = pack_bf16x2(acc.x, acc.y); p.y = pack_bf16x2(acc.z, acc.w);
      reinterpret_cast<uint2*>(out)[i] = p;
    }
  }
}

static inline int stem_grid(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = 148 * 16;
  return static_cast<int>(g < cap ? (g ? g : 1) : cap);
}

}  // namespace mmt

using namespace mmt;

extern "C" int mmt_layernorm_act(const float* x, int rows, int C, float eps, const float* gamma, const float* beta,
                                 int gelu, float* out_f32, void* out_bf16, int seg_rows, int out_seq_rows,
                                 int out_row_off, void* stream) {
  MMT_CHECK_ARG(x && gamma && beta && rows > 0 && C > 0 && C % 4 == 0 && C <= 2048 && (out_f32 || out_bf16));
  // row r -> (r / seg_rows) * out_seq_rows + out_row_off + r % seg_rows; out_seq_rows is a stride (>= seg_rows), the
  // offset may exceed it (online templates: rows T + b*T + i)
  MMT_CHECK_ARG(seg_rows <= 0 || (out_seq_rows >= seg_rows && out_row_off >= 0 && rows % seg_rows == 0));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int wpb = 8, grid = cdiv(rows, wpb);
  bf16* ob = reinterpret_cast<bf16*>(out_bf16);
#define LAUNCH(MAXV)                                                                                              \
  do {                                                                                                            \
    if (gelu) layernorm_act_kernel<MAXV, true><<<grid, wpb * 32, 0, s>>>(x, rows, C, eps, gamma, beta, out_f32, ob, \
                                                                        seg_rows, out_seq_rows, out_row_off);     \
    else layernorm_act_kernel<MAXV, false><<<grid, wpb * 32, 0, s>>>(x, rows, C, eps, gamma, beta, out_f32, ob,    \
                                                                     seg_rows, out_seq_rows, out_row_off);        \
  } while (0)
  if (C <= 512) LAUNCH(4);
  else if (C <= 1024) LAUNCH(8);
  else LAUNCH(16);
#undef LAUNCH
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_patchify2x2(const float* x, int B, int H, int W, int C, void* out, int out_bf16, void* stream) {
  MMT_CHECK_ARG(x && out && B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C % 4 == 0);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = static_cast<size_t>(B) * (H / 2) * (W / 2) * 4 * (C / 4);
  if (out_bf16) patchify2x2_kernel<bf16><<<stem_grid(total, 256), 256, 0, s>>>(x, B, H, W, C, reinterpret_cast<bf16*>(out));
  else patchify2x2_kernel<float><<<stem_grid(total, 256), 256, 0, s>>>(x, B, H, W, C, reinterpret_cast<float*>(out));
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_dwconv5x5(const void* in, const float* w, const float* bias, int B, int H, int W, int E, void* out,
                             int is_bf16, void* stream) {
  MMT_CHECK_ARG(in && w && bias && out && B > 0 && H > 0 && W > 0 && E % 4 == 0);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = static_cast<size_t>(B) * H * W * (E / 4);
  if (is_bf16) dwconv5x5_kernel<bf16><<<stem_grid(total, 256), 256, 0, s>>>(reinterpret_cast<const bf16*>(in), w, bias, B, H, W, E, reinterpret_cast<bf16*>(out));
  else dwconv5x5_kernel<float><<<stem_grid(total, 256), 256, 0, s>>>(reinterpret_cast<const float*>(in), w, bias, B, H, W, E, reinterpret_cast<float*>(out));
  MMT_RETURN_LAST_ERROR();
}
