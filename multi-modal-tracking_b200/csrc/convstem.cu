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

// Depthwise 5x5 (CBlock.attn): shared-memory tiled.  One CTA = 8 x 16 output pixels x 64 channels of one image: the
// 12 x 20 input window (halo 2, zeros outside the map) is staged once in smem as fp32-ready 16-byte channel vectors, the
// 25 x 64 filter taps beside it; a thread owns 8 channels x a 1 x 4 pixel strip (32 fp32 accumulators, 800 FMAs) and
// reads each of the 5 x 8 window pixels it needs once.  The first version (one thread per 4 channels of one pixel,
// 25 global gathers each) was instruction-bound at 613 M warp instructions per launch (ncu): 14x the HBM time.
constexpr int DW_TY = 8, DW_TX = 16;
template <typename T> struct DwCh { static constexpr int value = sizeof(T) == 2 ? 64 : 32; };   // 30 KB window either way
constexpr int DW_WY = DW_TY + 4, DW_WX = DW_TX + 4;

template <typename T>
__device__ __forceinline__ void dw_load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void dw_load8<float>(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void dw_load8<bf16>(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
template <typename T>
__device__ __forceinline__ void dw_store8(T* p, const float (&v)[8]);
template <>
__device__ __forceinline__ void dw_store8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void dw_store8<bf16>(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

template <typename T>
__global__ void __launch_bounds__(256)
dwconv5x5_kernel(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                 int B, int H, int W, int E, T* __restrict__ out) {
  constexpr int DW_CH = DwCh<T>::value;
  __shared__ __align__(16) T s_in[DW_WY * DW_WX * DW_CH];          // window, [wy][wx][channel block]
  __shared__ __align__(16) float s_w[25 * DW_CH];
  const int tiles_x = (W + DW_TX - 1) / DW_TX, tiles_y = (H + DW_TY - 1) / DW_TY, ncb = E / DW_CH;
  const int tid = threadIdx.x;
  int t = blockIdx.x;
  const int cb = t % ncb; t /= ncb;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int x0 = tx * DW_TX, y0 = ty * DW_TY, c0 = cb * DW_CH;
  constexpr int VE = 16 / sizeof(T);             // channels per 16-byte vector
  constexpr int VPP = DW_CH / VE;                // vectors per pixel
  for (int i = tid; i < DW_WY * DW_WX * VPP; i += 256) {
    const int v = i % VPP, px = i / VPP;
    const int wx = px % DW_WX, wy = px / DW_WX;
    const int yy = y0 + wy - 2, xx = x0 + wx - 2;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W)
      val = *reinterpret_cast<const uint4*>(in + ((static_cast<size_t>(b) * H + yy) * W + xx) * E + c0 + v * VE);
    reinterpret_cast<uint4*>(s_in)[i] = val;
  }
  for (int i = tid; i < 25 * DW_CH / 4; i += 256) {
    const int tap = i / (DW_CH / 4), q = i % (DW_CH / 4);
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(w + static_cast<size_t>(tap) * E + c0) + q);
  }
  __syncthreads();
  // thread -> channel group (8 channels) and a 1 x 4 pixel strip of the 8 x 16 tile
  constexpr int NCG = DW_CH / 8;
  const int cg = tid % NCG, strip = tid / NCG;   // 32 strips: 8 rows x 4 strips per row
  if (strip >= 32) return;                       // (fp32 mode: 4 channel groups -> 128 computing threads)
  const int ly = strip >> 2, lx = (strip & 3) * 4;
  float acc[4][8];
  {
    float bv[8];
    dw_load8<float>(bias + c0 + cg * 8, bv);
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[o][c] = bv[c];
  }
#pragma unroll
  for (int ky = 0; ky < 5; ++ky) {
    float px[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dw_load8<T>(s_in + ((ly + ky) * DW_WX + lx + j) * DW_CH + cg * 8, px[j]);
#pragma unroll
    for (int kx = 0; kx < 5; ++kx) {
      float wv[8];
      dw_load8<float>(s_w + (ky * 5 + kx) * DW_CH + cg * 8, wv);
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[o][c] = fmaf(wv[c], px[o + kx][c], acc[o][c]);
    }
  }
  const int y = y0 + ly;
  if (y < H) {
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int x = x0 + lx + o;
      if (x < W) dw_store8<T>(out + ((static_cast<size_t>(b) * H + y) * W + x) * E + c0 + cg * 8, acc[o]);
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
  const int chb = is_bf16 ? DwCh<bf16>::value : DwCh<float>::value;
  MMT_CHECK_ARG(in && w && bias && out && B > 0 && H > 0 && W > 0 && E % chb == 0);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long ctas = static_cast<long long>(B) * cdiv(H, DW_TY) * cdiv(W, DW_TX) * (E / chb);
  MMT_CHECK_ARG(ctas > 0 && ctas < (1ll << 31));
  if (is_bf16) dwconv5x5_kernel<bf16><<<static_cast<unsigned>(ctas), 256, 0, s>>>(reinterpret_cast<const bf16*>(in), w, bias, B, H, W, E, reinterpret_cast<bf16*>(out));
  else dwconv5x5_kernel<float><<<static_cast<unsigned>(ctas), 256, 0, s>>>(reinterpret_cast<const float*>(in), w, bias, B, H, W, E, reinterpret_cast<float*>(out));
  MMT_RETURN_LAST_ERROR();
}
