// Precise RoI Pooling, forward only.
// Replaces the reference's native op _prroi_pooling.prroi_pooling_forward_cuda
// (external/PreciseRoIPooling/pytorch/prroi_pool/src/prroi_pooling_gpu.c:22-44, kernel
// external/PreciseRoIPooling/src/prroi_pooling_gpu_impl.cu:149-212, cell integral :71-106):
//   out[r, c, ph, pw] = (1 / bin_area) * integral over the bin of the bilinearly interpolated feature map
//   (feature treated as 0 outside [0,H) x [0,W)); bin_area == 0 -> 0.
//
// The bilinear surface is separable, so the double integral over a bin is
//     sum_h sum_w wy(h) * wx(w) * feat[h, w]
// where wx(w) is the integral of pixel w's hat function over the bin's x-extent.  Each thread computes one
// output element; threads are laid out with the memory-contiguous dimension fastest so the gathers coalesce
// (channel for NHWC feature maps - the layout the rest of this library uses - or pw for NCHW).
#include "common.cuh"
#include "../../include/mmt_b200.h"

namespace mmt {

// integral over [lo, hi] of the hat function centred at pixel p:  max(0, 1 - |u - p|)
__device__ __forceinline__ float hat_integral(float lo, float hi, int p) {
  float acc = 0.f;
  const float fp = static_cast<float>(p);
  {  // right half: u in [p, p+1], weight 1 - (u - p)
    const float a = fmaxf(lo, fp) - fp, b = fminf(hi, fp + 1.f) - fp;
    if (b > a) acc += (b - 0.5f * b * b) - (a - 0.5f * a * a);
  }
  {  // left half: u in [p-1, p], weight 1 - (p - u); substitute v = p - u
    const float a = fp - fminf(hi, fp), b = fp - fmaxf(lo, fp - 1.f);
    if (b > a) acc += (b - 0.5f * b * b) - (a - 0.5f * a * a);
  }
  return acc;
}

template <bool NHWC>
__global__ void prroi_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ rois, float* __restrict__ out,
                                 int R, int C, int H, int W, int PH, int PW, float scale) {
  const size_t total = static_cast<size_t>(R) * C * PH * PW;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int c, pw, ph, r;
    if (NHWC) {  // c fastest
      c = i % C; pw = (i / C) % PW; ph = (i / C / PW) % PH; r = i / C / PW / PH;
    } else {     // pw fastest (== the reference's output linearisation)
      pw = i % PW; ph = (i / PW) % PH; c = (i / PW / PH) % C; r = i / PW / PH / C;
    }
    const float* roi = rois + r * 5;
    const int bi = static_cast<int>(roi[0]);
    const float x0 = roi[1] * scale, y0 = roi[2] * scale, x1 = roi[3] * scale, y1 = roi[4] * scale;
    const float bw = fmaxf(x1 - x0, 0.f) / PW, bh = fmaxf(y1 - y0, 0.f) / PH;
    const float ws = x0 + bw * pw, hs = y0 + bh * ph, we = ws + bw, he = hs + bh;
    const float area = fmaxf(0.f, bw * bh);
    float sum = 0.f;
    if (area > 0.f) {
      const int w_lo = max(0, static_cast<int>(floorf(ws))), w_hi = min(W - 1, static_cast<int>(ceilf(we)));
      const int h_lo = max(0, static_cast<int>(floorf(hs))), h_hi = min(H - 1, static_cast<int>(ceilf(he)));
      for (int h = h_lo; h <= h_hi; ++h) {
        const float wy = hat_integral(hs, he, h);
        if (wy == 0.f) continue;
        float rowacc = 0.f;
        for (int w = w_lo; w <= w_hi; ++w) {
          const float wx = hat_integral(ws, we, w);
          const float v = NHWC ? feat[((static_cast<size_t>(bi) * H + h) * W + w) * C + c]
                               : feat[((static_cast<size_t>(bi) * C + c) * H + h) * W + w];
          rowacc = fmaf(wx, v, rowacc);
        }
        sum = fmaf(wy, rowacc, sum);
      }
      sum /= area;
    }
    // output is always [R, C, PH, PW] (reference layout) for NCHW, [R, PH*PW, C] (token layout) for NHWC
    if (NHWC) out[((static_cast<size_t>(r) * PH + ph) * PW + pw) * C + c] = sum;
    else out[i] = sum;
  }
}

}  // namespace mmt

using namespace mmt;

extern "C" int mmt_prroi_fwd(const float* feat, const float* rois, float* out, int R, int C, int H, int W, int PH,
                             int PW, float spatial_scale, int channels_last, void* stream) {
  MMT_CHECK_ARG(feat && rois && out && R > 0 && C > 0 && H > 0 && W > 0 && PH > 0 && PW > 0);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = static_cast<size_t>(R) * C * PH * PW;
  size_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (channels_last) prroi_fwd_kernel<true><<<(int)g, 256, 0, s>>>(feat, rois, out, R, C, H, W, PH, PW, spatial_scale);
  else prroi_fwd_kernel<false><<<(int)g, 256, 0, s>>>(feat, rois, out, R, C, H, W, PH, PW, spatial_scale);
  MMT_RETURN_LAST_ERROR();
}
