// Shared device/host helpers for the sm_100a tracker-forward kernels.
// Everything in csrc/ is compiled with -gencode arch=compute_100a,code=sm_100a.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define MMT_OK 0
#define MMT_ERR_BAD_ARG 1000001
#define MMT_ERR_UNSUPPORTED 1000002

// Launch-error check used by every C-ABI entry point: returns the cudaError as int.
#define MMT_RETURN_LAST_ERROR()                          \
  do {                                                   \
    cudaError_t _e = cudaGetLastError();                 \
    return _e == cudaSuccess ? MMT_OK : (int)_e;         \
  } while (0)

#define MMT_CHECK_ARG(cond)                              \
  do {                                                   \
    if (!(cond)) return MMT_ERR_BAD_ARG;                 \
  } while (0)

typedef __nv_bfloat16 bf16;

namespace mmt {

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- activation storage type helpers (T = float in fp32 mode, bf16 otherwise) ----
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Exact-erf GELU (nn.GELU default) for the fp32 path.
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// erf-GELU for the bf16 GEMM epilogue: x * Phi(x) with Phi = 0.5 (1 + tanh(g(x))), g an odd degree-7
// polynomial fitted (least squares on [0,6]) so that tanh(g(x)) = erf(x / sqrt 2) to 1.3e-5 in
// x*Phi(x) - i.e. this approximates the EXACT erf GELU of nn.GELU(), not the "tanh GELU" (4.7e-4).
// One MUFU (tanh.approx, rel. err 2^-11) + 9 FP32 ops, so the tile epilogue stays under the MMA
// time of the tile.  |x| is clamped to 6 where g stops being monotone; tanh(g(6)) == 1 in fp32.
__device__ __forceinline__ float gelu_fast(float x) {
  const float xc = fminf(fmaxf(x, -6.0f), 6.0f);
  const float u = xc * xc;
  float g = fmaf(u, -8.21175444e-06f, -2.60437580e-04f);
  g = fmaf(g, u, 3.67492532e-02f);
  g = fmaf(g, u, 7.97674780e-01f);
  g *= xc;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(g));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace mmt
