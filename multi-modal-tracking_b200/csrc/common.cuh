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

// Programmatic dependent launch (ptx_sm100.cuh): kernels that call pdl_wait() before their first dependent access are
// launched through this helper, which sets cudaLaunchAttributeProgrammaticStreamSerialization unless switched off
// (mmt_config_pdl(0): A/B measurements, include/mmt_b200.h).
namespace mmt {
extern int g_pdl_enabled;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl_enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace mmt

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

// erf-GELU for the bf16 GEMM epilogue: x * Phi(x) with Phi = 0.5 (1 + tanh(g(x))), g(x) = x (c0 + c1 u + c2 u^2),
// u = min(x^2, 64), fitted (weighted least squares on [0, 3.6]) so that tanh(g(x)) = erf(x / sqrt 2) to 3.0e-5 in
// x * Phi(x) - i.e. this approximates the EXACT erf GELU of nn.GELU(), not the "tanh GELU" (4.7e-4) - and is monotone
// on |x| <= 8 (beyond, u is clamped and tanh saturates: g(8) = 13.6).  One MUFU (tanh.approx, rel. err 2^-11) + 7
// FP32/ALU ops per element: the epilogue's instruction stream costs GEMM rate on a power-capped part (fc1: 1448 TF/s
// without any epilogue math, 1131 with the previous 10-op form - profiles/r1_gemm_bound.md).
__device__ __forceinline__ float gelu_fast(float x) {
  const float u = fminf(x * x, 64.0f);
  float p = fmaf(u, -3.58004386e-04f, 3.70462776e-02f);
  p = fmaf(p, u, 7.97462465e-01f);
  const float g = p * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(g));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// Two gelu_fast at once on the packed fp32 pipe (fmul2 / ffma2: one issue slot for two lanes; the same operations in the same
// order per lane, so the result is bit-identical to gelu_fast).
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
  float2 u = __fmul2_rn(x, x);
  u.x = fminf(u.x, 64.0f);
  u.y = fminf(u.y, 64.0f);
  float2 p = __ffma2_rn(u, make_float2(-3.58004386e-04f, -3.58004386e-04f), make_float2(3.70462776e-02f, 3.70462776e-02f));
  p = __ffma2_rn(p, u, make_float2(7.97462465e-01f, 7.97462465e-01f));
  const float2 g = __fmul2_rn(p, x);
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(g.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(g.y));
  const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  return __ffma2_rn(hx, t, hx);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace mmt
