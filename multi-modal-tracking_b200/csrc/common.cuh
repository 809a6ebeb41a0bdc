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

// GELU with the Abramowitz-Stegun 7.1.26 erf (|err| <= 1.5e-7), written so that the
// negative tail has no 1-erf cancellation. ~14 instructions; used in the bf16 GEMM epilogue
// where the tile epilogue must stay under the MMA time of the tile.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = p * t * __expf(-z * z);  // = 1 - erf(z), in (0,1]
  return x >= 0.f ? x * (1.0f - 0.5f * e) : x * (0.5f * e);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace mmt
