// fp32 parity-mode GEMM: out = act(A[M,K] * W[N,K]^T + bias) + rowadd + resid, all fp32, FMA
// accumulation in fp32 (no TF32).  Same epilogue contract as the tcgen05 kernel (gemm_tc.cu);
// this is the arithmetic of the reference's fp32 nn.Linear / nn.Conv2d at test time
// ("All reference arithmetic is fp32 at test time").
//
// Classic register-tiled SIMT kernel: 128x128 output tile, BK=16, 256 threads, 8x8 outputs per
// thread, double-buffered smem with K-major -> [k][m] transposition on the way in.
#include "common.cuh"
#include "../../include/mmt_b200.h"

namespace mmt {

constexpr int F_BM = 128, F_BN = 128, F_BK = 16, F_THREADS = 256;

__global__ void __launch_bounds__(F_THREADS)
gemm_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw, int M, int N, int K,
                const float* __restrict__ bias, int act, const float* resid, int ldr,
                const float* __restrict__ rowadd, int rowadd_period, float* out, int ldo) {
  __shared__ float As[2][F_BK][F_BM + 4];
  __shared__ float Bs[2][F_BK][F_BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * F_BM, n0 = blockIdx.x * F_BN;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 8 x 8 (strided by 16 for conflict-free smem reads)

  // global->smem mapping: each thread loads 2 float4 of A and 2 of W per K-slice
  // tile is 128 rows x 16 k = 512 float4; thread t handles float4 #t and #t+256: row = idx/4, kq = idx%4
  float4 ra[2], rb[2];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * F_THREADS;
      const int r = idx >> 2, kq = (idx & 3) * 4;
      const int gm = m0 + r, gn = n0 + r, gk = k0 + kq;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (gm < M) {
        const float* p = A + static_cast<size_t>(gm) * lda + gk;
        if (gk + 3 < K && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) va = *reinterpret_cast<const float4*>(p);
        else {
          if (gk < K) va.x = p[0];
          if (gk + 1 < K) va.y = p[1];
          if (gk + 2 < K) va.z = p[2];
          if (gk + 3 < K) va.w = p[3];
        }
      }
      if (gn < N) {
        const float* p = W + static_cast<size_t>(gn) * ldw + gk;
        if (gk + 3 < K && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) vb = *reinterpret_cast<const float4*>(p);
        else {
          if (gk < K) vb.x = p[0];
          if (gk + 1 < K) vb.y = p[1];
          if (gk + 2 < K) vb.z = p[2];
          if (gk + 3 < K) vb.w = p[3];
        }
      }
      ra[i] = va; rb[i] = vb;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * F_THREADS;
      const int r = idx >> 2, kq = (idx & 3) * 4;
      As[buf][kq + 0][r] = ra[i].x; As[buf][kq + 1][r] = ra[i].y; As[buf][kq + 2][r] = ra[i].z; As[buf][kq + 3][r] = ra[i].w;
      Bs[buf][kq + 0][r] = rb[i].x; Bs[buf][kq + 1][r] = rb[i].y; Bs[buf][kq + 2][r] = rb[i].z; Bs[buf][kq + 3][r] = rb[i].w;
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = (K + F_BK - 1) / F_BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * F_BK);
#pragma unroll
    for (int k = 0; k < F_BK; ++k) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[buf][k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[buf][k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + ty + 16 * i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= N) continue;
      float x = acc[i][j];
      if (bias) x += bias[n];
      if (act == MMT_ACT_GELU) x = gelu_erf(x);
      else if (act == MMT_ACT_RELU) x = fmaxf(x, 0.f);
      if (rowadd) x += rowadd[static_cast<size_t>(row % rowadd_period) * N + n];
      if (resid) x += resid[static_cast<size_t>(row) * ldr + n];
      out[static_cast<size_t>(row) * ldo + n] = x;
    }
  }
}

}  // namespace mmt

extern "C" int mmt_gemm_f32(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const float* bias,
                            int act, const float* resid, int ldr, const float* rowadd, int rowadd_period, float* out,
                            int ldo, void* stream) {
  using namespace mmt;
  MMT_CHECK_ARG(A && W && out && M > 0 && N > 0 && K > 0);
  MMT_CHECK_ARG(lda >= K && ldw >= K && ldo >= N);
  MMT_CHECK_ARG(!rowadd || rowadd_period > 0);
  MMT_CHECK_ARG(!resid || ldr >= N);
  dim3 grid(cdiv(N, F_BN), cdiv(M, F_BM));
  gemm_f32_kernel<<<grid, F_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      A, lda, W, ldw, M, N, K, bias, act, resid, ldr, rowadd, rowadd_period, out, ldo);
  MMT_RETURN_LAST_ERROR();
}
