/*
 * mmt_b200.h - C ABI of libmmt_b200.so: the sm_100a kernels behind the per-frame network forward of the
 * MixViT RGB / RGB-T trackers (reference: LZ-QWQ/Multi-modal-Tracking).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - caller allocates every output and workspace; no hidden allocation, no hidden synchronisation;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream;
 *   - return value: 0 on success, a cudaError_t value for launch failures, or MMT_ERR_* (>= 1000001)
 *     for rejected arguments.  Nothing is printed.  (The reference's native ops raise through
 *     AT_ASSERTM / only printf launch errors: external/PreciseRoIPooling/pytorch/prroi_pool/src/
 *     prroi_pooling_gpu.c:42, lib/models/mixformer_vit_rgbt/deformable_attention/ops/src/cuda/
 *     ms_deform_im2col_cuda.cuh:948-952; the Python wrapper turns a non-zero status into RuntimeError.)
 *   - "T" buffers are bf16 when `bf16 != 0` and fp32 otherwise (the fp32 mode is the parity mode).
 *   - all token / pixel tensors are row-major [rows, channels] (NHWC for feature maps).
 *
 * Each entry point cites the reference code whose arithmetic it replaces (paths relative to the
 * reference root).
 */
#ifndef MMT_B200_H
#define MMT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMT_ACT_NONE 0
#define MMT_ACT_GELU 1 /* erf GELU (nn.GELU default), timm Mlp: lib/models/mixformer_vit/mixformer.py:123 */
#define MMT_ACT_RELU 2 /* head conv+BN+ReLU lib/models/mixformer_cvt/head.py:7-20; fusion FFN */

/* Version / capability probe (host only). Returns the ABI version; *sm gets the compiled SM (100). */
int mmt_abi_version(int* sm);

/*
 * out[M,N] = act(A[M,K] * W[N,K]^T + bias[N]) + rowadd[row % period][N] + resid[M,N]
 * A, W bf16 row-major (K contiguous, lda/ldw in elements, multiples of 8, 16-byte aligned base);
 * bias / rowadd / resid fp32 or NULL; out bf16 (out_fp32 = 0) or fp32, may alias resid.
 * tcgen05 / TMEM / TMA kernel.  max_ctas <= 0 means "all SMs".
 * Replaces nn.Linear / 1x1 nn.Conv2d / (with mmt_im2col3x3) 3x3 nn.Conv2d+BN on the path:
 *   qkv, proj, fc1, fc2      lib/models/mixformer_vit/mixformer.py:45-47,56,74,123
 *   patch-embed projection   lib/models/mixformer_vit/mixformer.py:25-33,192-203 (rowadd = pos_embed)
 *   fusion projections       lib/models/mixformer_vit_rgbt/deformable_attention/deformable_encoder_lnspecific.py:127-160
 *   head convs (BN folded)   lib/models/mixformer_cvt/head.py:7-20,159-198
 */
int mmt_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, int act,
                  const float* resid, int ldr, const float* rowadd, int rowadd_period, void* out, int ldo,
                  int out_fp32, int max_ctas, void* stream);

/* Same contract, fp32 operands and fp32 FMA accumulation (parity mode; SIMT kernel). */
int mmt_gemm_f32(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const float* bias, int act,
                 const float* resid, int ldr, const float* rowadd, int rowadd_period, float* out, int ldo,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMT_B200_H */
