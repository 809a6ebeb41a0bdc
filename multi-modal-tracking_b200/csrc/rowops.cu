// HBM-bound row kernels of the tracker forward: patchify, LayerNorm, GroupNorm, token row copies,
// fusion query/value staging, 3x3 im2col (with fused nearest upsampling) and the corner
// soft-argmax decode.  All are coalesced + vectorised; activations are "T" = bf16 (fast mode) or
// fp32 (parity mode); statistics and reductions are always fp32.
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "../../include/mmt_b200.h"

namespace mmt {

// ------------------------------------------------------------------------------------------------
// patchify: NCHW fp32 image -> patch matrix rows (Conv2d(3, C, P, P) as a GEMM;
// lib/models/mixformer_vit/mixformer.py:25-33).  k = c*P*P + ky*P + kx matches weight.view(C, -1).
// out row = b*tok_per_seq + tok_off + (py*(W/P) + px).
template <typename T>
__global__ void patchify_kernel(const float* __restrict__ img, T* __restrict__ out, int B, int Cin, int H, int W,
                                int P, int tok_off, int tok_per_seq) {
  const int gw = W / P;
  const int b = blockIdx.x / (H / P);
  const int py = blockIdx.x % (H / P);
  const int K = Cin * P * P;
  const int n = Cin * P * W;  // elements of this patch-row
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int x = i % W;
    const int ky = (i / W) % P;
    const int c = i / (W * P);
    const float v = img[((static_cast<size_t>(b) * Cin + c) * H + (py * P + ky)) * W + x];
    const int px = x / P, kx = x % P;
    const size_t row = static_cast<size_t>(b) * tok_per_seq + tok_off + py * gw + px;
    out[row * K + c * P * P + ky * P + kx] = from_f<T>(v);
  }
}

// bf16 output, P % 8 == 0, W % 8 == 0, 16-byte aligned image rows: a thread converts 8 consecutive pixels of one
// image row (two 16-byte loads) into one 16-byte store - the same values at the same places as patchify_kernel.
__global__ void __launch_bounds__(256)
patchify_bf16x8_kernel(const float* __restrict__ img, bf16* __restrict__ out, int B, int Cin, int H, int W, int P,
                       int tok_off, int tok_per_seq) {
  const int gw = W / P;
  const int b = blockIdx.x / (H / P);
  const int py = blockIdx.x % (H / P);
  const int K = Cin * P * P;
  const int w8 = W / 8;
  const int n = Cin * P * w8;  // 8-pixel units of this patch-row
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int x = (i % w8) * 8;
    const int ky = (i / w8) % P;
    const int c = i / (w8 * P);
    const float4* src = reinterpret_cast<const float4*>(img + ((static_cast<size_t>(b) * Cin + c) * H + (py * P + ky)) * W + x);
    const float4 v0 = __ldcs(src), v1 = __ldcs(src + 1);           // streamed: every pixel is read exactly once
    const int px = x / P, kx = x % P;
    const size_t row = static_cast<size_t>(b) * tok_per_seq + tok_off + py * gw + px;
    uint4 w;
    w.x = pack_bf16x2(v0.x, v0.y); w.y = pack_bf16x2(v0.z, v0.w);
    w.z = pack_bf16x2(v1.x, v1.y); w.w = pack_bf16x2(v1.z, v1.w);
    *reinterpret_cast<uint4*>(out + row * K + c * P * P + ky * P + kx) = w;
  }
}

// P % 4 == 0 (the ConvMAE stem's 4x4 / stride-4 convolution): 4 pixels per thread, one 16-byte load, one 8-byte store
__global__ void __launch_bounds__(256)
patchify_bf16x4_kernel(const float* __restrict__ img, bf16* __restrict__ out, int B, int Cin, int H, int W, int P,
                       int tok_off, int tok_per_seq) {
  const int gw = W / P;
  const int b = blockIdx.x / (H / P);
  const int py = blockIdx.x % (H / P);
  const int K = Cin * P * P;
  const int w4 = W / 4;
  const int n = Cin * P * w4;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int x = (i % w4) * 4;
    const int ky = (i / w4) % P;
    const int c = i / (w4 * P);
    const float4 v = __ldcs(reinterpret_cast<const float4*>(img + ((static_cast<size_t>(b) * Cin + c) * H + (py * P + ky)) * W + x));
    const int px = x / P, kx = x % P;
    const size_t row = static_cast<size_t>(b) * tok_per_seq + tok_off + py * gw + px;
    uint2 w;
    w.x = pack_bf16x2(v.x, v.y); w.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + row * K + c * P * P + ky * P + kx) = w;
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim, one warp per row, two-pass statistics in registers.
// gamma/beta set m = (row / period) & 1 (period <= 0: always set 0) implements the modality-specific
// norms of the shared-backbone variants (norm1_v/_i: lib/models/mixformer_vit_rgbt/mixformer_shared.py:143-157)
// and of the fusion encoder (deformable_encoder_lnspecific.py:143-160).
template <int MAXV>  // C <= MAXV * 128
__global__ void layernorm_kernel(const float* __restrict__ x, int rows, int C, float eps,
                                 const float* __restrict__ g0, const float* __restrict__ b0,
                                 const float* __restrict__ g1, const float* __restrict__ b1, int period,
                                 float* out_f32, bf16* out_bf16) {
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(warp) * C);
  const int nv = C >> 2;  // float4 per row
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      v[i] = xr[idx];
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  const int m = period > 0 ? ((warp / period) & 1) : 0;
  const float4* g = reinterpret_cast<const float4*>(m ? g1 : g0);
  const float4* bb = reinterpret_cast<const float4*>(m ? b1 : b0);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      const float4 gg = __ldg(g + idx), be = __ldg(bb + idx);
      float4 o;
      o.x = (v[i].x - mean) * rstd * gg.x + be.x;
      o.y = (v[i].y - mean) * rstd * gg.y + be.y;
      o.z = (v[i].z - mean) * rstd * gg.z + be.z;
      o.w = (v[i].w - mean) * rstd * gg.w + be.w;
      if (out_f32) reinterpret_cast<float4*>(out_f32 + static_cast<size_t>(warp) * C)[idx] = o;
      if (out_bf16) {
        uint2 p;
        p.x = pack_bf16x2(o.x, o.y);
        p.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(out_bf16 + static_cast<size_t>(warp) * C)[idx] = p;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Row statistics + bf16 cast: the stand-alone producer of the folded LayerNorm (mmt_gemm_bf16_ex).  For rows whose
// producer is not a fused GEMM epilogue (the token embedding, the rows candidate elimination re-gathered, launch shapes
// without the CTA-pair epilogue) this kernel leaves EXACTLY what that epilogue leaves: a bf16 copy of the fp32 row and, per
// 128-column statistics slot, the partial (sum, sum of squares) accumulated in the epilogue's own order - slot s covers the
// 32-column chunks {h, h+2, h+4, h+6} (h = s & 1) of the 256-column tile s >> 1, columns ascending, four at a time
// (gemm_tc.cu, TMA fp32 epilogue).  Equal partial sums mean equal mean / rstd bits downstream, so a sequence's boxes do not
// depend on whether its batch was large enough for the CTA-pair kernel (tests: sub-batch bit-identity).
// One warp per row, one pass over HBM; the row is parked in shared memory (33-word pitch) for the ordered slot sums.
template <int MAXV>  // C <= MAXV * 128
__global__ void rowstats_cast_kernel(const float* __restrict__ x, int rows, int C, bf16* __restrict__ xb, int ld_xb,
                                     float* __restrict__ stats, int slots, int slot_stride) {
  extern __shared__ float rs_smem[];
  ptx::pdl_wait();                  // launched with programmatic stream serialization: the producer of x has completed
  ptx::pdl_launch_dependents();     // the consumer GEMM may start its prologue
  const int wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  float* row = rs_smem + wib * (MAXV * 128 + MAXV * 4);
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(warp) * C);
  const int nv = C >> 2;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      const float4 v = xr[idx];
      uint2 p;
      p.x = pack_bf16x2(v.x, v.y);
      p.y = pack_bf16x2(v.z, v.w);
      reinterpret_cast<uint2*>(xb + static_cast<size_t>(warp) * ld_xb)[idx] = p;
      const int c = idx * 4;
      float* d = row + c + (c >> 5);                 // 33-word pitch per 32-column chunk: the slot lanes hit distinct banks
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
  }
  __syncwarp();
  if (lane < slots) {
    const int tile = lane >> 1, h = lane & 1;
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < 4; ++k) {
      const int c0 = tile * 256 + (h + 2 * k) * 32;
      if (c0 >= C) break;
      const float* d = row + c0 + (c0 >> 5);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y0 = d[4 * j], y1 = d[4 * j + 1], y2 = d[4 * j + 2], y3 = d[4 * j + 3];
        s1 += (y0 + y1) + (y2 + y3);
        s2 = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, s2))));
      }
    }
    reinterpret_cast<float2*>(stats)[static_cast<size_t>(lane) * slot_stride + warp] = make_float2(s1, s2);
  }
}

int launch_rowstats_cast(const float* x, int rows, int C, void* xb, int ld_xb, float* stats, int slots, int slot_stride,
                         cudaStream_t s) {
  // slots = C / 128 (one per 128 columns), the layout the consumer epilogue sums: C a multiple of 128
  if (!(x && xb && stats && rows > 0 && C > 0 && C % 128 == 0 && C <= 2048 && slots == C / 128 && ld_xb >= C &&
        ld_xb % 4 == 0 && slot_stride >= rows))
    return MMT_ERR_BAD_ARG;
  const int wpb = 8;
  const int grid = cdiv(rows, wpb);
  bf16* o = reinterpret_cast<bf16*>(xb);
  auto smem = [&](int maxv) { return static_cast<size_t>(wpb) * (maxv * 128 + maxv * 4) * sizeof(float); };
  cudaError_t le = cudaSuccess;
  if (C <= 512) le = launch_pdl(rowstats_cast_kernel<4>, dim3(grid), dim3(wpb * 32), smem(4), s, x, rows, C, o, ld_xb, stats, slots, slot_stride);
  else if (C <= 1024) le = launch_pdl(rowstats_cast_kernel<8>, dim3(grid), dim3(wpb * 32), smem(8), s, x, rows, C, o, ld_xb, stats, slots, slot_stride);
  else {
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(rowstats_cast_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem(16));
      if (e != cudaSuccess) return (int)e;
      attr = true;
    }
    le = launch_pdl(rowstats_cast_kernel<16>, dim3(grid), dim3(wpb * 32), smem(16), s, x, rows, C, o, ld_xb, stats, slots, slot_stride);
  }
  if (le != cudaSuccess) return (int)le;
  MMT_RETURN_LAST_ERROR();
}

// ------------------------------------------------------------------------------------------------
// GroupNorm(G, C) on NHWC rows [B, HW, C] (nn.GroupNorm after the fusion 1x1 convs,
// lib/models/mixformer_vit_rgbt/fusion_utils.py:252-268).  One CTA per (sample, 8 groups).
// Output row of (b, r) is b*out_seq_rows + out_row_off + r, so a modality's map can be written straight into
// its half of the [B, 2*HW, C] fusion token tensor (torch.cat(src_flatten, 1), deformable_encoder_lnspecific.py:101).
__global__ void groupnorm_kernel(const float* __restrict__ x, int HW, int C, int G, float eps,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, float* out_f32,
                                 bf16* out_bf16, int out_seq_rows, int out_row_off) {
  const int cpg = C / G;               // channels per group (multiple of 4)
  const int lanes = 8 * cpg / 4;       // float4 lanes covering this CTA's 8 groups
  const int slices = blockDim.x / lanes;
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * 8 * cpg;
  const int l = threadIdx.x % lanes, sl = threadIdx.x / lanes;
  const int grp = (l * 4) / cpg;       // 0..7
  __shared__ float part[512];
  __shared__ float stat[2][8];
  const float* xb = x + static_cast<size_t>(b) * HW * C + c0 + l * 4;
  const float cnt = static_cast<float>(HW) * cpg;
  const int lpg = cpg / 4;             // float4 lanes per group
  // fixed-order block reduction (run-to-run deterministic): thread g < 8 sums the partials of group g
  auto group_total = [&](float v) -> float {
    part[threadIdx.x] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 8) {
      for (int s2 = 0; s2 < slices; ++s2)
        for (int j = 0; j < lpg; ++j) t += part[s2 * lanes + threadIdx.x * lpg + j];
    }
    return t;
  };

  float s = 0.f;
  if (sl < slices)
    for (int r = sl; r < HW; r += slices) {
      const float4 v = *reinterpret_cast<const float4*>(xb + static_cast<size_t>(r) * C);
      s += v.x + v.y + v.z + v.w;
    }
  {
    const float t = group_total(s);
    if (threadIdx.x < 8) stat[0][threadIdx.x] = t / cnt;
  }
  __syncthreads();
  const float mean = stat[0][grp];
  float q = 0.f;
  if (sl < slices)
    for (int r = sl; r < HW; r += slices) {
      const float4 v = *reinterpret_cast<const float4*>(xb + static_cast<size_t>(r) * C);
      const float a = v.x - mean, bq = v.y - mean, c = v.z - mean, d = v.w - mean;
      q += a * a + bq * bq + c * c + d * d;
    }
  {
    const float t = group_total(q);
    if (threadIdx.x < 8) stat[1][threadIdx.x] = rsqrtf(t / cnt + eps);
  }
  __syncthreads();
  const float rstd = stat[1][grp];
  const float4 gg = *reinterpret_cast<const float4*>(gamma + c0 + l * 4);
  const float4 be = *reinterpret_cast<const float4*>(beta + c0 + l * 4);
  if (sl < slices)
    for (int r = sl; r < HW; r += slices) {
      const float4 v = *reinterpret_cast<const float4*>(x + (static_cast<size_t>(b) * HW + r) * C + c0 + l * 4);
      const size_t off = (static_cast<size_t>(b) * out_seq_rows + out_row_off + r) * C + c0 + l * 4;
      float4 o;
      o.x = (v.x - mean) * rstd * gg.x + be.x;
      o.y = (v.y - mean) * rstd * gg.y + be.y;
      o.z = (v.z - mean) * rstd * gg.z + be.z;
      o.w = (v.w - mean) * rstd * gg.w + be.w;
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + off) = o;
      if (out_bf16) {
        uint2 p;
        p.x = pack_bf16x2(o.x, o.y);
        p.y = pack_bf16x2(o.z, o.w);
        *reinterpret_cast<uint2*>(out_bf16 + off) = p;
      }
    }
}

// The same with the CTA's [HW x 8 groups] slab held in registers (one HBM pass instead of three, every load of a thread in
// flight at once): HW <= MAXR * slices rows.  Same arithmetic and the same fixed-order reductions as groupnorm_kernel -
// bit-identical results; the three dependent passes made the small launches latency chains (20 us at one sequence,
// profiles/r2_launches.md).
template <int MAXR>
__global__ void __launch_bounds__(512) groupnorm_reg_kernel(const float* __restrict__ x, int HW, int C, int G, float eps,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, float* out_f32,
                                     bf16* out_bf16, int out_seq_rows, int out_row_off) {
  const int cpg = C / G;
  const int lanes = 8 * cpg / 4;
  const int slices = blockDim.x / lanes;
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * 8 * cpg;
  const int l = threadIdx.x % lanes, sl = threadIdx.x / lanes;
  const int grp = (l * 4) / cpg;
  __shared__ float part[512];
  __shared__ float stat[2][8];
  const float* xb = x + static_cast<size_t>(b) * HW * C + c0 + l * 4;
  const float cnt = static_cast<float>(HW) * cpg;
  const int lpg = cpg / 4;
  auto group_total = [&](float v) -> float {
    part[threadIdx.x] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 8) {
      for (int s2 = 0; s2 < slices; ++s2)
        for (int j = 0; j < lpg; ++j) t += part[s2 * lanes + threadIdx.x * lpg + j];
    }
    return t;
  };
  float4 v[MAXR];
#pragma unroll
  for (int i = 0; i < MAXR; ++i) {
    const int r = sl + i * slices;
    v[i] = r < HW ? *reinterpret_cast<const float4*>(xb + static_cast<size_t>(r) * C) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXR; ++i)
    if (sl + i * slices < HW) s += v[i].x + v[i].y + v[i].z + v[i].w;
  {
    const float t = group_total(s);
    if (threadIdx.x < 8) stat[0][threadIdx.x] = t / cnt;
  }
  __syncthreads();
  const float mean = stat[0][grp];
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXR; ++i)
    if (sl + i * slices < HW) {
      const float a = v[i].x - mean, bq = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + bq * bq + c * c + d * d;
    }
  {
    const float t = group_total(q);
    if (threadIdx.x < 8) stat[1][threadIdx.x] = rsqrtf(t / cnt + eps);
  }
  __syncthreads();
  const float rstd = stat[1][grp];
  const float4 gg = *reinterpret_cast<const float4*>(gamma + c0 + l * 4);
  const float4 be = *reinterpret_cast<const float4*>(beta + c0 + l * 4);
#pragma unroll
  for (int i = 0; i < MAXR; ++i) {
    const int r = sl + i * slices;
    if (r < HW) {
      const size_t off = (static_cast<size_t>(b) * out_seq_rows + out_row_off + r) * C + c0 + l * 4;
      float4 o;
      o.x = (v[i].x - mean) * rstd * gg.x + be.x;
      o.y = (v[i].y - mean) * rstd * gg.y + be.y;
      o.z = (v[i].z - mean) * rstd * gg.z + be.z;
      o.w = (v[i].w - mean) * rstd * gg.w + be.w;
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + off) = o;
      if (out_bf16) {
        uint2 p;
        p.x = pack_bf16x2(o.x, o.y);
        p.y = pack_bf16x2(o.z, o.w);
        *reinterpret_cast<uint2*>(out_bf16 + off) = p;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// copy_rows: dst[s*rows_per_seq + r, :] = src[s*seq_stride + row_off + r, :]  (fp32 -> T), e.g. the search
// tokens of every sequence out of the [t, ot, s] token layout (mixformer.py:208-214).
template <typename T>
__global__ void copy_rows_kernel(const float* __restrict__ src, int seq_stride, int row_off, int rows_per_seq,
                                 int nseq, int C, T* __restrict__ dst) {
  const int nv = C >> 2;
  const size_t total = static_cast<size_t>(nseq) * rows_per_seq * nv;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = i % nv;
    const size_t row = i / nv;
    const int s = row / rows_per_seq, r = row % rows_per_seq;
    const float4 v =
        reinterpret_cast<const float4*>(src + (static_cast<size_t>(s) * seq_stride + row_off + r) * C)[cv];
    if (sizeof(T) == 4) {
      reinterpret_cast<float4*>(dst)[i] = v;
    } else {
      uint2 p;
      p.x = pack_bf16x2(v.x, v.y);
      p.y = pack_bf16x2(v.z, v.w);
      reinterpret_cast<uint2*>(dst)[i] = p;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fusion_prep: from the fusion token tensor src [B, 2*L, C] fp32 (first L tokens = RGB, last L = TIR) build
//   out_val [B*2L, C]  = T(src)                                   (value_proj input, no positional term)
//   out_q   [B*L, 2C]  = T(cat_channel(src_v + pos_v, src_i + pos_i))   (query_bimodal)
// (lib/models/mixformer_vit_rgbt/deformable_attention/ops/modules/ms_deform_attn_bimodal.py:97-111;
//  pos = sine embedding + level_embed, deformable_encoder_lnspecific.py:86-100).  pos may be NULL.
template <typename T>
__global__ void fusion_prep_kernel(const float* __restrict__ src, const float* __restrict__ pos, int B, int L, int C,
                                   T* out_val, T* out_q) {
  const int nv = C >> 2;
  const size_t total = static_cast<size_t>(B) * 2 * L * nv;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = i % nv;
    const size_t row = i / nv;            // b*2L + t
    const int t = row % (2 * L);
    const int b = row / (2 * L);
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    if (out_val) {
      if (sizeof(T) == 4) reinterpret_cast<float4*>(out_val)[i] = v;
      else {
        uint2 p; p.x = pack_bf16x2(v.x, v.y); p.y = pack_bf16x2(v.z, v.w);
        reinterpret_cast<uint2*>(out_val)[i] = p;
      }
    }
    if (out_q) {
      float4 q = v;
      if (pos) {
        const float4 pp = reinterpret_cast<const float4*>(pos)[static_cast<size_t>(t) * nv + cv];
        q.x += pp.x; q.y += pp.y; q.z += pp.z; q.w += pp.w;
      }
      const int m = t >= L, p_ = t - m * L;
      const size_t o = (static_cast<size_t>(b) * L + p_) * (2 * nv) + m * nv + cv;
      if (sizeof(T) == 4) reinterpret_cast<float4*>(out_q)[o] = q;
      else {
        uint2 p; p.x = pack_bf16x2(q.x, q.y); p.y = pack_bf16x2(q.z, q.w);
        reinterpret_cast<uint2*>(out_q)[o] = p;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// im2col for Conv2d(k=3, p=1) on NHWC maps, with the nearest-neighbour upsampling of the pyramid head
// folded into the gather: the virtual input at output resolution (H, W) is
//     in(b, y, x, c) = src1[b, y / s1, x / s1, c] (+ src2[b, y / s2, x / s2, c])
// (F.interpolate(scale_factor=2|4) + add, lib/models/mixformer_cvt/head.py:166-178).
// out[(b*H + y)*W + x, (ky*3 + kx)*C + c], zero outside the map.  8 channels per thread.
template <typename T>
__global__ void im2col3x3_kernel(const T* __restrict__ src1, int ld1, int s1, const T* __restrict__ src2, int ld2,
                                 int s2, int B, int H, int W, int C, T* __restrict__ out) {
  constexpr int VE = 16 / sizeof(T);  // elements per 16-byte vector
  const int cvn = C / VE;
  const size_t total = static_cast<size_t>(B) * H * W * 9 * cvn;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = i % cvn;
    size_t r = i / cvn;
    const int tap = r % 9;
    r /= 9;
    const int x = r % W;
    r /= W;
    const int y = r % H;
    const int b = r / H;
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const int H1 = H / s1, W1 = W / s1;
      v = *reinterpret_cast<const uint4*>(src1 + (static_cast<size_t>(b) * H1 * W1 + (yy / s1) * W1 + xx / s1) * ld1 +
                                          cv * VE);
      if (src2) {
        const int H2 = H / s2, W2 = W / s2;
        const uint4 w = *reinterpret_cast<const uint4*>(
            src2 + (static_cast<size_t>(b) * H2 * W2 + (yy / s2) * W2 + xx / s2) * ld2 + cv * VE);
        if (sizeof(T) == 4) {
          float4 a = *reinterpret_cast<float4*>(&v);
          const float4 c = *reinterpret_cast<const float4*>(&w);
          a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
          v = *reinterpret_cast<uint4*>(&a);
        } else {
          __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&v);
          const __nv_bfloat162* c = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // add in fp32 and round once (the reference adds fp32 maps)
            const float2 fa = __bfloat1622float2(a[k]), fc = __bfloat1622float2(c[k]);
            a[k] = __floats2bfloat162_rn(fa.x + fc.x, fa.y + fc.y);
          }
        }
      }
    }
    reinterpret_cast<uint4*>(out)[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// upsample_add: out(b, y, x, :) = src1[b, y / s1, x / s1, :] (+ src2[b, y / s2, x / s2, :]) on NHWC maps, bf16, the
// add in fp32 with one rounding (F.interpolate(scale_factor=2|4) + add, lib/models/mixformer_cvt/head.py:166-178).
// Materialises the input of the implicit-GEMM 3x3 convolutions of the pyramid head once (instead of 9 im2col copies).
// One CTA per output row (b, y): the source rows are fixed per CTA, the per-element index arithmetic is two 32-bit
// operations (the first form decomposed a 64-bit linear index with five divisions per 16-byte element: 33 us per launch for
// 30-60 MB, three times what the traffic costs).
__global__ void upsample_add_kernel(const bf16* __restrict__ src1, int ld1, int s1, const bf16* __restrict__ src2,
                                    int ld2, int s2, int B, int H, int W, int C, bf16* __restrict__ out) {
  const int cvn = C / 8;
  const int y = blockIdx.x % H, b = blockIdx.x / H;
  const int W1 = W / s1, H1 = H / s1;
  const bf16* r1 = src1 + (static_cast<size_t>(b) * H1 * W1 + static_cast<size_t>(y / s1) * W1) * ld1;
  const bf16* r2 = nullptr;
  int W2 = 0;
  if (src2) {
    W2 = W / s2;
    r2 = src2 + (static_cast<size_t>(b) * (H / s2) * W2 + static_cast<size_t>(y / s2) * W2) * ld2;
  }
  uint4* orow = reinterpret_cast<uint4*>(out) + (static_cast<size_t>(b) * H + y) * W * cvn;
  const int n = W * cvn;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {      // (four elements per trip with the loads issued first: no faster)
    const int x = e / cvn, cv = e - x * cvn;
    uint4 v = *reinterpret_cast<const uint4*>(r1 + static_cast<size_t>(x / s1) * ld1 + cv * 8);
    if (r2) {
      const uint4 w = *reinterpret_cast<const uint4*>(r2 + static_cast<size_t>(x / s2) * ld2 + cv * 8);
      __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&v);
      const __nv_bfloat162* c = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 fa = __bfloat1622float2(a[k]), fc = __bfloat1622float2(c[k]);
        a[k] = __floats2bfloat162_rn(fa.x + fc.x, fa.y + fc.y);
      }
    }
    orow[e] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Corner decode: score = conv5_1x1(x4) + up4(a3) + up2(a4); softmax over S*S; soft-argmax expectation
// with coord = stride * index; output (x, y) / img_sz  (lib/models/mixformer_cvt/head.py:181,198-212,
// coordinate tables :138-145).  grid = (B, 2 corners); corner c uses the c-th pointer set.
struct CornerArgs {
  const void* x4[2];   // [B*S*S, ld4x] T, C4 channels used
  const void* a3[2];   // [B*(S/4)^2, lda3] T, 1 channel used
  const void* a4[2];   // [B*(S/2)^2, lda4] T, 1 channel used
  const float* w5[2];  // [C4]
  float b5[2];
  int ld4x, lda3, lda4;
};

template <typename T>
__global__ void corner_decode_kernel(CornerArgs a, int S, int C4, float stride_px, float inv_img, float* score_maps,
                                     float* xyxy) {
  extern __shared__ float logit[];  // S*S
  __shared__ float red[32];
  __shared__ float bc[3];
  const int b = blockIdx.x, cor = blockIdx.y;
  const int n = S * S;
  const T* x4 = reinterpret_cast<const T*>(a.x4[cor]);
  const T* a3 = reinterpret_cast<const T*>(a.a3[cor]);
  const T* a4 = reinterpret_cast<const T*>(a.a4[cor]);
  const float* w5 = a.w5[cor];
  __shared__ float w5s[256];                       // conv5 weights (C4 <= 256, checked by the launcher)
  for (int c = threadIdx.x; c < C4; c += blockDim.x) w5s[c] = __ldg(w5 + c);
  __syncthreads();
  const bool vec8 = (C4 % 8) == 0 && (a.ld4x % 8) == 0 && (reinterpret_cast<uintptr_t>(x4) & 15) == 0;
  const int S2 = S / 2, S4 = S / 4;
  float lmax = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int y = i / S, x = i % S;
    const T* row = x4 + (static_cast<size_t>(b) * n + i) * a.ld4x;
    float acc = 0.f;
    if (sizeof(T) == 2 && vec8) {
      // bf16 rows, 16-byte aligned, C4 a multiple of 8: eight channels per load (the scalar loop's 2-byte loads made this
      // kernel 51 us at one sequence - a latency chain of 480 loads per thread)
      const uint4* r8 = reinterpret_cast<const uint4*>(row);
      for (int c8 = 0; c8 < C4 / 8; ++c8) {
        const uint4 q = r8[c8];
        const uint32_t w32[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          acc = fmaf(__uint_as_float(w32[k] << 16), w5s[c8 * 8 + 2 * k], acc);
          acc = fmaf(__uint_as_float(w32[k] & 0xffff0000u), w5s[c8 * 8 + 2 * k + 1], acc);
        }
      }
    } else {
      for (int c = 0; c < C4; ++c) acc = fmaf(to_f<T>(row[c]), w5s[c], acc);
    }
    acc += a.b5[cor];
    if (a3) acc += to_f<T>(a3[(static_cast<size_t>(b) * S4 * S4 + (y / 4) * S4 + x / 4) * a.lda3]);   // pyramid head only
    if (a4) acc += to_f<T>(a4[(static_cast<size_t>(b) * S2 * S2 + (y / 2) * S2 + x / 2) * a.lda4]);
    logit[i] = acc;
    if (score_maps) score_maps[(static_cast<size_t>(b) * 2 + cor) * n + i] = acc;
    lmax = fmaxf(lmax, acc);
  }
  // block max
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  lmax = warp_max(lmax);
  if (lane == 0) red[wid] = lmax;
  __syncthreads();
  if (wid == 0) {
    float m = lane < nw ? red[lane] : -INFINITY;
    m = warp_max(m);
    if (lane == 0) bc[0] = m;
  }
  __syncthreads();
  const float gmax = bc[0];
  float se = 0.f, sx = 0.f, sy = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float e = expf(logit[i] - gmax);
    se += e;
    sx = fmaf(e, static_cast<float>(i % S) * stride_px, sx);
    sy = fmaf(e, static_cast<float>(i / S) * stride_px, sy);
  }
  se = warp_sum(se); sx = warp_sum(sx); sy = warp_sum(sy);
  __syncthreads();
  if (lane == 0) { red[wid] = se; }
  __syncthreads();
  float t0 = 0.f, t1 = 0.f, t2 = 0.f;
  if (wid == 0) { float v = lane < nw ? red[lane] : 0.f; t0 = warp_sum(v); }
  __syncthreads();
  if (lane == 0) { red[wid] = sx; }
  __syncthreads();
  if (wid == 0) { float v = lane < nw ? red[lane] : 0.f; t1 = warp_sum(v); }
  __syncthreads();
  if (lane == 0) { red[wid] = sy; }
  __syncthreads();
  if (wid == 0) { float v = lane < nw ? red[lane] : 0.f; t2 = warp_sum(v); }
  if (threadIdx.x == 0) {
    xyxy[b * 4 + cor * 2 + 0] = (t1 / t0) * inv_img;
    xyxy[b * 4 + cor * 2 + 1] = (t2 / t0) * inv_img;
  }
}

// box_xyxy_to_cxcywh (lib/utils/box_ops.py:27-31) on [B,4].
__global__ void xyxy_to_cxcywh_kernel(const float* __restrict__ xyxy, float* __restrict__ out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float x0 = xyxy[b * 4], y0 = xyxy[b * 4 + 1], x1 = xyxy[b * 4 + 2], y1 = xyxy[b * 4 + 3];
  out[b * 4 + 0] = (x0 + x1) / 2;
  out[b * 4 + 1] = (y0 + y1) / 2;
  out[b * 4 + 2] = x1 - x0;
  out[b * 4 + 3] = y1 - y0;
}

// out[r, :C] = a[r, :], out[r, C:2C] = b[r, :]  (torch.cat([v, i], dim=1) of two NHWC maps, fusion_utils.py:106), T rows.
template <typename T>
__global__ void concat_cols_kernel(const T* __restrict__ a, const T* __restrict__ b, int rows, int C, T* __restrict__ out) {
  constexpr int VE = 16 / sizeof(T);
  const int nv = C / VE;
  const size_t total = static_cast<size_t>(rows) * 2 * nv;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = i % (2 * nv);
    const size_t r = i / (2 * nv);
    const T* src = cv < nv ? a + r * C + cv * VE : b + r * C + (cv - nv) * VE;
    reinterpret_cast<uint4*>(out)[i] = *reinterpret_cast<const uint4*>(src);
  }
}

// rois[b] = (b, xyxy[b] * scale): the target_roi tensor of the SPM (lib/models/mixformer_cvt/score_decoder.py:37-44:
// normalised box * feature width, batch index = arange(B)).
__global__ void spm_rois_kernel(const float* __restrict__ xyxy, int B, float scale, float* __restrict__ rois) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  rois[b * 5 + 0] = static_cast<float>(b);
#pragma unroll
  for (int k = 0; k < 4; ++k) rois[b * 5 + 1 + k] = xyxy[b * 4 + k] * scale;
}

static inline int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = 148 * 16;
  return static_cast<int>(g < cap ? (g ? g : 1) : cap);
}

}  // namespace mmt

using namespace mmt;

extern "C" int mmt_patchify(const float* img, void* out, int B, int Cin, int H, int W, int P, int tok_off,
                            int tok_per_seq, int out_bf16, void* stream) {
  MMT_CHECK_ARG(img && out && B > 0 && Cin > 0 && P > 0 && H % P == 0 && W % P == 0);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = B * (H / P);
  const bool vec8 = out_bf16 && (P % 8 == 0) && (W % 8 == 0) && ((reinterpret_cast<uintptr_t>(img) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const bool vec4 = out_bf16 && (P % 4 == 0) && (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(img) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
  if (vec8) patchify_bf16x8_kernel<<<grid, 256, 0, s>>>(img, reinterpret_cast<bf16*>(out), B, Cin, H, W, P, tok_off, tok_per_seq);
  else if (vec4) patchify_bf16x4_kernel<<<grid, 256, 0, s>>>(img, reinterpret_cast<bf16*>(out), B, Cin, H, W, P, tok_off, tok_per_seq);
  else if (out_bf16) patchify_kernel<bf16><<<grid, 256, 0, s>>>(img, reinterpret_cast<bf16*>(out), B, Cin, H, W, P, tok_off, tok_per_seq);
  else patchify_kernel<float><<<grid, 256, 0, s>>>(img, reinterpret_cast<float*>(out), B, Cin, H, W, P, tok_off, tok_per_seq);
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_layernorm(const float* x, int rows, int C, float eps, const float* g0, const float* b0,
                             const float* g1, const float* b1, int period, float* out_f32, void* out_bf16,
                             void* stream) {
  MMT_CHECK_ARG(x && g0 && b0 && rows > 0 && C > 0 && C % 4 == 0 && C <= 2048);
  MMT_CHECK_ARG(out_f32 || out_bf16);
  if (!g1) { g1 = g0; b1 = b0; }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int wpb = 8;  // warps per block
  const int grid = cdiv(rows, wpb);
  bf16* ob = reinterpret_cast<bf16*>(out_bf16);
  cudaError_t le;
  if (C <= 512) le = launch_pdl(layernorm_kernel<4>, dim3(grid), dim3(wpb * 32), 0, s, x, rows, C, eps, g0, b0, g1, b1, period, out_f32, ob);
  else if (C <= 1024) le = launch_pdl(layernorm_kernel<8>, dim3(grid), dim3(wpb * 32), 0, s, x, rows, C, eps, g0, b0, g1, b1, period, out_f32, ob);
  else le = launch_pdl(layernorm_kernel<16>, dim3(grid), dim3(wpb * 32), 0, s, x, rows, C, eps, g0, b0, g1, b1, period, out_f32, ob);
  if (le != cudaSuccess) return (int)le;
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_rowstats_cast(const float* x, int rows, int C, void* xb, int ld_xb, float* stats, int slots,
                                int slot_stride, void* stream) {
  return mmt::launch_rowstats_cast(x, rows, C, xb, ld_xb, stats, slots, slot_stride, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mmt_groupnorm(const float* x, int B, int HW, int C, int G, float eps, const float* gamma,
                             const float* beta, float* out_f32, void* out_bf16, int out_seq_rows, int out_row_off,
                             void* stream) {
  MMT_CHECK_ARG(x && gamma && beta && B > 0 && HW > 0 && G > 0 && C % G == 0 && G % 8 == 0);
  const int cpg = C / G;
  MMT_CHECK_ARG(cpg % 4 == 0 && cpg <= 64 && (out_f32 || out_bf16));
  if (out_seq_rows <= 0) { out_seq_rows = HW; out_row_off = 0; }
  MMT_CHECK_ARG(out_row_off >= 0 && out_row_off + HW <= out_seq_rows);
  const int lanes = 8 * cpg / 4;
  int slices = 512 / lanes;
  if (slices > 16) slices = 16;
  dim3 grid(G / 8, B);
  cudaStream_t gs = reinterpret_cast<cudaStream_t>(stream);
  if (HW <= 21 * slices && lanes * slices <= 512)       // the slab fits in registers (84 per thread): one pass over HBM
    groupnorm_reg_kernel<21><<<grid, lanes * slices, 0, gs>>>(
      x, HW, C, G, eps, gamma, beta, out_f32, reinterpret_cast<bf16*>(out_bf16), out_seq_rows, out_row_off);
  else
    groupnorm_kernel<<<grid, lanes * slices, 0, gs>>>(
      x, HW, C, G, eps, gamma, beta, out_f32, reinterpret_cast<bf16*>(out_bf16), out_seq_rows, out_row_off);
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_copy_rows(const float* src, int seq_stride, int row_off, int rows_per_seq, int nseq, int C,
                             void* dst, int out_bf16, void* stream) {
  MMT_CHECK_ARG(src && dst && nseq > 0 && rows_per_seq > 0 && C % 4 == 0);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = static_cast<size_t>(nseq) * rows_per_seq * (C / 4);
  const int grid = grid_for(total, 256);
  if (out_bf16) copy_rows_kernel<bf16><<<grid, 256, 0, s>>>(src, seq_stride, row_off, rows_per_seq, nseq, C, reinterpret_cast<bf16*>(dst));
  else copy_rows_kernel<float><<<grid, 256, 0, s>>>(src, seq_stride, row_off, rows_per_seq, nseq, C, reinterpret_cast<float*>(dst));
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_fusion_prep(const float* src, const float* pos, int B, int L, int C, void* out_val, void* out_q,
                               int out_bf16, void* stream) {
  MMT_CHECK_ARG(src && B > 0 && L > 0 && C % 4 == 0 && (out_val || out_q));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = static_cast<size_t>(B) * 2 * L * (C / 4);
  const int grid = grid_for(total, 256);
  if (out_bf16) fusion_prep_kernel<bf16><<<grid, 256, 0, s>>>(src, pos, B, L, C, reinterpret_cast<bf16*>(out_val), reinterpret_cast<bf16*>(out_q));
  else fusion_prep_kernel<float><<<grid, 256, 0, s>>>(src, pos, B, L, C, reinterpret_cast<float*>(out_val), reinterpret_cast<float*>(out_q));
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_im2col3x3(const void* src1, int ld1, int s1, const void* src2, int ld2, int s2, int B, int H, int W,
                             int C, void* out, int is_bf16, void* stream) {
  MMT_CHECK_ARG(src1 && out && B > 0 && H > 0 && W > 0 && s1 > 0 && H % s1 == 0 && W % s1 == 0);
  MMT_CHECK_ARG(!src2 || (s2 > 0 && H % s2 == 0 && W % s2 == 0));
  const int ve = is_bf16 ? 8 : 4;
  MMT_CHECK_ARG(C % ve == 0 && ld1 % ve == 0 && (!src2 || ld2 % ve == 0));
  MMT_CHECK_ARG((reinterpret_cast<uintptr_t>(src1) & 15) == 0 && (reinterpret_cast<uintptr_t>(src2) & 15) == 0);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = static_cast<size_t>(B) * H * W * 9 * (C / ve);
  const int grid = grid_for(total, 256);
  if (is_bf16)
    im2col3x3_kernel<bf16><<<grid, 256, 0, s>>>(reinterpret_cast<const bf16*>(src1), ld1, s1, reinterpret_cast<const bf16*>(src2), ld2, s2, B, H, W, C, reinterpret_cast<bf16*>(out));
  else
    im2col3x3_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(src1), ld1, s1, reinterpret_cast<const float*>(src2), ld2, s2, B, H, W, C, reinterpret_cast<float*>(out));
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_concat_cols(const void* a, const void* b, int rows, int C, void* out, int is_bf16, void* stream) {
  MMT_CHECK_ARG(a && b && out && rows > 0 && C > 0 && C % (is_bf16 ? 8 : 4) == 0);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t total = static_cast<size_t>(rows) * 2 * (C / (is_bf16 ? 8 : 4));
  if (is_bf16) concat_cols_kernel<bf16><<<grid_for(total, 256), 256, 0, s>>>(reinterpret_cast<const bf16*>(a), reinterpret_cast<const bf16*>(b), rows, C, reinterpret_cast<bf16*>(out));
  else concat_cols_kernel<float><<<grid_for(total, 256), 256, 0, s>>>(reinterpret_cast<const float*>(a), reinterpret_cast<const float*>(b), rows, C, reinterpret_cast<float*>(out));
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_spm_rois(const float* xyxy, int B, float scale, float* rois, void* stream) {
  MMT_CHECK_ARG(xyxy && rois && B > 0);
  spm_rois_kernel<<<cdiv(B, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(xyxy, B, scale, rois);
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_upsample_add(const void* src1, int ld1, int s1, const void* src2, int ld2, int s2, int B, int H, int W,
                                int C, void* out, void* stream) {
  MMT_CHECK_ARG(src1 && out && B > 0 && H > 0 && W > 0 && s1 > 0 && H % s1 == 0 && W % s1 == 0);
  MMT_CHECK_ARG(!src2 || (s2 > 0 && H % s2 == 0 && W % s2 == 0));
  MMT_CHECK_ARG(C % 8 == 0 && ld1 % 8 == 0 && (!src2 || ld2 % 8 == 0));
  MMT_CHECK_ARG((reinterpret_cast<uintptr_t>(src1) & 15) == 0 && (reinterpret_cast<uintptr_t>(src2) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  MMT_CHECK_ARG(static_cast<long long>(B) * H < (1ll << 31));
  upsample_add_kernel<<<B * H, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(src1), ld1, s1, reinterpret_cast<const bf16*>(src2), ld2, s2, B, H, W, C,
      reinterpret_cast<bf16*>(out));
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_corner_decode(const void* x4_tl, const void* x4_br, int ld4x, int C4, const float* w5_tl,
                                 const float* w5_br, float b5_tl, float b5_br, const void* a3_tl, const void* a3_br,
                                 int lda3, const void* a4_tl, const void* a4_br, int lda4, int B, int S,
                                 float stride_px, float img_sz, float* score_maps, float* xyxy, float* cxcywh,
                                 int is_bf16, void* stream) {
  MMT_CHECK_ARG(x4_tl && x4_br && w5_tl && w5_br && xyxy && cxcywh);
  // a3 / a4 (the pyramid head's side maps) come in pairs or not at all (plain Corner_Predictor, head.py:23-94)
  MMT_CHECK_ARG((a3_tl != nullptr) == (a3_br != nullptr) && (a4_tl != nullptr) == (a4_br != nullptr));
  MMT_CHECK_ARG(B > 0 && S > 0 && (S % 4 == 0 || (!a3_tl && !a4_tl)) && S * S * 4 <= 96 * 1024 && C4 > 0 && C4 <= 256 &&
                img_sz > 0);
  CornerArgs a;
  a.x4[0] = x4_tl; a.x4[1] = x4_br; a.a3[0] = a3_tl; a.a3[1] = a3_br; a.a4[0] = a4_tl; a.a4[1] = a4_br;
  a.w5[0] = w5_tl; a.w5[1] = w5_br; a.b5[0] = b5_tl; a.b5[1] = b5_br;
  a.ld4x = ld4x; a.lda3 = lda3; a.lda4 = lda4;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(S) * S * sizeof(float);
  dim3 grid(B, 2);
  if (is_bf16) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(corner_decode_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    corner_decode_kernel<bf16><<<grid, 1024, smem, s>>>(a, S, C4, stride_px, 1.0f / img_sz, score_maps, xyxy);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(corner_decode_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    corner_decode_kernel<float><<<grid, 1024, smem, s>>>(a, S, C4, stride_px, 1.0f / img_sz, score_maps, xyxy);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  xyxy_to_cxcywh_kernel<<<cdiv(B, 128), 128, 0, s>>>(xyxy, cxcywh, B);
  MMT_RETURN_LAST_ERROR();
}
