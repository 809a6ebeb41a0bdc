// TEST INFRASTRUCTURE - C entry point around the reference's OWN deformable-attention forward kernel, compiled from
// the header where it lies under /root/reference (never copied): oracle/build_ref.py passes
//   -I <reference>/lib/models/mixformer_vit_rgbt/deformable_attention/ops/src/cuda
// The reference's host wrapper (ms_deform_attn_cuda.cu:20-80) needs two edits to build against torch 2.11
// (`value.type()` -> `value.scalar_type()`) and is bypassed: this file calls the kernel launcher
// `ms_deformable_im2col_cuda<float>` (ms_deform_im2col_cuda.cuh:924-1010) exactly as that wrapper does, one im2col_step
// batch slice at a time.
#include "ms_deform_im2col_cuda.cuh"

extern "C" int msda_ref_forward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                                const float* sampling_loc, const float* attn_weight, float* out, int batch, int spatial_size,
                                int num_heads, int channels, int num_levels, int num_query, int num_point, int im2col_step,
                                void* stream) {
  if (im2col_step <= 0 || batch % im2col_step != 0) return -1;        // ms_deform_attn_cuda.cu:50-52
  const long long per_value = (long long)spatial_size * num_heads * channels;
  const long long per_loc = (long long)num_query * num_heads * num_levels * num_point * 2;
  const long long per_w = (long long)num_query * num_heads * num_levels * num_point;
  const long long per_out = (long long)num_query * num_heads * channels;
  for (int n = 0; n < batch / im2col_step; ++n) {
    ms_deformable_im2col_cuda<float>(static_cast<cudaStream_t>(stream), value + n * im2col_step * per_value, spatial_shapes,
                                     level_start_index, sampling_loc + n * im2col_step * per_loc,
                                     attn_weight + n * im2col_step * per_w, im2col_step, spatial_size, num_heads, channels,
                                     num_levels, num_query, num_point, out + n * im2col_step * per_out);
  }
  return (int)cudaGetLastError();
}
