// ORACLE — TEST INFRASTRUCTURE ONLY. Host wrapper that instantiates the REFERENCE's own forward kernel
// (ms_deformable_im2col_cuda<float>, /root/reference/MOTR/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:923-954)
// from the source where it lies; nothing is copied into this repository. Built by oracle/Makefile into
// oracle/_ref/libmsda_ref_cuda.so for sm_100a. Used (a) as a second, independent checker of the
// gather kernel on the GPU and (b) as the "existing CUDA implementation" timed beside ours in
// profiles/ (the reference kernel has no sm_100-specific path).
#include "ms_deform_im2col_cuda.cuh"

extern "C" __attribute__((visibility("default"))) int ref_msda_im2col_f32(
    void* stream, const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
    const float* sampling_loc, const float* attn_weight, int batch, int spatial_size, int num_heads, int channels,
    int num_levels, int num_query, int num_point, float* out) {
  ms_deformable_im2col_cuda<float>(static_cast<cudaStream_t>(stream), value, spatial_shapes, level_start_index,
                                   sampling_loc, attn_weight, batch, spatial_size, num_heads, channels, num_levels,
                                   num_query, num_point, out);
  return static_cast<int>(cudaGetLastError());
}

// The reference's own backward launcher (ms_deformable_col2im_cuda<float>, same header :956-1130) for one
// im2col step covering the whole batch. grad_value / grad_sampling_loc / grad_attn_weight must be zero-filled.
extern "C" __attribute__((visibility("default"))) int ref_msda_col2im_f32(
    void* stream, const float* grad_col, const float* value, const int64_t* spatial_shapes,
    const int64_t* level_start_index, const float* sampling_loc, const float* attn_weight, int batch, int spatial_size,
    int num_heads, int channels, int num_levels, int num_query, int num_point, float* grad_value,
    float* grad_sampling_loc, float* grad_attn_weight) {
  ms_deformable_col2im_cuda<float>(static_cast<cudaStream_t>(stream), grad_col, value, spatial_shapes,
                                   level_start_index, sampling_loc, attn_weight, batch, spatial_size, num_heads,
                                   channels, num_levels, num_query, num_point, grad_value, grad_sampling_loc,
                                   grad_attn_weight);
  return static_cast<int>(cudaGetLastError());
}
