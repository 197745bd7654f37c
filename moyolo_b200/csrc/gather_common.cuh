// Device helpers of the deformable gather shared by msda.cu and decoder_cluster.cu: bilinear corner staging
// (pixel = loc*size - 0.5, zero padding), predicated 128-bit value loads and the packed fp32x2 accumulate.
#pragma once
#include "common.cuh"

namespace moyolo {

// Bilinear corner staging shared by both kernels: pixel = loc*size - 0.5, zero padding
// (grid_sample align_corners=False / ms_deform_im2col_cuda.cuh:285-291 give the same numbers).
struct Corners {
  int pos[4];
  float w[4];
};
__device__ __forceinline__ Corners make_corners(float loc_x, float loc_y, int H, int W, int start,
                                                float aw) {
  Corners c;
  const float x = loc_x * static_cast<float>(W) - 0.5f;
  const float y = loc_y * static_cast<float>(H) - 0.5f;
  const float xf = floorf(x), yf = floorf(y);
  const float lx = x - xf, ly = y - yf;
  const float hx = 1.0f - lx, hy = 1.0f - ly;
  // Guard the float->int conversion: anything at or beyond one pixel outside contributes nothing.
  const bool inside = (x > -1.0f) && (y > -1.0f) && (x < static_cast<float>(W)) &&
                      (y < static_cast<float>(H));
  const int x0 = inside ? static_cast<int>(xf) : -2;
  const int y0 = inside ? static_cast<int>(yf) : -2;
  const int x1 = x0 + 1, y1 = y0 + 1;
  const bool vx0 = (x0 >= 0) && (x0 < W), vx1 = (x1 >= 0) && (x1 < W);
  const bool vy0 = (y0 >= 0) && (y0 < H), vy1 = (y1 >= 0) && (y1 < H);
  c.pos[0] = (inside && vy0 && vx0) ? start + y0 * W + x0 : -1;
  c.pos[1] = (inside && vy0 && vx1) ? start + y0 * W + x1 : -1;
  c.pos[2] = (inside && vy1 && vx0) ? start + y1 * W + x0 : -1;
  c.pos[3] = (inside && vy1 && vx1) ? start + y1 * W + x1 : -1;
  c.w[0] = hy * hx * aw;
  c.w[1] = hy * lx * aw;
  c.w[2] = ly * hx * aw;
  c.w[3] = ly * lx * aw;
  return c;
}

// acc[0..7] += w * (8 bf16 channels of v): four packed fp32x2 FMAs (sm_100 FFMA2; bit-identical to eight
// scalar round-to-nearest FMAs).
__device__ __forceinline__ void fma_bf16x8(float (&a)[8], const uint4& v, float w) {
  const float2 w2 = make_float2(w, w);
  float2 t;
  t = __ffma2_rn(w2, bf16x2_to_float2(v.x), make_float2(a[0], a[1])); a[0] = t.x; a[1] = t.y;
  t = __ffma2_rn(w2, bf16x2_to_float2(v.y), make_float2(a[2], a[3])); a[2] = t.x; a[3] = t.y;
  t = __ffma2_rn(w2, bf16x2_to_float2(v.z), make_float2(a[4], a[5])); a[4] = t.x; a[5] = t.y;
  t = __ffma2_rn(w2, bf16x2_to_float2(v.w), make_float2(a[6], a[7])); a[6] = t.x; a[7] = t.y;
}
__device__ __forceinline__ void fma_f32x4(float (&a)[4], const uint4& v, float w) {
  const float2 w2 = make_float2(w, w);
  float2 t;
  t = __ffma2_rn(w2, make_float2(__uint_as_float(v.x), __uint_as_float(v.y)), make_float2(a[0], a[1])); a[0] = t.x; a[1] = t.y;
  t = __ffma2_rn(w2, make_float2(__uint_as_float(v.z), __uint_as_float(v.w)), make_float2(a[2], a[3])); a[2] = t.x; a[3] = t.y;
}

// 128-bit read-only load that is skipped (result = zeros) when `valid` is false, WITHOUT a branch:
// grid_sample's zero padding for corners outside the map (the address is then never dereferenced).
__device__ __forceinline__ uint4 ldg128_if(const void* ptr, bool valid) {
  uint4 v;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b32 %0, 0;\n\t"
      "mov.b32 %1, 0;\n\t"
      "mov.b32 %2, 0;\n\t"
      "mov.b32 %3, 0;\n\t"
      "@p ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];\n\t"
      "}"
      : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
      : "l"(ptr), "r"(static_cast<int>(valid)));
  return v;
}

}  // namespace moyolo
