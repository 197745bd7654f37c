/* ORACLE — TEST INFRASTRUCTURE ONLY. Never linked or loaded by the product package.
 *
 * Plain-C restatement of the multi-scale deformable attention core (a1):
 *   out[b,q,m,c] = sum_l sum_p w[b,q,m,l,p] * bilinear(value[b, start_l + ., m, c], loc[b,q,m,l,p])
 * following the reference's arithmetic, not its code:
 *   - Python form: ultralytics/nn/modules/utils.py:41-78 (grid = 2*loc-1, F.grid_sample bilinear,
 *     padding_mode zeros, align_corners False  =>  pixel = loc*size - 0.5)
 *   - CUDA form:   MOTR/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:33-84 (four bounds-checked
 *     corners) and :285-291 (pixel coordinates, early-out unless -1 < h,w < size).
 * Pinned by tests/golden/kat0_*.npz (the reference's own known-answer case, MOTR/models/ops/test.py:21-60)
 * and tests/golden/core_*.npz (reference Python core run by oracle/make_golden.py).
 * Built by oracle/Makefile into oracle/_build/libmsda_oracle.so.
 */
#include <math.h>
#include <stdint.h>

#define DEFINE_CORE(NAME, T, FLOOR)                                                                   \
  void NAME(const T* value, const int32_t* shapes_hw, int n_levels, int B, int64_t Lv, int H, int D,   \
            const T* loc, const T* w, int Q, int P, T* out) {                                          \
    int64_t start[16];                                                                                \
    int64_t acc = 0;                                                                                  \
    for (int l = 0; l < n_levels; ++l) {                                                              \
      start[l] = acc;                                                                                 \
      acc += (int64_t)shapes_hw[2 * l] * shapes_hw[2 * l + 1];                                        \
    }                                                                                                 \
    for (int b = 0; b < B; ++b) for (int q = 0; q < Q; ++q)                                             \
        for (int m = 0; m < H; ++m) {                                                                 \
      T* o = out + (((int64_t)b * Q + q) * H + m) * D;                                                \
      for (int c = 0; c < D; ++c) o[c] = 0;                                                           \
      for (int l = 0; l < n_levels; ++l) {                                                            \
        const int hh = shapes_hw[2 * l], ww = shapes_hw[2 * l + 1];                                   \
        for (int p = 0; p < P; ++p) {                                                                 \
          const int64_t i = ((((int64_t)b * Q + q) * H + m) * n_levels + l) * P + p;                   \
          const T x = loc[2 * i] * ww - (T)0.5, y = loc[2 * i + 1] * hh - (T)0.5;                      \
          if (!(y > -1 && x > -1 && y < hh && x < ww)) continue;                                      \
          const T x0f = FLOOR(x), y0f = FLOOR(y);                                                     \
          const int x0 = (int)x0f, y0 = (int)y0f;                                                     \
          const T fx = x - x0f, fy = y - y0f, a = w[i];                                               \
          const T cw[4] = {(1 - fy) * (1 - fx), (1 - fy) * fx, fy * (1 - fx), fy * fx};               \
          const int cy[4] = {y0, y0, y0 + 1, y0 + 1}, cx[4] = {x0, x0 + 1, x0, x0 + 1};               \
          for (int k = 0; k < 4; ++k) {                                                               \
            if (cy[k] < 0 || cy[k] >= hh || cx[k] < 0 || cx[k] >= ww) continue;                       \
            const T* v = value + (((int64_t)b * Lv + start[l] + (int64_t)cy[k] * ww + cx[k]) * H + m) * D; \
            const T s = cw[k] * a;                                                                    \
            for (int c = 0; c < D; ++c) o[c] += s * v[c];                                             \
          }                                                                                           \
        }                                                                                             \
      }                                                                                               \
    }                                                                                                 \
  }

DEFINE_CORE(msda_core_f32, float, floorf)
DEFINE_CORE(msda_core_f64, double, floor)

/* Backward of the core, restating the arithmetic of
 * MOTR/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:88-159 (ms_deform_attn_col2im_bilinear: per channel
 * top_grad_value = top_grad * attn_weight; corner k: grad_value += w_k * top_grad_value,
 * grad_h_weight / grad_w_weight accumulate -/+ the opposite fraction times the corner value;
 * grad_attn_weight = top_grad * sampled value; grad_sampling_loc = (width * grad_w_weight, height * grad_h_weight)
 * * top_grad_value) and the channel sums of :301-920. grad_value [B,Lv,H,D] must be zero-filled by the caller.
 * Pinned by tests/golden/core_grad_*.npz (autograd of the reference's multi_scale_deformable_attn_pytorch). */
#define DEFINE_CORE_BWD(NAME, T, FLOOR)                                                               \
  void NAME(const T* value, const int32_t* shapes_hw, int n_levels, int B, int64_t Lv, int H, int D,   \
            const T* loc, const T* w, int Q, int P, const T* grad_out, T* grad_value, T* grad_loc,     \
            T* grad_w) {                                                                               \
    int64_t start[16];                                                                                \
    int64_t acc = 0;                                                                                  \
    for (int l = 0; l < n_levels; ++l) {                                                              \
      start[l] = acc;                                                                                 \
      acc += (int64_t)shapes_hw[2 * l] * shapes_hw[2 * l + 1];                                        \
    }                                                                                                 \
    for (int b = 0; b < B; ++b) for (int q = 0; q < Q; ++q)                                             \
        for (int m = 0; m < H; ++m) {                                                                 \
      const T* go = grad_out + (((int64_t)b * Q + q) * H + m) * D;                                    \
      for (int l = 0; l < n_levels; ++l) {                                                            \
        const int hh = shapes_hw[2 * l], ww = shapes_hw[2 * l + 1];                                   \
        for (int p = 0; p < P; ++p) {                                                                 \
          const int64_t i = ((((int64_t)b * Q + q) * H + m) * n_levels + l) * P + p;                   \
          grad_loc[2 * i] = 0; grad_loc[2 * i + 1] = 0; grad_w[i] = 0;                                \
          const T x = loc[2 * i] * ww - (T)0.5, y = loc[2 * i + 1] * hh - (T)0.5;                      \
          if (!(y > -1 && x > -1 && y < hh && x < ww)) continue;                                      \
          const T x0f = FLOOR(x), y0f = FLOOR(y);                                                     \
          const int x0 = (int)x0f, y0 = (int)y0f;                                                     \
          const T lw = x - x0f, lh = y - y0f, hw = 1 - lw, hhh = 1 - lh, a = w[i];                     \
          const T cw[4] = {hhh * hw, hhh * lw, lh * hw, lh * lw};                                     \
          const T dh[4] = {-hw, -lw, hw, lw}, dw[4] = {-hhh, hhh, -lh, lh};                           \
          const int cy[4] = {y0, y0, y0 + 1, y0 + 1}, cx[4] = {x0, x0 + 1, x0, x0 + 1};               \
          T g_a = 0, g_x = 0, g_y = 0;                                                                \
          for (int c = 0; c < D; ++c) {                                                               \
            const T tgv = go[c] * a;                                                                  \
            T val = 0, ghw = 0, gww = 0;                                                              \
            for (int k = 0; k < 4; ++k) {                                                             \
              if (cy[k] < 0 || cy[k] >= hh || cx[k] < 0 || cx[k] >= ww) continue;                     \
              const int64_t o = (((int64_t)b * Lv + start[l] + (int64_t)cy[k] * ww + cx[k]) * H + m) * D + c; \
              const T v = value[o];                                                                   \
              ghw += dh[k] * v; gww += dw[k] * v; val += cw[k] * v;                                   \
              grad_value[o] += cw[k] * tgv;                                                           \
            }                                                                                         \
            g_a += go[c] * val; g_x += ww * gww * tgv; g_y += hh * ghw * tgv;                         \
          }                                                                                           \
          grad_w[i] = g_a; grad_loc[2 * i] = g_x; grad_loc[2 * i + 1] = g_y;                          \
        }                                                                                             \
      }                                                                                               \
    }                                                                                                 \
  }

DEFINE_CORE_BWD(msda_core_backward_f32, float, floorf)
DEFINE_CORE_BWD(msda_core_backward_f64, double, floor)
