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
