// Backward of the multi-scale deformable attention gather for sm_100a.
//
// Replaces: ms_deformable_col2im_cuda / ms_deform_attn_col2im_bilinear (MOTR/models/ops/src/cuda/
// ms_deform_im2col_cuda.cuh:88-159, 301-920), i.e. MSDA.ms_deform_attn_backward behind
// MSDeformAttnFunction.backward (MOTR/models/ops/functions/ms_deform_attn_func.py:33-41), and the autograd
// of multi_scale_deformable_attn_pytorch (ultralytics/nn/modules/utils.py:41-78). Written from the arithmetic:
//
//   out[b,q,m,:]   = sum_{l,p} A * (w1 v1 + w2 v2 + w3 v3 + w4 v4)         (bilinear corners, zero padding)
//   dA             = <g, w1 v1 + w2 v2 + w3 v3 + w4 v4>
//   dloc_x         = W_l * A * <g, hy (v2 - v1) + ly (v4 - v3)>
//   dloc_y         = H_l * A * <g, hx (v3 - v1) + lx (v4 - v2)>
//   dvalue[corner] += w_k * A * g                                            (scatter-add)
//
// The reference launches one thread per (query, head, channel) and reduces the location / weight gradients over
// the channels through shared memory (seven kernel variants by channel count). Here one WARP owns a
// (query row, head) item, exactly like the forward kernel: the head's value row is covered by G adjacent lanes with
// one 128-bit load each, the warp's 32/G sub-groups each take one sampling point per round (four corner rows in
// flight per lane), the channel reduction is log2(G) shuffles, and the scatter-add is one 128-bit vector
// reduction (RED.ADD.v4.f32) per lane and corner instead of Dh scalar atomics.
#include <algorithm>

#include "common.cuh"

namespace moyolo {

struct MsdaBwdParams {
  const void* value;
  int64_t value_batch_stride;  // elements
  int64_t value_pos_stride;
  LevelTable lv;
  int batch;
  int n_heads;
  int n_points;
  const void* loc;       // [rows, H, L, P, 2]
  const void* weights;   // [rows, H, L, P]
  const void* grad_out;  // [rows, >= H*Dh]
  int64_t grad_out_row_stride;
  int64_t rows;
  int64_t rows_per_batch;
  const int32_t* row_offsets;
  void* grad_value;      // [B, Lv, H, Dh] contiguous, accumulated into
  void* grad_loc;        // [rows, H, L, P, 2]
  void* grad_weights;    // [rows, H, L, P]
};

constexpr int kBwdWarps = 8;
constexpr int kBwdMaxPoints = 64;

struct __align__(8) BwdPoint {
  int pos[4];   // spatial index of the corner or -1 (outside: contributes nothing, as zero padding)
  float lx, ly; // fractions
  float a;      // attention weight
  float wf, hf; // level width / height as float
  int pad;
};

template <typename VT, int DH>
__global__ void __launch_bounds__(kBwdWarps * 32) msda_backward_kernel(const MsdaBwdParams p) {
  constexpr int G = DH * static_cast<int>(sizeof(VT)) / 16;  // lanes per value row
  constexpr int NSG = 32 / G;                                // sub-groups = points per round
  constexpr int CPL = 16 / static_cast<int>(sizeof(VT));     // channels per lane
  __shared__ BwdPoint s_pts[kBwdWarps][kBwdMaxPoints];
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = static_cast<int64_t>(blockIdx.x) * kBwdWarps + warp;
  if (item >= p.rows * p.n_heads) return;
  const int64_t row = item / p.n_heads;
  const int head = static_cast<int>(item % p.n_heads);
  const int b = batch_of_row(row, p.row_offsets, p.batch, p.rows_per_batch);
  const int LP = p.lv.n * p.n_points;
  BwdPoint* pts = s_pts[warp];
  const float* loc = static_cast<const float*>(p.loc);
  const float* wts = static_cast<const float*>(p.weights);
  // ---- phase 1: lane pt stages the geometry of sampling point pt (pixel = loc*size - 0.5, .cuh:285-291 analogue) ----
  for (int pt = lane; pt < LP; pt += 32) {
    const int64_t i = item * LP + pt;
    const float2 l2 = *reinterpret_cast<const float2*>(loc + 2 * i);
    const int level = pt / p.n_points;
    const int H = p.lv.h[level], W = p.lv.w[level], start = p.lv.start[level];
    const float x = l2.x * static_cast<float>(W) - 0.5f;
    const float y = l2.y * static_cast<float>(H) - 0.5f;
    const bool inside = (x > -1.0f) && (y > -1.0f) && (x < static_cast<float>(W)) && (y < static_cast<float>(H));
    const float xf = floorf(x), yf = floorf(y);
    const int x0 = inside ? static_cast<int>(xf) : -2, y0 = inside ? static_cast<int>(yf) : -2;
    const int x1 = x0 + 1, y1 = y0 + 1;
    const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
    BwdPoint q;
    q.pos[0] = (inside && vy0 && vx0) ? start + y0 * W + x0 : -1;
    q.pos[1] = (inside && vy0 && vx1) ? start + y0 * W + x1 : -1;
    q.pos[2] = (inside && vy1 && vx0) ? start + y1 * W + x0 : -1;
    q.pos[3] = (inside && vy1 && vx1) ? start + y1 * W + x1 : -1;
    q.lx = x - xf;
    q.ly = y - yf;
    q.a = wts[i];
    q.wf = static_cast<float>(W);
    q.hf = static_cast<float>(H);
    q.pad = 0;
    pts[pt] = q;
  }
  __syncwarp();
  // ---- phase 2: sub-group `sub` walks points sub, sub + NSG, ...; lane gl of it owns channels [gl*CPL, gl*CPL+CPL) ----
  const int sub = lane / G, gl = lane % G;
  float g[CPL];
  {
    const float* go = static_cast<const float*>(p.grad_out) + row * p.grad_out_row_stride + head * DH + gl * CPL;
#pragma unroll
    for (int c = 0; c < CPL; c += 4) {
      const float4 t = *reinterpret_cast<const float4*>(go + c);
      g[c] = t.x; g[c + 1] = t.y; g[c + 2] = t.z; g[c + 3] = t.w;
    }
  }
  const VT* vbase = static_cast<const VT*>(p.value) + static_cast<int64_t>(b) * p.value_batch_stride + head * DH + gl * CPL;
  float* gvbase = static_cast<float*>(p.grad_value) +
                  (static_cast<int64_t>(b) * (p.lv.start[p.lv.n - 1] + p.lv.h[p.lv.n - 1] * p.lv.w[p.lv.n - 1])) *
                      (static_cast<int64_t>(p.n_heads) * DH) +
                  head * DH + gl * CPL;
  const int64_t gv_pos_stride = static_cast<int64_t>(p.n_heads) * DH;
  const int rounds = (LP + NSG - 1) / NSG;
  for (int it = 0; it < rounds; ++it) {
    const int pt = it * NSG + sub;
    const bool act = pt < LP;
    BwdPoint q;
    if (act) {
      q = pts[pt];
    } else {
      q.pos[0] = q.pos[1] = q.pos[2] = q.pos[3] = -1;
      q.lx = q.ly = q.a = q.wf = q.hf = 0.0f;
    }
    float v[4][CPL];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint4 raw = make_uint4(0u, 0u, 0u, 0u);
      if (q.pos[k] >= 0) raw = ldg128(vbase + static_cast<int64_t>(q.pos[k]) * p.value_pos_stride);
      if constexpr (sizeof(VT) == 2) {
        const float2 a0 = bf16x2_to_float2(raw.x), a1 = bf16x2_to_float2(raw.y), a2 = bf16x2_to_float2(raw.z),
                     a3 = bf16x2_to_float2(raw.w);
        v[k][0] = a0.x; v[k][1] = a0.y; v[k][2] = a1.x; v[k][3] = a1.y;
        v[k][4] = a2.x; v[k][5] = a2.y; v[k][6] = a3.x; v[k][7] = a3.y;
      } else {
        v[k][0] = __uint_as_float(raw.x); v[k][1] = __uint_as_float(raw.y);
        v[k][2] = __uint_as_float(raw.z); v[k][3] = __uint_as_float(raw.w);
      }
    }
    const float hx = 1.0f - q.lx, hy = 1.0f - q.ly;
    const float w1 = hy * hx, w2 = hy * q.lx, w3 = q.ly * hx, w4 = q.ly * q.lx;
    float s_val = 0.0f, s_w = 0.0f, s_h = 0.0f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const float v1 = v[0][c], v2 = v[1][c], v3 = v[2][c], v4 = v[3][c];
      s_val = fmaf(g[c], w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4, s_val);
      s_w = fmaf(g[c], hy * (v2 - v1) + q.ly * (v4 - v3), s_w);
      s_h = fmaf(g[c], hx * (v3 - v1) + q.lx * (v4 - v2), s_h);
    }
    // scatter-add into grad_value: one 128-bit vector reduction per 4 channels and corner
    const float wk[4] = {w1 * q.a, w2 * q.a, w3 * q.a, w4 * q.a};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (q.pos[k] >= 0) {
        float* dst = gvbase + static_cast<int64_t>(q.pos[k]) * gv_pos_stride;
#pragma unroll
        for (int c = 0; c < CPL; c += 4)
          atomicAdd(reinterpret_cast<float4*>(dst + c),
                    make_float4(wk[k] * g[c], wk[k] * g[c + 1], wk[k] * g[c + 2], wk[k] * g[c + 3]));
      }
    }
    // channel reduction across the G lanes of the sub-group
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      s_val += __shfl_xor_sync(0xffffffffu, s_val, o);
      s_w += __shfl_xor_sync(0xffffffffu, s_w, o);
      s_h += __shfl_xor_sync(0xffffffffu, s_h, o);
    }
    if (act && gl == 0) {
      const int64_t i = item * LP + pt;
      static_cast<float*>(p.grad_weights)[i] = s_val;
      *reinterpret_cast<float2*>(static_cast<float*>(p.grad_loc) + 2 * i) =
          make_float2(q.wf * s_w * q.a, q.hf * s_h * q.a);
    }
  }
}

// Generic path: any head_dim, fp32 or fp64 arithmetic (the reference's own gradient test runs gradcheck in
// double, MOTR/models/ops/test.py:63-79). One thread per (row, head, point), scalar atomics.
template <typename VT, typename AT>
__global__ void msda_backward_generic_kernel(const MsdaBwdParams p, int head_dim) {
  pdl_trigger();
  pdl_wait();
  const int LP = p.lv.n * p.n_points;
  const int64_t total = p.rows * p.n_heads * LP;
  const int64_t len_v = p.lv.start[p.lv.n - 1] + static_cast<int64_t>(p.lv.h[p.lv.n - 1]) * p.lv.w[p.lv.n - 1];
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int pt = static_cast<int>(idx % LP);
    const int head = static_cast<int>((idx / LP) % p.n_heads);
    const int64_t row = idx / (static_cast<int64_t>(LP) * p.n_heads);
    const int b = batch_of_row(row, p.row_offsets, p.batch, p.rows_per_batch);
    const int level = pt / p.n_points;
    const int H = p.lv.h[level], W = p.lv.w[level];
    const AT lx = static_cast<const AT*>(p.loc)[idx * 2], ly = static_cast<const AT*>(p.loc)[idx * 2 + 1];
    const AT a = static_cast<const AT*>(p.weights)[idx];
    AT* gl = static_cast<AT*>(p.grad_loc) + idx * 2;
    AT* gw = static_cast<AT*>(p.grad_weights) + idx;
    const AT x = lx * W - AT(0.5), y = ly * H - AT(0.5);
    if (!(x > AT(-1) && y > AT(-1) && x < AT(W) && y < AT(H))) {
      gl[0] = 0; gl[1] = 0; *gw = 0;
      continue;
    }
    const AT xf = floor(x), yf = floor(y);
    const AT fx = x - xf, fy = y - yf, hx = AT(1) - fx, hy = AT(1) - fy;
    const int x0 = static_cast<int>(xf), y0 = static_cast<int>(yf);
    const int ys[4] = {y0, y0, y0 + 1, y0 + 1}, xs[4] = {x0, x0 + 1, x0, x0 + 1};
    const AT wk[4] = {hy * hx, hy * fx, fy * hx, fy * fx};
    const VT* vb = static_cast<const VT*>(p.value) + static_cast<int64_t>(b) * p.value_batch_stride + head * head_dim;
    AT* gvb = static_cast<AT*>(p.grad_value) + (static_cast<int64_t>(b) * len_v) * (static_cast<int64_t>(p.n_heads) * head_dim) +
              head * head_dim;
    const AT* go = static_cast<const AT*>(p.grad_out) + row * p.grad_out_row_stride + head * head_dim;
    AT s_val = 0, s_w = 0, s_h = 0;
    for (int c = 0; c < head_dim; ++c) {
      AT v[4];
      const AT tg = go[c];
      for (int k = 0; k < 4; ++k) {
        const bool ok = ys[k] >= 0 && ys[k] < H && xs[k] >= 0 && xs[k] < W;
        const int64_t pos = p.lv.start[level] + static_cast<int64_t>(ys[k]) * W + xs[k];
        v[k] = ok ? static_cast<AT>(vb[pos * p.value_pos_stride + c]) : AT(0);
        if (ok) atomicAdd(gvb + pos * (static_cast<int64_t>(p.n_heads) * head_dim) + c, wk[k] * a * tg);
      }
      s_val += tg * (wk[0] * v[0] + wk[1] * v[1] + wk[2] * v[2] + wk[3] * v[3]);
      s_w += tg * (hy * (v[1] - v[0]) + fy * (v[3] - v[2]));
      s_h += tg * (hx * (v[2] - v[0]) + fx * (v[3] - v[1]));
    }
    *gw = s_val;
    gl[0] = AT(W) * s_w * a;
    gl[1] = AT(H) * s_h * a;
  }
}

}  // namespace moyolo

using namespace moyolo;

extern "C" int moyolo_msda_sampled_backward(const void* value, int value_dtype, int64_t value_batch_stride,
                                            int64_t value_pos_stride, const int32_t* shapes_hw_host, int n_levels,
                                            int batch, int64_t len_v, int n_heads, int head_dim, int n_points,
                                            const void* loc, const void* weights, int aux_dtype,
                                            const void* grad_out, int64_t grad_out_row_stride, int64_t rows,
                                            const int32_t* row_offsets, void* grad_value, void* grad_loc,
                                            void* grad_weights, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(value && shapes_hw_host && loc && weights && grad_out && grad_value && grad_loc && grad_weights,
                 MOYOLO_ERR_BAD_ARG, "msda_sampled_backward: null pointer");
  MOYOLO_REQUIRE(value_dtype == MOYOLO_F32 || value_dtype == MOYOLO_BF16 || value_dtype == MOYOLO_F64,
                 MOYOLO_ERR_UNSUPPORTED, "unsupported value dtype %d", value_dtype);
  MOYOLO_REQUIRE(batch > 0 && n_heads > 0 && head_dim > 0 && n_points > 0 && rows >= 0, MOYOLO_ERR_BAD_ARG,
                 "batch/n_heads/head_dim/n_points must be positive");
  MOYOLO_REQUIRE(row_offsets != nullptr || rows % batch == 0, MOYOLO_ERR_BAD_SHAPE,
                 "dense rows (%lld) must be divisible by batch (%d)", (long long)rows, batch);
  MOYOLO_REQUIRE(value_pos_stride >= static_cast<int64_t>(n_heads) * head_dim &&
                     grad_out_row_stride >= static_cast<int64_t>(n_heads) * head_dim,
                 MOYOLO_ERR_BAD_SHAPE, "value_pos_stride / grad_out_row_stride smaller than n_heads*head_dim");
  MOYOLO_REQUIRE((value_dtype == MOYOLO_F64) == (aux_dtype == MOYOLO_F64) &&
                     (aux_dtype == MOYOLO_F32 || aux_dtype == MOYOLO_F64),
                 MOYOLO_ERR_UNSUPPORTED, "loc/weights/gradients are fp32 (fp32/bf16 value) or fp64 (fp64 value)");
  MsdaBwdParams p{};
  int rc = make_levels(shapes_hw_host, n_levels, len_v, &p.lv);
  if (rc != MOYOLO_OK) return rc;
  p.value = value; p.value_batch_stride = value_batch_stride; p.value_pos_stride = value_pos_stride;
  p.batch = batch; p.n_heads = n_heads; p.n_points = n_points;
  p.loc = loc; p.weights = weights; p.grad_out = grad_out; p.grad_out_row_stride = grad_out_row_stride;
  p.rows = rows; p.rows_per_batch = rows / batch > 0 ? rows / batch : 1; p.row_offsets = row_offsets;
  p.grad_value = grad_value; p.grad_loc = grad_loc; p.grad_weights = grad_weights;
  if (rows == 0) return MOYOLO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int LP = n_levels * n_points;
  const int esz = value_dtype == MOYOLO_BF16 ? 2 : 4;
  const bool fast_ok = value_dtype != MOYOLO_F64 && (head_dim == 32 || head_dim == 64) && LP <= kBwdMaxPoints &&
                       aligned16(value) && aligned16(grad_out) && aligned16(grad_value) &&
                       (reinterpret_cast<uintptr_t>(loc) & 7u) == 0 && (reinterpret_cast<uintptr_t>(grad_loc) & 7u) == 0 &&
                       (value_pos_stride * esz) % 16 == 0 && (value_batch_stride * esz) % 16 == 0 &&
                       (grad_out_row_stride * 4) % 16 == 0 && rows * n_heads * LP < INT32_MAX;
  if (fast_ok) {
    const int64_t items = rows * n_heads;
    const dim3 grid(static_cast<unsigned>((items + kBwdWarps - 1) / kBwdWarps)), block(kBwdWarps * 32);
    if (value_dtype == MOYOLO_BF16) {
      if (head_dim == 32) launch_k(msda_backward_kernel<__nv_bfloat16, 32>, grid, block, 0, st, p);
      else launch_k(msda_backward_kernel<__nv_bfloat16, 64>, grid, block, 0, st, p);
    } else {
      if (head_dim == 32) launch_k(msda_backward_kernel<float, 32>, grid, block, 0, st, p);
      else launch_k(msda_backward_kernel<float, 64>, grid, block, 0, st, p);
    }
    return check_launch("msda_backward_kernel");
  }
  const int64_t total = rows * n_heads * LP;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((total + 127) / 128, 148 * 16));
  if (value_dtype == MOYOLO_F64)
    launch_k(msda_backward_generic_kernel<double, double>, dim3(blocks), dim3(128), 0, st, p, head_dim);
  else if (value_dtype == MOYOLO_F32)
    launch_k(msda_backward_generic_kernel<float, float>, dim3(blocks), dim3(128), 0, st, p, head_dim);
  else
    launch_k(msda_backward_generic_kernel<__nv_bfloat16, float>, dim3(blocks), dim3(128), 0, st, p, head_dim);
  return check_launch("msda_backward_generic_kernel");
}
