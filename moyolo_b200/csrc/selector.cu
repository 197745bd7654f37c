// Encoder-side query selection of MYDecoder (ultralytics/nn/modules/head.py:993-1113), the step that
// runs every frame right before the decoder: enc_output (Linear + LayerNorm) and the class-score head over
// all Lv pyramid positions, top-k(300) by max-class logit, gathers, and the box head + anchors for the
// selected rows only (the reference evaluates the 3-layer box MLP on all Lv rows, head.py:1044, and then keeps
// 300 of them, :1053 — the MLP is row-wise, so evaluating it on the selected rows gives identical values).
#include <float.h>

#include "common.cuh"

namespace moyolo {

int linear_rowln_tcgen05(const void* x, int64_t ldx, const void* w, const RowLnArgs& a, cudaStream_t st);
bool linear_tcgen05_supported(const void* x, int64_t ldx, const void* w, int64_t M, int N, int K);

// ---------------------------------------------------------------------------------------------
// top-k of one score row per CTA: MSB-first radix select of the k-th largest key (4 passes of 8 bits,
// per-warp shared-memory histograms), ordered compaction of the k winners, bitonic sort of the winners.
// Output order = torch.topk(sorted=True): descending score; equal scores by ascending index.
// ---------------------------------------------------------------------------------------------
constexpr int kTopkThreads = 1024;
constexpr int kTopkMaxK = 1024;

__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone: larger float <-> larger key
}

__global__ void __launch_bounds__(kTopkThreads) topk_kernel(const float* __restrict__ scores, int64_t row_stride, int n,
                                                           int k, int32_t* __restrict__ idx_out,
                                                           float* __restrict__ val_out) {
  pdl_trigger();
  pdl_wait();
  __shared__ uint32_t s_hist[32][256];        // per-warp histograms
  __shared__ uint32_t s_tot[256];
  __shared__ unsigned long long s_sel[kTopkMaxK];  // (key << 32) | (~index): sorts descending key, ascending index
  __shared__ uint32_t s_prefix, s_remaining;
  __shared__ int s_warp[33];
  __shared__ int s_count_gt, s_count_eq;
  const float* row = scores + static_cast<int64_t>(blockIdx.x) * row_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { s_prefix = 0u; s_remaining = static_cast<uint32_t>(k); }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = threadIdx.x; i < 32 * 256; i += kTopkThreads) (&s_hist[0][0])[i] = 0u;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    for (int i = threadIdx.x; i < n; i += kTopkThreads) {
      const uint32_t u = float_key(row[i]);
      if (pass == 0 || (u >> (shift + 8)) == prefix) atomicAdd(&s_hist[warp][(u >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 256) {
      uint32_t t = 0;
#pragma unroll 8
      for (int w = 0; w < 32; ++w) t += s_hist[w][threadIdx.x];
      s_tot[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // 256 bins, top digit first
      uint32_t rem = s_remaining, d = 255;
      for (;; --d) {
        const uint32_t c = s_tot[d];
        if (c >= rem || d == 0) break;
        rem -= c;
      }
      s_remaining = rem;
      s_prefix = (prefix << 8) | d;
    }
    __syncthreads();
  }
  const uint32_t thr = s_prefix;            // key of the k-th largest element
  const int take_eq = static_cast<int>(s_remaining);  // how many elements equal to it belong to the top-k
  // ordered compaction: keys > thr all, keys == thr the first take_eq by index
  if (threadIdx.x == 0) { s_count_gt = 0; s_count_eq = 0; }
  for (int i = threadIdx.x; i < kTopkMaxK; i += kTopkThreads) s_sel[i] = 0ull;  // padding sorts last
  __syncthreads();
  for (int base = 0; base < n; base += kTopkThreads) {
    const int i = base + threadIdx.x;
    const uint32_t u = i < n ? float_key(row[i]) : 0u;
    const bool gt = i < n && u > thr, eq = i < n && u == thr;
    // block-wide exclusive ranks of gt and eq in index order
    const unsigned bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
    const int rg = __popc(bg & ((1u << lane) - 1u)), re = __popc(be & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = (__popc(bg) << 16) | __popc(be);
    __syncthreads();
    int og = 0, oe = 0;
    for (int w = 0; w < warp; ++w) { og += s_warp[w] >> 16; oe += s_warp[w] & 0xFFFF; }
    const int cg = s_count_gt, ce = s_count_eq;
    const unsigned long long packed = (static_cast<unsigned long long>(u) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(i));
    if (gt) s_sel[cg + og + rg] = packed;
    __syncthreads();
    if (eq && ce + oe + re < take_eq) s_sel[(k - take_eq) + ce + oe + re] = packed;
    if (threadIdx.x == kTopkThreads - 1) {
      s_count_gt = cg + og + rg + (gt ? 1 : 0);
      s_count_eq = ce + oe + re + (eq ? 1 : 0);
    }
    __syncthreads();
  }
  // bitonic sort (descending) of kTopkMaxK packed entries, one per thread
  for (int size = 2; size <= kTopkMaxK; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int t = threadIdx.x, partner = t ^ stride;
      if (partner > t) {
        const unsigned long long a = s_sel[t], b = s_sel[partner];
        const bool desc = (t & size) == 0;
        if (desc ? a < b : a > b) { s_sel[t] = b; s_sel[partner] = a; }
      }
      __syncthreads();
    }
  }
  for (int j = threadIdx.x; j < k; j += kTopkThreads) {
    const int id = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(s_sel[j] & 0xFFFFFFFFull));
    idx_out[static_cast<int64_t>(blockIdx.x) * k + j] = id;
    if (val_out != nullptr) val_out[static_cast<int64_t>(blockIdx.x) * k + j] = row[id];
  }
}

// one CTA (64 threads) per selected row: det_embed = features[idx] (fp32 + GEMM-operand copy), enc_scores = logits[idx]
__global__ void __launch_bounds__(64) select_gather_kernel(const float* __restrict__ features, const float* __restrict__ logits,
                                                          const int32_t* __restrict__ idx, int k, int64_t len_v, int C, int nc,
                                                          float* __restrict__ embed, void* __restrict__ embed_lp, int lp_bf16,
                                                          float* __restrict__ enc_scores) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x;              // b * k + j
  const int b = r / k;
  const int64_t src = static_cast<int64_t>(b) * len_v + idx[r];
  for (int c = threadIdx.x * 4; c < C; c += 64 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(features + src * C + c);
    *reinterpret_cast<float4*>(embed + static_cast<int64_t>(r) * C + c) = v;
    if (embed_lp != nullptr) {
      if (lp_bf16) {
        uint2 u;
        u.x = float2_to_bf16x2(v.x, v.y);
        u.y = float2_to_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(embed_lp) + static_cast<int64_t>(r) * C + c) = u;
      } else {
        *reinterpret_cast<float4*>(static_cast<float*>(embed_lp) + static_cast<int64_t>(r) * C + c) = v;
      }
    }
  }
  if (enc_scores != nullptr && logits != nullptr && threadIdx.x < nc) enc_scores[static_cast<int64_t>(r) * nc + threadIdx.x] = logits[src * nc + threadIdx.x];
}

// refer[r, :] = h[r, :] . w3^T + b3 + anchor(idx[r])   (head.py:1044 with the anchors of :993-1010)
template <typename HT>
__global__ void __launch_bounds__(128) anchor_box_kernel(const HT* __restrict__ h, int64_t ldh, const float* __restrict__ w3,
                                                        const float* __restrict__ b3, const int32_t* __restrict__ idx,
                                                        LevelTable lv, float grid_size, float eps, float* __restrict__ refer,
                                                        int64_t rows, int K) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float d[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  for (int k = lane; k < K; k += 32) {
    const float hv = to_float<HT>(h[row * ldh + k]);
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = fmaf(hv, w3[j * K + k], d[j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) d[j] = warp_sum(d[j]);
  if (lane < 4) {
    const int p = idx[row];
    int l = 0;
    while (l + 1 < lv.n && p >= lv.start[l + 1]) ++l;
    const int q = p - lv.start[l];
    const int y = q / lv.w[l], x = q - y * lv.w[l];
    // sic: the reference divides (x, y) by (h, w) (valid_WH = [h, w], head.py:1000-1001)
    float a[4];
    a[0] = (static_cast<float>(x) + 0.5f) / static_cast<float>(lv.h[l]);
    a[1] = (static_cast<float>(y) + 0.5f) / static_cast<float>(lv.w[l]);
    a[2] = a[3] = grid_size * exp2f(static_cast<float>(l));
    bool valid = true;
#pragma unroll
    for (int j = 0; j < 4; ++j) valid = valid && (a[j] > eps) && (a[j] < 1.0f - eps);
    const float me = lane == 0 ? a[0] : (lane == 1 ? a[1] : (lane == 2 ? a[2] : a[3]));
    const float anchor = valid ? logf(me / (1.0f - me)) : INFINITY;
    const float t = lane == 0 ? d[0] : (lane == 1 ? d[1] : (lane == 2 ? d[2] : d[3]));
    refer[row * 4 + lane] = t + b3[lane] + anchor;
  }
}

// out[r, :] = zero_rows[r % period] ? 0 : x[r, :]   (valid_mask * feats, head.py:1039; fp32 path)
__global__ void mask_rows_kernel(const float* __restrict__ x, const uint8_t* __restrict__ zero_rows, int64_t period,
                                 float* __restrict__ out, int64_t rows, int C) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = rows * C;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / C;
    out[i] = zero_rows[r % period] ? 0.0f : x[i];
  }
}

// valid[p] = every anchor coordinate of pyramid position p lies in (eps, 1 - eps) (head.py:1006)
__global__ void anchor_invalid_kernel(LevelTable lv, float grid_size, float eps, uint8_t* __restrict__ invalid, int total) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    int l = 0;
    while (l + 1 < lv.n && p >= lv.start[l + 1]) ++l;
    const int q = p - lv.start[l];
    const int y = q / lv.w[l], x = q - y * lv.w[l];
    const float a0 = (static_cast<float>(x) + 0.5f) / static_cast<float>(lv.h[l]);
    const float a1 = (static_cast<float>(y) + 0.5f) / static_cast<float>(lv.w[l]);
    const float a2 = grid_size * exp2f(static_cast<float>(l));
    const bool valid = a0 > eps && a0 < 1.0f - eps && a1 > eps && a1 < 1.0f - eps && a2 > eps && a2 < 1.0f - eps;
    invalid[p] = valid ? 0 : 1;
  }
}

}  // namespace moyolo

using namespace moyolo;

extern "C" int moyolo_enc_output_scores(const void* x, int64_t ldx, const void* w, const float* bias, const float* gamma,
                                        const float* beta, float eps, const uint8_t* zero_in_rows, const float* score_w,
                                        const float* score_b, int nc, int64_t M, float* out_f32, void* out_lp,
                                        float* logits, float* max_logit, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(x && w && gamma && beta, MOYOLO_ERR_BAD_ARG, "enc_output_scores: null pointer");
  MOYOLO_REQUIRE(M >= 0 && nc >= 0 && nc <= 8 && ldx >= 256, MOYOLO_ERR_BAD_SHAPE, "enc_output_scores: bad sizes (nc <= 8)");
  MOYOLO_REQUIRE(nc == 0 || score_w != nullptr, MOYOLO_ERR_BAD_ARG, "enc_output_scores: nc > 0 needs score_w");
  if (M == 0) return MOYOLO_OK;
  MOYOLO_REQUIRE(linear_tcgen05_supported(x, ldx, w, M, 256, 256) && (out_f32 == nullptr || aligned16(out_f32)) &&
                     (out_lp == nullptr || aligned16(out_lp)),
                 MOYOLO_ERR_ALIGNMENT, "enc_output_scores: operands must be 16-byte aligned bf16");
  RowLnArgs a{};
  a.bias = bias; a.gamma = gamma; a.beta = beta; a.eps = eps; a.zero_acc_rows = zero_in_rows;
  a.score_w = score_w; a.score_b = score_b; a.nc = nc; a.M = M;
  a.out_f32 = out_f32; a.out_lp = out_lp; a.logits = logits; a.max_logit = max_logit;
  return linear_rowln_tcgen05(x, ldx, w, a, static_cast<cudaStream_t>(stream));
}

extern "C" int moyolo_topk(const float* scores, int64_t row_stride, int n, int batch, int k, int32_t* idx_out,
                           float* val_out, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(scores && idx_out, MOYOLO_ERR_BAD_ARG, "topk: null pointer");
  MOYOLO_REQUIRE(batch > 0 && n > 0 && k > 0 && k <= n && k <= kTopkMaxK && row_stride >= n, MOYOLO_ERR_BAD_SHAPE,
                 "topk: need 0 < k <= min(n, %d), got n=%d k=%d", kTopkMaxK, n, k);
  launch_k(topk_kernel, dim3(batch), dim3(kTopkThreads), 0, static_cast<cudaStream_t>(stream), scores, row_stride, n, k,
           idx_out, val_out);
  return check_launch("topk_kernel");
}

extern "C" int moyolo_select_gather(const float* features, const float* logits, const int32_t* idx, int batch, int k,
                                    int64_t len_v, int C, int nc, float* embed, void* embed_lp, int lp_dtype,
                                    float* enc_scores, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(features && idx && embed, MOYOLO_ERR_BAD_ARG, "select_gather: null pointer");
  MOYOLO_REQUIRE(batch > 0 && k > 0 && C > 0 && C % 4 == 0 && nc <= 64, MOYOLO_ERR_BAD_SHAPE, "select_gather: bad sizes");
  MOYOLO_REQUIRE(aligned16(features) && aligned16(embed) && (embed_lp == nullptr || aligned16(embed_lp)),
                 MOYOLO_ERR_ALIGNMENT, "select_gather: buffers must be 16-byte aligned");
  launch_k(select_gather_kernel, dim3(batch * k), dim3(64), 0, static_cast<cudaStream_t>(stream), features, logits, idx,
           k, len_v, C, nc, embed, embed_lp, lp_dtype == MOYOLO_BF16 ? 1 : 0, enc_scores);
  return check_launch("select_gather_kernel");
}

extern "C" int moyolo_anchor_box(const void* h, int64_t ldh, int h_dtype, const float* w3, const float* b3,
                                 const int32_t* idx, const int32_t* shapes_hw_host, int n_levels, int64_t len_v,
                                 float grid_size, float eps, float* refer, int64_t rows, int K, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(h && w3 && b3 && idx && refer && shapes_hw_host, MOYOLO_ERR_BAD_ARG, "anchor_box: null pointer");
  MOYOLO_REQUIRE(rows >= 0 && K > 0 && ldh >= K, MOYOLO_ERR_BAD_SHAPE, "anchor_box: bad sizes");
  LevelTable lv;
  int rc = make_levels(shapes_hw_host, n_levels, len_v, &lv);
  if (rc != MOYOLO_OK || rows == 0) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>((rows + 3) / 4);
  if (h_dtype == MOYOLO_BF16)
    launch_k(anchor_box_kernel<__nv_bfloat16>, dim3(blocks), dim3(128), 0, st, static_cast<const __nv_bfloat16*>(h), ldh,
             w3, b3, idx, lv, grid_size, eps, refer, rows, K);
  else if (h_dtype == MOYOLO_F32)
    launch_k(anchor_box_kernel<float>, dim3(blocks), dim3(128), 0, st, static_cast<const float*>(h), ldh, w3, b3, idx, lv,
             grid_size, eps, refer, rows, K);
  else
    return fail(MOYOLO_ERR_UNSUPPORTED, "anchor_box: unsupported h_dtype %d", h_dtype);
  return check_launch("anchor_box_kernel");
}

extern "C" int moyolo_anchor_invalid(const int32_t* shapes_hw_host, int n_levels, int64_t len_v, float grid_size, float eps,
                                     uint8_t* invalid, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(shapes_hw_host && invalid, MOYOLO_ERR_BAD_ARG, "anchor_invalid: null pointer");
  LevelTable lv;
  int rc = make_levels(shapes_hw_host, n_levels, len_v, &lv);
  if (rc != MOYOLO_OK) return rc;
  launch_k(anchor_invalid_kernel, dim3(64), dim3(256), 0, static_cast<cudaStream_t>(stream), lv, grid_size, eps, invalid,
           static_cast<int>(len_v));
  return check_launch("anchor_invalid_kernel");
}

extern "C" int moyolo_mask_rows(const float* x, const uint8_t* zero_rows, int64_t period, float* out, int64_t rows, int C,
                                moyolo_stream_t stream) {
  MOYOLO_REQUIRE(x && zero_rows && out && period > 0 && rows >= 0 && C > 0, MOYOLO_ERR_BAD_ARG, "mask_rows: bad arguments");
  if (rows == 0) return MOYOLO_OK;
  const int64_t n = rows * C;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  launch_k(mask_rows_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), x, zero_rows, period, out, rows,
           C);
  return check_launch("mask_rows_kernel");
}
