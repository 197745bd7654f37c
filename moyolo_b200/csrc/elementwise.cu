// Row-wise fused epilogues of the decoder layer: residual + LayerNorm (+ with_pos_embed operand),
// box refinement, score head, positional embedding. One warp per row, fp32 math throughout.
// Reference arithmetic: ultralytics/nn/modules/transformer.py:183-190 (pos2posemb), :637-647 and
// :576-580 (post-norm residual blocks), :709 (box refinement), :717-721 (score head);
// ultralytics/nn/modules/utils.py:34-38 (inverse_sigmoid); ultralytics/nn/modules/head.py:310.
#include "common.cuh"

namespace moyolo {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float inverse_sigmoidf_(float x) {
  x = fminf(fmaxf(x, 0.0f), 1.0f);
  const float x1 = fmaxf(x, 1e-5f);
  const float x2 = fmaxf(1.0f - x, 1e-5f);
  return logf(x1 / x2);
}

constexpr int kRowWarps = 4;  // warps (rows) per block

template <typename LP_T>
__global__ void __launch_bounds__(kRowWarps * 32) add_layernorm_kernel(
    const float* __restrict__ x, const float* __restrict__ residual, const float* __restrict__ gamma,
    const float* __restrict__ beta, float eps, int64_t rows, int C, float* __restrict__ out_f32,
    LP_T* __restrict__ out_lp, const float* __restrict__ pos, LP_T* __restrict__ out_pos_lp) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * C;
  const float* rr = residual ? residual + row * C : nullptr;
  // C <= 1024 is held in registers (32 per lane); larger rows re-read from L1/L2.
  constexpr int kMaxPerLane = 32;
  float v[kMaxPerLane];
  const int per_lane = (C + 31) / 32;
  const bool in_regs = per_lane <= kMaxPerLane;
  float sum = 0.0f;
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int c = lane + i * 32;
      v[i] = (i < per_lane && c < C) ? xr[c] + (rr ? rr[c] : 0.0f) : 0.0f;
      sum += v[i];
    }
  } else {
    for (int c = lane; c < C; c += 32) sum += xr[c] + (rr ? rr[c] : 0.0f);
  }
  const float mean = warp_sum(sum) / static_cast<float>(C);
  float sq = 0.0f;
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int c = lane + i * 32;
      const float d = (i < per_lane && c < C) ? v[i] - mean : 0.0f;
      sq += d * d;
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      const float d = xr[c] + (rr ? rr[c] : 0.0f) - mean;
      sq += d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(C) + eps);
  auto emit = [&](int c, float val) {
    const float o = (val - mean) * rstd * gamma[c] + beta[c];
    if (out_f32) out_f32[row * C + c] = o;
    if (out_lp) out_lp[row * C + c] = from_float<LP_T>(o);
    if (out_pos_lp) out_pos_lp[row * C + c] = from_float<LP_T>(o + pos[row * C + c]);
  };
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int c = lane + i * 32;
      if (i < per_lane && c < C) emit(c, v[i]);
    }
  } else {
    for (int c = lane; c < C; c += 32) emit(c, xr[c] + (rr ? rr[c] : 0.0f));
  }
}

// d_model = 256 fast path: one warp per row, every lane owns two float4 column groups
// (cols 4*lane..+3 and 128+4*lane..+3) so all global accesses are full 512-byte warp transactions.
template <typename LP_T>
__device__ __forceinline__ void store_lp4(LP_T* dst, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store_lp4<float>(float* dst, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(dst) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store_lp4<__nv_bfloat16>(__nv_bfloat16* dst, float a, float b, float c, float d) {
  uint2 u;
  u.x = float2_to_bf16x2(a, b);
  u.y = float2_to_bf16x2(c, d);
  *reinterpret_cast<uint2*>(dst) = u;
}

constexpr int kLn256Warps = 8;

template <typename LP_T>
__global__ void __launch_bounds__(kLn256Warps * 32) add_layernorm256_kernel(
    const float* __restrict__ x, const float* __restrict__ residual, const float* __restrict__ gamma,
    const float* __restrict__ beta, float eps, int64_t rows, float* __restrict__ out_f32,
    LP_T* __restrict__ out_lp, const float* __restrict__ pos, LP_T* __restrict__ out_pos_lp) {
  pdl_trigger();
  constexpr int C = 256;
  const int lane = threadIdx.x & 31;
  const float4 ga = *reinterpret_cast<const float4*>(gamma + lane * 4), gb = *reinterpret_cast<const float4*>(gamma + 128 + lane * 4);
  const float4 ba = *reinterpret_cast<const float4*>(beta + lane * 4), bb = *reinterpret_cast<const float4*>(beta + 128 + lane * 4);
  pdl_wait();
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kLn256Warps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int64_t o0 = row * C + lane * 4, o1 = o0 + 128;
  float4 a = *reinterpret_cast<const float4*>(x + o0), b = *reinterpret_cast<const float4*>(x + o1);
  if (residual) {
    const float4 ra = *reinterpret_cast<const float4*>(residual + o0), rb = *reinterpret_cast<const float4*>(residual + o1);
    a.x += ra.x; a.y += ra.y; a.z += ra.z; a.w += ra.w;
    b.x += rb.x; b.y += rb.y; b.z += rb.z; b.w += rb.w;
  }
  const float mean = warp_sum(((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w))) * (1.0f / C);
  a.x -= mean; a.y -= mean; a.z -= mean; a.w -= mean;
  b.x -= mean; b.y -= mean; b.z -= mean; b.w -= mean;
  const float sq = warp_sum(((a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w)) +
                            ((b.x * b.x + b.y * b.y) + (b.z * b.z + b.w * b.w)));
  const float rstd = rsqrtf(sq * (1.0f / C) + eps);
  a.x = a.x * rstd * ga.x + ba.x; a.y = a.y * rstd * ga.y + ba.y; a.z = a.z * rstd * ga.z + ba.z; a.w = a.w * rstd * ga.w + ba.w;
  b.x = b.x * rstd * gb.x + bb.x; b.y = b.y * rstd * gb.y + bb.y; b.z = b.z * rstd * gb.z + bb.z; b.w = b.w * rstd * gb.w + bb.w;
  if (out_f32) {
    *reinterpret_cast<float4*>(out_f32 + o0) = a;
    *reinterpret_cast<float4*>(out_f32 + o1) = b;
  }
  if (out_lp) {
    store_lp4<LP_T>(out_lp + o0, a.x, a.y, a.z, a.w);
    store_lp4<LP_T>(out_lp + o1, b.x, b.y, b.z, b.w);
  }
  if (out_pos_lp) {
    const float4 pa = *reinterpret_cast<const float4*>(pos + o0), pb = *reinterpret_cast<const float4*>(pos + o1);
    store_lp4<LP_T>(out_pos_lp + o0, a.x + pa.x, a.y + pa.y, a.z + pa.z, a.w + pa.w);
    store_lp4<LP_T>(out_pos_lp + o1, b.x + pb.x, b.y + pb.y, b.z + pb.z, b.w + pb.w);
  }
}

template <typename LP_T>
__global__ void add_cast_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                LP_T* __restrict__ out, int64_t n) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = from_float<LP_T>(a[i] + (b ? b[i] : 0.0f));
}

template <typename HT>
__global__ void __launch_bounds__(kRowWarps * 32) box_refine_kernel(
    const HT* __restrict__ h, int64_t ldh, const float* __restrict__ w3, const float* __restrict__ b3,
    const float* __restrict__ ref, float* __restrict__ new_ref, int64_t rows, int K) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
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
    float t = lane == 0 ? d[0] : (lane == 1 ? d[1] : (lane == 2 ? d[2] : d[3]));
    t += b3[lane] + inverse_sigmoidf_(ref[row * 4 + lane]);
    new_ref[row * 4 + lane] = sigmoidf_(t);
  }
}

// K == 256, bf16 hidden: every lane owns 8 consecutive k (one 128-bit load of h); its w3 slices are
// immutable and are fetched before the programmatic-dependency wait.
__global__ void __launch_bounds__(kRowWarps * 32) box_refine256_kernel(
    const __nv_bfloat16* __restrict__ h, int64_t ldh, const float* __restrict__ w3, const float* __restrict__ b3,
    const float* __restrict__ ref, float* __restrict__ new_ref, int64_t rows) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  float w[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w3 + j * 256 + lane * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(w3 + j * 256 + lane * 8 + 4));
    w[j][0] = a.x; w[j][1] = a.y; w[j][2] = a.z; w[j][3] = a.w;
    w[j][4] = b.x; w[j][5] = b.y; w[j][6] = b.z; w[j][7] = b.w;
  }
  const float bias = lane < 4 ? __ldg(b3 + lane) : 0.0f;
  pdl_wait();
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint4 hv = *reinterpret_cast<const uint4*>(h + row * ldh + lane * 8);
  const float rf = lane < 4 ? ref[row * 4 + lane] : 0.5f;
  const float2 h0 = bf16x2_to_float2(hv.x), h1 = bf16x2_to_float2(hv.y), h2 = bf16x2_to_float2(hv.z),
               h3 = bf16x2_to_float2(hv.w);
  const float hf[8] = {h0.x, h0.y, h1.x, h1.y, h2.x, h2.y, h3.x, h3.y};
  float d[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = fmaf(hf[k], w[j][k], acc);
    d[j] = warp_sum(acc);
  }
  if (lane < 4) {
    const float t = (lane == 0 ? d[0] : (lane == 1 ? d[1] : (lane == 2 ? d[2] : d[3]))) + bias + inverse_sigmoidf_(rf);
    new_ref[row * 4 + lane] = sigmoidf_(t);
  }
}

template <typename XT>
__global__ void __launch_bounds__(kRowWarps * 32) score_head_kernel(
    const XT* __restrict__ x, int64_t ldx, const float* __restrict__ w, const float* __restrict__ b,
    float* __restrict__ logits, float* __restrict__ scores, int32_t* __restrict__ labels, int64_t rows,
    int K, int nc, float* __restrict__ max_logit) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  float best = -INFINITY;
  int best_c = 0;
  for (int c = 0; c < nc; ++c) {
    float d = 0.0f;
    for (int k = lane; k < K; k += 32) d = fmaf(to_float<XT>(x[row * ldx + k]), w[c * K + k], d);
    d = warp_sum(d) + b[c];
    if (lane == 0 && logits) logits[row * nc + c] = d;
    if (d > best) { best = d; best_c = c; }  // first maximum wins, as torch.max
  }
  if (lane == 0) {
    if (scores) scores[row] = sigmoidf_(best);
    if (labels) labels[row] = best_c;
    if (max_logit) max_logit[row] = best;
  }
}

__global__ void sigmoid_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, int inverse) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    y[i] = inverse ? inverse_sigmoidf_(x[i]) : sigmoidf_(x[i]);
}

__global__ void pos2posemb_kernel(const float* __restrict__ pos, float* __restrict__ emb, int64_t rows,
                                  int n_coord, int F, float temperature) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = rows * n_coord * F;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx % F);
    const int64_t rc = idx / F;  // row * n_coord + coord
    const float p = pos[rc] * 6.283185307179586f;
    const float dim_t = powf(temperature, static_cast<float>(2 * (i / 2)) / static_cast<float>(F));
    const float e = p / dim_t;
    emb[idx] = (i & 1) ? cosf(e) : sinf(e);
  }
}

template <typename TO>
__global__ void linear_k4_relu_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                      const float* __restrict__ b, TO* __restrict__ y, int64_t rows, int N) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = rows * N;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(idx % N);
    const int64_t r = idx / N;
    float v = b ? b[n] : 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) v = fmaf(x[r * 4 + k], w[n * 4 + k], v);
    y[idx] = from_float<TO>(fmaxf(v, 0.0f));
  }
}

static unsigned ew_blocks(int64_t n) {
  return static_cast<unsigned>(n <= 0 ? 1 : ((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8));
}
static unsigned row_blocks(int64_t rows) { return static_cast<unsigned>((rows + kRowWarps - 1) / kRowWarps); }

// fp32 -> three-term bf16 expansion x = x0 + x1 + x2 (x0 = bf16(x), x1 = bf16(x - x0), x2 = bf16(x - x0 - x1)), written
// as SIX column blocks of K so that ONE bf16 tensor-core GEMM over K' = 6K computes an fp32-grade product:
//   activations (role 0): [x2 | x0 | x1 | x1 | x0 | x0]     weights (role 1): [w0 | w2 | w1 | w0 | w1 | w0]
//   -> sum_k = x2 w0 + x0 w2 + x1 w1 + x1 w0 + x0 w1 + x0 w0   (smallest terms first in the accumulation order)
// The dropped terms are O(2^-24): measured 3e-8 relative to rms against an fp64 product (the fp32 tolerance is 1e-4).
__global__ void split_bf16x3_kernel(const float* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out, int64_t ldo,
                                    int64_t M, int K, int role) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = M * K;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / K;
    const int c = static_cast<int>(i % K);
    const float v = x[r * ldx + c];
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h0);
    const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(r1 - __bfloat162float(h1));
    __nv_bfloat16* o = out + r * ldo + c;
    if (role == 0) { o[0] = h2; o[K] = h0; o[2 * K] = h1; o[3 * K] = h1; o[4 * K] = h0; o[5 * K] = h0; }
    else           { o[0] = h0; o[K] = h2; o[2 * K] = h1; o[3 * K] = h0; o[4 * K] = h1; o[5 * K] = h0; }
  }
}

}  // namespace moyolo

using namespace moyolo;

extern "C" int moyolo_add_layernorm(const float* x, const float* residual, const float* gamma,
                                    const float* beta, float eps, int64_t rows, int C, float* out_f32,
                                    void* out_lp, const float* pos, void* out_pos_lp, int lp_dtype,
                                    moyolo_stream_t stream) {
  MOYOLO_REQUIRE(x && gamma && beta, MOYOLO_ERR_BAD_ARG, "add_layernorm: null x/gamma/beta");
  MOYOLO_REQUIRE(rows >= 0 && C > 0, MOYOLO_ERR_BAD_SHAPE, "add_layernorm: bad rows/C");
  MOYOLO_REQUIRE(out_pos_lp == nullptr || pos != nullptr, MOYOLO_ERR_BAD_ARG,
                 "add_layernorm: out_pos_lp requested without pos");
  if (rows == 0) return MOYOLO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool fast = C == 256 && aligned16(x) && aligned16(gamma) && aligned16(beta) &&
                    (residual == nullptr || aligned16(residual)) && (out_f32 == nullptr || aligned16(out_f32)) &&
                    (out_lp == nullptr || aligned16(out_lp)) && (pos == nullptr || aligned16(pos)) &&
                    (out_pos_lp == nullptr || aligned16(out_pos_lp));
  if (fast && (lp_dtype == MOYOLO_BF16 || lp_dtype == MOYOLO_F32)) {
    const unsigned blocks = static_cast<unsigned>((rows + kLn256Warps - 1) / kLn256Warps);
    if (lp_dtype == MOYOLO_BF16)
      launch_k(add_layernorm256_kernel<__nv_bfloat16>, dim3(blocks), dim3(kLn256Warps * 32), 0, st, 
          x, residual, gamma, beta, eps, rows, out_f32, static_cast<__nv_bfloat16*>(out_lp), pos,
          static_cast<__nv_bfloat16*>(out_pos_lp));
    else
      launch_k(add_layernorm256_kernel<float>, dim3(blocks), dim3(kLn256Warps * 32), 0, st, 
          x, residual, gamma, beta, eps, rows, out_f32, static_cast<float*>(out_lp), pos,
          static_cast<float*>(out_pos_lp));
    return check_launch("add_layernorm256_kernel");
  }
  if (lp_dtype == MOYOLO_BF16) {
    launch_k(add_layernorm_kernel<__nv_bfloat16>, dim3(row_blocks(rows)), dim3(kRowWarps * 32), 0, st, 
        x, residual, gamma, beta, eps, rows, C, out_f32, static_cast<__nv_bfloat16*>(out_lp), pos,
        static_cast<__nv_bfloat16*>(out_pos_lp));
  } else if (lp_dtype == MOYOLO_F32) {
    launch_k(add_layernorm_kernel<float>, dim3(row_blocks(rows)), dim3(kRowWarps * 32), 0, st, 
        x, residual, gamma, beta, eps, rows, C, out_f32, static_cast<float*>(out_lp), pos,
        static_cast<float*>(out_pos_lp));
  } else {
    return fail(MOYOLO_ERR_UNSUPPORTED, "add_layernorm: unsupported lp_dtype %d", lp_dtype);
  }
  return check_launch("add_layernorm_kernel");
}

extern "C" int moyolo_add_cast(const float* a, const float* b, void* out_lp, int lp_dtype, int64_t n,
                               moyolo_stream_t stream) {
  MOYOLO_REQUIRE(a && out_lp && n >= 0, MOYOLO_ERR_BAD_ARG, "add_cast: bad arguments");
  if (n == 0) return MOYOLO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (lp_dtype == MOYOLO_BF16)
    launch_k(add_cast_kernel<__nv_bfloat16>, dim3(ew_blocks(n)), dim3(256), 0, st, a, b, static_cast<__nv_bfloat16*>(out_lp), n);
  else if (lp_dtype == MOYOLO_F32)
    launch_k(add_cast_kernel<float>, dim3(ew_blocks(n)), dim3(256), 0, st, a, b, static_cast<float*>(out_lp), n);
  else
    return fail(MOYOLO_ERR_UNSUPPORTED, "add_cast: unsupported lp_dtype %d", lp_dtype);
  return check_launch("add_cast_kernel");
}

extern "C" int moyolo_box_refine(const void* h, int64_t ldh, int h_dtype, const float* w3, const float* b3,
                                 const float* ref, float* new_ref, int64_t rows, int K,
                                 moyolo_stream_t stream) {
  MOYOLO_REQUIRE(h && w3 && b3 && ref && new_ref, MOYOLO_ERR_BAD_ARG, "box_refine: null pointer");
  MOYOLO_REQUIRE(rows >= 0 && K > 0 && ldh >= K, MOYOLO_ERR_BAD_SHAPE, "box_refine: bad sizes");
  if (rows == 0) return MOYOLO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h_dtype == MOYOLO_BF16 && K == 256 && aligned16(h) && (ldh * 2) % 16 == 0 && aligned16(w3))
    launch_k(box_refine256_kernel, dim3(row_blocks(rows)), dim3(kRowWarps * 32), 0, st,
             static_cast<const __nv_bfloat16*>(h), ldh, w3, b3, ref, new_ref, rows);
  else if (h_dtype == MOYOLO_BF16)
    launch_k(box_refine_kernel<__nv_bfloat16>, dim3(row_blocks(rows)), dim3(kRowWarps * 32), 0, st, 
        static_cast<const __nv_bfloat16*>(h), ldh, w3, b3, ref, new_ref, rows, K);
  else if (h_dtype == MOYOLO_F32)
    launch_k(box_refine_kernel<float>, dim3(row_blocks(rows)), dim3(kRowWarps * 32), 0, st, static_cast<const float*>(h), ldh,
                                                                        w3, b3, ref, new_ref, rows, K);
  else
    return fail(MOYOLO_ERR_UNSUPPORTED, "box_refine: unsupported h_dtype %d", h_dtype);
  return check_launch("box_refine_kernel");
}

extern "C" int moyolo_score_head(const void* x, int64_t ldx, int x_dtype, const float* w, const float* b,
                                 float* logits, float* scores, int32_t* labels, int64_t rows, int K, int nc,
                                 float* max_logit, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(x && w && b, MOYOLO_ERR_BAD_ARG, "score_head: null pointer");
  MOYOLO_REQUIRE(rows >= 0 && K > 0 && nc > 0 && ldx >= K, MOYOLO_ERR_BAD_SHAPE, "score_head: bad sizes");
  if (rows == 0) return MOYOLO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (x_dtype == MOYOLO_BF16)
    launch_k(score_head_kernel<__nv_bfloat16>, dim3(row_blocks(rows)), dim3(kRowWarps * 32), 0, st, 
        static_cast<const __nv_bfloat16*>(x), ldx, w, b, logits, scores, labels, rows, K, nc, max_logit);
  else if (x_dtype == MOYOLO_F32)
    launch_k(score_head_kernel<float>, dim3(row_blocks(rows)), dim3(kRowWarps * 32), 0, st, 
        static_cast<const float*>(x), ldx, w, b, logits, scores, labels, rows, K, nc, max_logit);
  else
    return fail(MOYOLO_ERR_UNSUPPORTED, "score_head: unsupported x_dtype %d", x_dtype);
  return check_launch("score_head_kernel");
}

extern "C" int moyolo_sigmoid(const float* x, float* y, int64_t n, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(x && y && n >= 0, MOYOLO_ERR_BAD_ARG, "sigmoid: bad arguments");
  if (n == 0) return MOYOLO_OK;
  launch_k(sigmoid_kernel, dim3(ew_blocks(n)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, y, n, 0);
  return check_launch("sigmoid_kernel");
}

extern "C" int moyolo_inverse_sigmoid(const float* x, float* y, int64_t n, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(x && y && n >= 0, MOYOLO_ERR_BAD_ARG, "inverse_sigmoid: bad arguments");
  if (n == 0) return MOYOLO_OK;
  launch_k(sigmoid_kernel, dim3(ew_blocks(n)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, y, n, 1);
  return check_launch("inverse_sigmoid_kernel");
}

extern "C" int moyolo_pos2posemb(const float* pos, float* emb, int64_t rows, int n_coord, int num_pos_feats,
                                 float temperature, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(pos && emb, MOYOLO_ERR_BAD_ARG, "pos2posemb: null pointer");
  MOYOLO_REQUIRE(rows >= 0 && n_coord > 0 && num_pos_feats > 0, MOYOLO_ERR_BAD_SHAPE, "pos2posemb: bad sizes");
  if (rows == 0) return MOYOLO_OK;
  launch_k(pos2posemb_kernel, dim3(ew_blocks(rows * n_coord * num_pos_feats)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      pos, emb, rows, n_coord, num_pos_feats, temperature);
  return check_launch("pos2posemb_kernel");
}

extern "C" int moyolo_linear_k4_relu(const float* x, const float* w, const float* b, void* y, int out_dtype,
                                     int64_t rows, int N, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(x && w && y, MOYOLO_ERR_BAD_ARG, "linear_k4_relu: null pointer");
  MOYOLO_REQUIRE(rows >= 0 && N > 0, MOYOLO_ERR_BAD_SHAPE, "linear_k4_relu: bad sizes");
  if (rows == 0) return MOYOLO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (out_dtype == MOYOLO_BF16)
    launch_k(linear_k4_relu_kernel<__nv_bfloat16>, dim3(ew_blocks(rows * N)), dim3(256), 0, st, 
        x, w, b, static_cast<__nv_bfloat16*>(y), rows, N);
  else if (out_dtype == MOYOLO_F32)
    launch_k(linear_k4_relu_kernel<float>, dim3(ew_blocks(rows * N)), dim3(256), 0, st, x, w, b, static_cast<float*>(y), rows, N);
  else
    return fail(MOYOLO_ERR_UNSUPPORTED, "linear_k4_relu: unsupported out_dtype %d", out_dtype);
  return check_launch("linear_k4_relu_kernel");
}

extern "C" int moyolo_split_bf16x3(const float* x, int64_t ldx, void* out, int64_t ldo, int64_t M, int K, int role,
                                   moyolo_stream_t stream) {
  MOYOLO_REQUIRE(x && out, MOYOLO_ERR_BAD_ARG, "split_bf16x3: null pointer");
  MOYOLO_REQUIRE(M >= 0 && K > 0 && ldx >= K && ldo >= 6 * static_cast<int64_t>(K) && (role == 0 || role == 1),
                 MOYOLO_ERR_BAD_SHAPE, "split_bf16x3: bad sizes");
  if (M == 0) return MOYOLO_OK;
  launch_k(split_bf16x3_kernel, dim3(ew_blocks(M * K)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, ldx,
           static_cast<__nv_bfloat16*>(out), ldo, M, K, role);
  return check_launch("split_bf16x3_kernel");
}
