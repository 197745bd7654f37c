// CUDA-core GEMM: y = act(x . w^T + bias), fp32 accumulate in strict K order.
// This is the exact-parity engine (fp32 operands reproduce nn.Linear to ~1e-6 relative); the
// throughput engine for bf16 operands is the tcgen05 kernel in gemm_tcgen05.cu.
// Replaces nn.Linear call sites ultralytics/nn/modules/transformer.py:264,268,269,286,576-580.
#include "common.cuh"

namespace moyolo {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) linear_simt_kernel(const TI* __restrict__ x, int64_t ldx,
                                                          const TI* __restrict__ w,
                                                          const float* __restrict__ bias,
                                                          TO* __restrict__ y, int64_t ldy, int64_t M,
                                                          int N, int K, int relu,
                                                          const uint8_t* __restrict__ zero_rows) {
  pdl_trigger();
  pdl_wait();
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t m0 = static_cast<int64_t>(blockIdx.y) * BM;
  const int n0 = blockIdx.x * BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  // loader mapping: 256 threads cover a 64x16 tile, 4 elements each (one row, 4 consecutive k)
  const int lr = tid / 4;        // 0..63
  const int lk = (tid % 4) * 4;  // 0,4,8,12

  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = k0 + lk + i;
      const int64_t gm = m0 + lr;
      const int gn = n0 + lr;
      As[lk + i][lr] = (gm < M && kk < K) ? to_float<TI>(x[gm * ldx + kk]) : 0.0f;
      Ws[lk + i][lr] = (gn < N && kk < K) ? to_float<TI>(w[static_cast<int64_t>(gn) * K + kk]) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Ws[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t gm = m0 + ty * TM + i;
    if (gm >= M) continue;
    const bool zero = zero_rows != nullptr && zero_rows[gm] != 0;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn >= N) continue;
      float v = acc[i][j] + (bias ? bias[gn] : 0.0f);
      if (relu) v = fmaxf(v, 0.0f);
      if (zero) v = 0.0f;
      y[gm * ldy + gn] = from_float<TO>(v);
    }
  }
}

template <typename TI, typename TO>
static int launch_simt(const void* x, int64_t ldx, const void* w, const float* bias, void* y, int64_t ldy,
                       int64_t M, int N, int K, int relu, const uint8_t* zero_rows, cudaStream_t st) {
  dim3 grid((N + BN - 1) / BN, static_cast<unsigned>((M + BM - 1) / BM));
  launch_k(linear_simt_kernel<TI, TO>, dim3(grid), dim3(256), 0, st, static_cast<const TI*>(x), ldx, static_cast<const TI*>(w),
                                                   bias, static_cast<TO*>(y), ldy, M, N, K, relu, zero_rows);
  return check_launch("linear_simt_kernel");
}

int linear_simt(const void* x, int64_t ldx, const void* w, const float* bias, void* y, int64_t ldy,
                int64_t M, int N, int K, int in_dtype, int out_dtype, int relu, const uint8_t* zero_rows,
                cudaStream_t st) {
  if (in_dtype == MOYOLO_F32 && out_dtype == MOYOLO_F32)
    return launch_simt<float, float>(x, ldx, w, bias, y, ldy, M, N, K, relu, zero_rows, st);
  if (in_dtype == MOYOLO_F32 && out_dtype == MOYOLO_BF16)
    return launch_simt<float, __nv_bfloat16>(x, ldx, w, bias, y, ldy, M, N, K, relu, zero_rows, st);
  if (in_dtype == MOYOLO_BF16 && out_dtype == MOYOLO_F32)
    return launch_simt<__nv_bfloat16, float>(x, ldx, w, bias, y, ldy, M, N, K, relu, zero_rows, st);
  if (in_dtype == MOYOLO_BF16 && out_dtype == MOYOLO_BF16)
    return launch_simt<__nv_bfloat16, __nv_bfloat16>(x, ldx, w, bias, y, ldy, M, N, K, relu, zero_rows, st);
  return fail(MOYOLO_ERR_UNSUPPORTED, "moyolo_linear: unsupported dtype pair (%d -> %d)", in_dtype, out_dtype);
}

// implemented in gemm_tcgen05.cu
bool ffn_ln_tcgen05_supported(const void* x, int64_t ldx, const void* w1, const void* w2, const void* h, int64_t M, int C,
                              int F);
int ffn_ln_tcgen05(const void* x, int64_t ldx, const void* w1, const float* b1, const void* w2, const float* b2, void* h,
                   int F, const float* residual, const float* gamma, const float* beta, float eps, int64_t M,
                   float* out_f32, void* out_lp, const float* pos, void* out_pos_lp, cudaStream_t st);
bool linear_tall_supported(const void* x, int64_t ldx, const void* w, const void* y, int64_t ldy, int64_t M, int N, int K);
int linear_tall(const void* x, int64_t ldx, const void* w, const float* bias, void* y, int64_t ldy, int64_t M, int N,
                const uint8_t* zero_rows, int max_ctas, cudaStream_t st);
int linear_tcgen05(const void* x, int64_t ldx, const void* x2, int64_t ldx2, int n_split, const void* w,
                   const float* bias, void* y, int64_t ldy, int64_t M, int N, int K, int out_dtype, int relu,
                   const uint8_t* zero_rows, cudaStream_t st);
bool linear_tcgen05_supported(const void* x, int64_t ldx, const void* w, int64_t M, int N, int K);
bool linear_ln_tcgen05_supported(const void* x, int64_t ldx, const void* w, int64_t M, int N, int K);
int linear_ln_tcgen05(const void* x, int64_t ldx, const void* w, const float* bias, const float* residual,
                      const float* gamma, const float* beta, float eps, int64_t M, int K, float* out_f32, void* out_lp,
                      const float* pos, void* out_pos_lp, const float* score_w, const float* score_b, int score_nc,
                      float* logits, float* scores, int32_t* labels, cudaStream_t st);

}  // namespace moyolo

extern "C" int moyolo_linear(const void* x, int64_t ldx, const void* w, const float* bias, void* y,
                             int64_t ldy, int64_t M, int N, int K, int in_dtype, int out_dtype,
                             int epilogue, const uint8_t* zero_rows, int engine, moyolo_stream_t stream) {
  using namespace moyolo;
  MOYOLO_REQUIRE(x && w && y, MOYOLO_ERR_BAD_ARG, "moyolo_linear: null x/w/y pointer");
  MOYOLO_REQUIRE(M >= 0 && N > 0 && K > 0 && ldx >= K && ldy >= N, MOYOLO_ERR_BAD_SHAPE,
                 "moyolo_linear: bad sizes M=%lld N=%d K=%d ldx=%lld ldy=%lld", (long long)M, N, K,
                 (long long)ldx, (long long)ldy);
  MOYOLO_REQUIRE(epilogue == MOYOLO_EPI_NONE || epilogue == MOYOLO_EPI_RELU, MOYOLO_ERR_BAD_ARG,
                 "moyolo_linear: bad epilogue %d", epilogue);
  if (M == 0) return MOYOLO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int relu = epilogue == MOYOLO_EPI_RELU;
  if (engine == MOYOLO_GEMM_AUTO)
    engine = (in_dtype == MOYOLO_BF16 && linear_tcgen05_supported(x, ldx, w, M, N, K)) ? MOYOLO_GEMM_TCGEN05
                                                                                       : MOYOLO_GEMM_SIMT;
  if (engine == MOYOLO_GEMM_TCGEN05) {
    MOYOLO_REQUIRE(in_dtype == MOYOLO_BF16, MOYOLO_ERR_UNSUPPORTED,
                   "moyolo_linear: the tcgen05 engine takes bf16 operands");
    MOYOLO_REQUIRE(linear_tcgen05_supported(x, ldx, w, M, N, K), MOYOLO_ERR_ALIGNMENT,
                   "moyolo_linear: tcgen05 engine needs K%%64==0, N%%32==0, 16B-aligned x/w/ldx");
    return linear_tcgen05(x, ldx, nullptr, 0, 0, w, bias, y, ldy, M, N, K, out_dtype, relu, zero_rows, st);
  }
  MOYOLO_REQUIRE(engine == MOYOLO_GEMM_SIMT, MOYOLO_ERR_BAD_ARG, "moyolo_linear: bad engine %d", engine);
  return linear_simt(x, ldx, w, bias, y, ldy, M, N, K, in_dtype, out_dtype, relu, zero_rows, st);
}

extern "C" int moyolo_linear_tall(const void* x, int64_t ldx, const void* w, const float* bias, void* y, int64_t ldy,
                                  int64_t M, int N, int K, const uint8_t* zero_rows, int max_ctas,
                                  moyolo_stream_t stream) {
  using namespace moyolo;
  MOYOLO_REQUIRE(x && w && y, MOYOLO_ERR_BAD_ARG, "moyolo_linear_tall: null x/w/y pointer");
  MOYOLO_REQUIRE(M >= 0 && N > 0 && K > 0 && ldx >= K && ldy >= N && max_ctas >= 0, MOYOLO_ERR_BAD_SHAPE,
                 "moyolo_linear_tall: bad sizes");
  if (M == 0) return MOYOLO_OK;
  MOYOLO_REQUIRE(linear_tall_supported(x, ldx, w, y, ldy, M, N, K), MOYOLO_ERR_UNSUPPORTED,
                 "moyolo_linear_tall: needs bf16, K == 256, N %% 128 == 0, 16B-aligned x/w/y and row strides");
  return linear_tall(x, ldx, w, bias, y, ldy, M, N, zero_rows, max_ctas, static_cast<cudaStream_t>(stream));
}

extern "C" int moyolo_linear_dual(const void* x1, int64_t ldx1, const void* x2, int64_t ldx2, int n_split,
                                  const void* w, const float* bias, void* y, int64_t ldy, int64_t M, int N, int K,
                                  int out_dtype, moyolo_stream_t stream) {
  using namespace moyolo;
  MOYOLO_REQUIRE(x1 && x2 && w && y, MOYOLO_ERR_BAD_ARG, "moyolo_linear_dual: null pointer");
  MOYOLO_REQUIRE(M >= 0 && N > 0 && K > 0 && ldx1 >= K && ldx2 >= K && ldy >= N, MOYOLO_ERR_BAD_SHAPE,
                 "moyolo_linear_dual: bad sizes");
  if (M == 0) return MOYOLO_OK;
  MOYOLO_REQUIRE(linear_tcgen05_supported(x1, ldx1, w, M, N, K) && linear_tcgen05_supported(x2, ldx2, w, M, N, K),
                 MOYOLO_ERR_ALIGNMENT, "moyolo_linear_dual: needs K%%64==0, N%%32==0, 16B-aligned operands");
  return linear_tcgen05(x1, ldx1, x2, ldx2, n_split, w, bias, y, ldy, M, N, K, out_dtype, 0, nullptr,
                        static_cast<cudaStream_t>(stream));
}

extern "C" int moyolo_linear_add_layernorm(const void* x, int64_t ldx, const void* w, const float* bias,
                                           const float* residual, const float* gamma, const float* beta, float eps,
                                           int64_t M, int N, int K, float* out_f32, void* out_lp, const float* pos,
                                           void* out_pos_lp, moyolo_stream_t stream) {
  using namespace moyolo;
  MOYOLO_REQUIRE(x && w && gamma && beta, MOYOLO_ERR_BAD_ARG, "moyolo_linear_add_layernorm: null pointer");
  MOYOLO_REQUIRE(out_pos_lp == nullptr || pos != nullptr, MOYOLO_ERR_BAD_ARG,
                 "moyolo_linear_add_layernorm: out_pos_lp requested without pos");
  MOYOLO_REQUIRE(M >= 0 && K > 0 && ldx >= K, MOYOLO_ERR_BAD_SHAPE, "moyolo_linear_add_layernorm: bad sizes");
  if (M == 0) return MOYOLO_OK;
  MOYOLO_REQUIRE(linear_ln_tcgen05_supported(x, ldx, w, M, N, K), MOYOLO_ERR_UNSUPPORTED,
                 "moyolo_linear_add_layernorm: needs N == 256, K %% 64 == 0 and 16B-aligned bf16 operands");
  MOYOLO_REQUIRE((residual == nullptr || aligned16(residual)) && (out_f32 == nullptr || aligned16(out_f32)) &&
                     (out_lp == nullptr || aligned16(out_lp)) && (pos == nullptr || aligned16(pos)) &&
                     (out_pos_lp == nullptr || aligned16(out_pos_lp)),
                 MOYOLO_ERR_ALIGNMENT, "moyolo_linear_add_layernorm: row buffers must be 16-byte aligned");
  return linear_ln_tcgen05(x, ldx, w, bias, residual, gamma, beta, eps, M, K, out_f32, out_lp, pos, out_pos_lp, nullptr,
                           nullptr, 0, nullptr, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int moyolo_linear_add_layernorm_scores(const void* x, int64_t ldx, const void* w, const float* bias,
                                                  const float* residual, const float* gamma, const float* beta,
                                                  float eps, int64_t M, int N, int K, float* out_f32, void* out_lp,
                                                  const float* score_w, const float* score_b, int nc, float* logits,
                                                  float* scores, int32_t* labels, moyolo_stream_t stream) {
  using namespace moyolo;
  MOYOLO_REQUIRE(x && w && gamma && beta && score_w && score_b, MOYOLO_ERR_BAD_ARG,
                 "moyolo_linear_add_layernorm_scores: null pointer");
  MOYOLO_REQUIRE(M >= 0 && K > 0 && ldx >= K && nc > 0 && nc <= 8, MOYOLO_ERR_BAD_SHAPE,
                 "moyolo_linear_add_layernorm_scores: bad sizes (1 <= nc <= 8)");
  if (M == 0) return MOYOLO_OK;
  MOYOLO_REQUIRE(linear_ln_tcgen05_supported(x, ldx, w, M, N, K), MOYOLO_ERR_UNSUPPORTED,
                 "moyolo_linear_add_layernorm_scores: needs N == 256, K %% 64 == 0 and 16B-aligned bf16 operands");
  MOYOLO_REQUIRE((residual == nullptr || aligned16(residual)) && (out_f32 == nullptr || aligned16(out_f32)) &&
                     (out_lp == nullptr || aligned16(out_lp)),
                 MOYOLO_ERR_ALIGNMENT, "moyolo_linear_add_layernorm_scores: row buffers must be 16-byte aligned");
  return linear_ln_tcgen05(x, ldx, w, bias, residual, gamma, beta, eps, M, K, out_f32, out_lp, nullptr, nullptr, score_w,
                           score_b, nc, logits, scores, labels, static_cast<cudaStream_t>(stream));
}

extern "C" int moyolo_ffn_add_layernorm(const void* x, int64_t ldx, const void* w1, const float* b1, const void* w2,
                                        const float* b2, void* h, int F, const float* residual, const float* gamma,
                                        const float* beta, float eps, int64_t M, int C, float* out_f32, void* out_lp,
                                        const float* pos, void* out_pos_lp, moyolo_stream_t stream) {
  using namespace moyolo;
  MOYOLO_REQUIRE(x && w1 && w2 && h && gamma && beta, MOYOLO_ERR_BAD_ARG, "moyolo_ffn_add_layernorm: null pointer");
  MOYOLO_REQUIRE(out_pos_lp == nullptr || pos != nullptr, MOYOLO_ERR_BAD_ARG,
                 "moyolo_ffn_add_layernorm: out_pos_lp requested without pos");
  MOYOLO_REQUIRE(M >= 0 && C > 0 && F > 0 && ldx >= C, MOYOLO_ERR_BAD_SHAPE, "moyolo_ffn_add_layernorm: bad sizes");
  if (M == 0) return MOYOLO_OK;
  MOYOLO_REQUIRE(ffn_ln_tcgen05_supported(x, ldx, w1, w2, h, M, C, F), MOYOLO_ERR_UNSUPPORTED,
                 "moyolo_ffn_add_layernorm: needs d_model == 256, hidden in {256, 1024} and 16B-aligned bf16 operands");
  MOYOLO_REQUIRE((residual == nullptr || aligned16(residual)) && (out_f32 == nullptr || aligned16(out_f32)) &&
                     (out_lp == nullptr || aligned16(out_lp)) && (pos == nullptr || aligned16(pos)) &&
                     (out_pos_lp == nullptr || aligned16(out_pos_lp)),
                 MOYOLO_ERR_ALIGNMENT, "moyolo_ffn_add_layernorm: row buffers must be 16-byte aligned");
  return ffn_ln_tcgen05(x, ldx, w1, b1, w2, b2, h, F, residual, gamma, beta, eps, M, out_f32, out_lp, pos, out_pos_lp,
                        static_cast<cudaStream_t>(stream));
}
