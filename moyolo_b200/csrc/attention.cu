// Query self-attention of the decoder layer (nn.MultiheadAttention semantics,
// ultralytics/nn/modules/transformer.py:637-641; QIM self-attention MOTR/models/qim.py:276).
// Sequences are ragged (carried tracks + detect queries differ per lock-step sequence), keys never
// cross a sequence boundary. Flash-style: one CTA = (sequence, head, 16 queries); K/V are streamed
// through shared memory in 64-key tiles with an online softmax; fp32 math, exp via expf.
#include "common.cuh"

namespace moyolo {

constexpr int kAttThreads = 128;
constexpr int kQT = 16;   // queries per CTA (4 per warp)
constexpr int kKT = 64;   // keys per tile
constexpr int kQPW = 4;   // queries per warp

template <typename T, int DH>
__global__ void __launch_bounds__(kAttThreads) self_attention_kernel(
    const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, int64_t ldk, const T* __restrict__ v,
    int64_t ldv, T* __restrict__ out, int64_t ldo, int batch, const int32_t* __restrict__ row_offsets,
    const int32_t* __restrict__ seg_len, const float* __restrict__ attn_mask) {
  constexpr int DPL = DH / 32;  // output dims per lane
  __shared__ float s_k[kKT][DH + 1];
  __shared__ float s_v[kKT][DH];
  __shared__ __align__(16) float s_q[DH][kQT];       // [d][query] so 4 queries read as one float4
  __shared__ __align__(16) float s_p[kAttThreads / 32][kKT][kQPW];

  // locate (sequence, query tile) of this CTA
  int b = 0, tile = blockIdx.x, seq_start = 0, seq_len = 0;
  for (; b < batch; ++b) {
    seq_start = row_offsets[b];
    seq_len = seg_len ? seg_len[b] : row_offsets[b + 1] - seq_start;
    const int nt = (seq_len + kQT - 1) / kQT;
    if (tile < nt) break;
    tile -= nt;
  }
  if (b == batch) return;
  const int head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = tile * kQT;
  const float scale = rsqrtf(static_cast<float>(DH));

  for (int i = threadIdx.x; i < kQT * DH; i += kAttThreads) {
    const int qi = i / DH, d = i % DH;
    const int ql = q0 + qi;
    s_q[d][qi] = ql < seq_len ? to_float<T>(q[static_cast<int64_t>(seq_start + ql) * ldq + head * DH + d]) * scale
                              : 0.0f;
  }

  float m[kQPW], l[kQPW], o[kQPW][DPL];
#pragma unroll
  for (int i = 0; i < kQPW; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.0f;
#pragma unroll
    for (int e = 0; e < DPL; ++e) o[i][e] = 0.0f;
  }

  for (int k0 = 0; k0 < seq_len; k0 += kKT) {
    __syncthreads();  // previous tile fully consumed (also orders the s_q fill on the first pass)
    for (int i = threadIdx.x; i < kKT * DH; i += kAttThreads) {
      const int kj = i / DH, d = i % DH;
      const int kl = k0 + kj;
      const bool ok = kl < seq_len;
      const int64_t r = seq_start + kl;
      s_k[kj][d] = ok ? to_float<T>(k[r * ldk + head * DH + d]) : 0.0f;
      s_v[kj][d] = ok ? to_float<T>(v[r * ldv + head * DH + d]) : 0.0f;
    }
    __syncthreads();

    // scores for keys (lane, lane+32) x the warp's 4 queries
    float s0[kQPW], s1[kQPW];
#pragma unroll
    for (int i = 0; i < kQPW; ++i) { s0[i] = 0.0f; s1[i] = 0.0f; }
#pragma unroll 8
    for (int d = 0; d < DH; ++d) {
      const float ka = s_k[lane][d], kb = s_k[lane + 32][d];
      const float4 qq = *reinterpret_cast<const float4*>(&s_q[d][warp * kQPW]);
      s0[0] = fmaf(qq.x, ka, s0[0]); s1[0] = fmaf(qq.x, kb, s1[0]);
      s0[1] = fmaf(qq.y, ka, s0[1]); s1[1] = fmaf(qq.y, kb, s1[1]);
      s0[2] = fmaf(qq.z, ka, s0[2]); s1[2] = fmaf(qq.z, kb, s1[2]);
      s0[3] = fmaf(qq.w, ka, s0[3]); s1[3] = fmaf(qq.w, kb, s1[3]);
    }
    const bool ok0 = (k0 + lane) < seq_len, ok1 = (k0 + lane + 32) < seq_len;
    float corr[kQPW];
#pragma unroll
    for (int i = 0; i < kQPW; ++i) {
      const int ql = q0 + warp * kQPW + i;
      if (attn_mask != nullptr && ql < seq_len) {
        if (ok0) s0[i] += attn_mask[static_cast<int64_t>(ql) * seq_len + k0 + lane];
        if (ok1) s1[i] += attn_mask[static_cast<int64_t>(ql) * seq_len + k0 + lane + 32];
      }
      const float a = ok0 ? s0[i] : -INFINITY, c = ok1 ? s1[i] : -INFINITY;
      const float mt = warp_max(fmaxf(a, c));
      const float mn = fmaxf(m[i], mt);
      const float p0 = ok0 ? expf(a - mn) : 0.0f, p1 = ok1 ? expf(c - mn) : 0.0f;
      corr[i] = (m[i] == -INFINITY) ? 0.0f : expf(m[i] - mn);
      l[i] = l[i] * corr[i] + warp_sum(p0 + p1);
      m[i] = mn;
      s_p[warp][lane][i] = p0;
      s_p[warp][lane + 32][i] = p1;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kQPW; ++i)
#pragma unroll
      for (int e = 0; e < DPL; ++e) o[i][e] *= corr[i];
#pragma unroll 8
    for (int j = 0; j < kKT; ++j) {
      const float4 pp = *reinterpret_cast<const float4*>(&s_p[warp][j][0]);
#pragma unroll
      for (int e = 0; e < DPL; ++e) {
        const float vv = s_v[j][lane + 32 * e];
        o[0][e] = fmaf(pp.x, vv, o[0][e]);
        o[1][e] = fmaf(pp.y, vv, o[1][e]);
        o[2][e] = fmaf(pp.z, vv, o[2][e]);
        o[3][e] = fmaf(pp.w, vv, o[3][e]);
      }
    }
    __syncwarp();
  }

#pragma unroll
  for (int i = 0; i < kQPW; ++i) {
    const int ql = q0 + warp * kQPW + i;
    if (ql >= seq_len) continue;
    const float inv = 1.0f / l[i];
#pragma unroll
    for (int e = 0; e < DPL; ++e)
      out[static_cast<int64_t>(seq_start + ql) * ldo + head * DH + lane + 32 * e] = from_float<T>(o[i][e] * inv);
  }
}

}  // namespace moyolo

using namespace moyolo;

extern "C" int moyolo_self_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                     int64_t ldv, void* out, int64_t ldo, int dtype, int batch,
                                     const int32_t* row_offsets, const int32_t* row_offsets_host,
                                     const int32_t* seg_len, int n_heads, int head_dim, const float* attn_mask,
                                     moyolo_stream_t stream) {
  MOYOLO_REQUIRE(q && k && v && out && row_offsets && row_offsets_host, MOYOLO_ERR_BAD_ARG,
                 "self_attention: null pointer");
  MOYOLO_REQUIRE(batch > 0 && n_heads > 0, MOYOLO_ERR_BAD_ARG, "self_attention: bad batch/n_heads");
  MOYOLO_REQUIRE(head_dim == 32 || head_dim == 64, MOYOLO_ERR_UNSUPPORTED,
                 "self_attention: head_dim must be 32 or 64, got %d", head_dim);
  int64_t tiles = 0;
  for (int b = 0; b < batch; ++b) {
    const int n = row_offsets_host[b + 1] - row_offsets_host[b];
    MOYOLO_REQUIRE(n >= 0, MOYOLO_ERR_BAD_SHAPE, "self_attention: row_offsets must be non-decreasing");
    tiles += (n + kQT - 1) / kQT;
  }
  if (tiles == 0) return MOYOLO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(static_cast<unsigned>(tiles), n_heads);
#define LAUNCH(T, DH)                                                                              \
  self_attention_kernel<T, DH><<<grid, kAttThreads, 0, st>>>(                                       \
      static_cast<const T*>(q), ldq, static_cast<const T*>(k), ldk, static_cast<const T*>(v), ldv, \
      static_cast<T*>(out), ldo, batch, row_offsets, seg_len, attn_mask)
  if (dtype == MOYOLO_F32) {
    if (head_dim == 32) LAUNCH(float, 32); else LAUNCH(float, 64);
  } else if (dtype == MOYOLO_BF16) {
    if (head_dim == 32) LAUNCH(__nv_bfloat16, 32); else LAUNCH(__nv_bfloat16, 64);
  } else {
    return fail(MOYOLO_ERR_UNSUPPORTED, "self_attention: unsupported dtype %d", dtype);
  }
#undef LAUNCH
  return check_launch("self_attention_kernel");
}
