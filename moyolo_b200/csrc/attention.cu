// Query self-attention of the decoder layer (nn.MultiheadAttention semantics,
// ultralytics/nn/modules/transformer.py:637-641; QIM self-attention MOTR/models/qim.py:276).
// Sequences are ragged (carried tracks + detect queries differ per lock-step sequence), keys never
// cross a sequence boundary. Flash-style: one CTA = (sequence, head, 16 queries); K/V are streamed
// through shared memory in 64-key tiles with an online softmax; fp32 math, exp via expf.
#include <stdlib.h>

#include <atomic>

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
  pdl_trigger();
  pdl_wait();
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


// ------------------------------------------------------------------------------------------------
// bf16 tensor-core variant (eval path: no attn_mask). One CTA = (sequence, head, 64 queries), four
// warps of 16 queries; K/V stream through a double-buffered cp.async ring in 64-key tiles; S = Q K^T
// and O += P V run on mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with the online softmax held in
// registers (the S accumulator fragment is re-used as the A fragment of P V). The problem is
// 300..1200 queries x head_dim 32: ~0.1 GFLOP per launch, so the design goal is latency (few dependent
// steps, no shared-memory round trip for P), not tensor throughput.
// ------------------------------------------------------------------------------------------------
constexpr int kMQT = 64;  // queries per CTA
constexpr int kMKT = 64;  // keys per tile

template <int DH>
__global__ void __launch_bounds__(128, 3) self_attention_mma_kernel(
    const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k, int64_t ldk,
    const __nv_bfloat16* __restrict__ v, int64_t ldv, __nv_bfloat16* __restrict__ out, int64_t ldo, int batch,
    const int32_t* __restrict__ row_offsets, const int32_t* __restrict__ seg_len) {
  pdl_trigger();
  pdl_wait();
  constexpr int PITCH = DH + 8;      // bf16 elements; (DH+8)*2 bytes keeps ldmatrix rows on distinct banks
  constexpr int KS = DH / 16;        // k-steps of S = Q K^T
  constexpr int NT = DH / 8;         // n-tiles of O
  constexpr int CPR = DH / 8;        // 16-byte chunks per row
  __shared__ __align__(16) __nv_bfloat16 s_q[kMQT][PITCH];
  __shared__ __align__(16) __nv_bfloat16 s_k[2][kMKT][PITCH];
  __shared__ __align__(16) __nv_bfloat16 s_v[2][kMKT][PITCH];

  int b = 0, tile = blockIdx.x, seq_start = 0, seq_len = 0;
  for (; b < batch; ++b) {
    seq_start = row_offsets[b];
    seq_len = seg_len ? seg_len[b] : row_offsets[b + 1] - seq_start;
    const int nt = (seq_len + kMQT - 1) / kMQT;
    if (tile < nt) break;
    tile -= nt;
  }
  if (b == batch) return;
  const int head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = tile * kMQT;
  const int n_kt = (seq_len + kMKT - 1) / kMKT;

  auto load_kv = [&](int kt, int buf) {
    for (int i = threadIdx.x; i < kMKT * CPR; i += 128) {
      const int r = i / CPR, c = i % CPR;
      const int kl = kt * kMKT + r;
      const bool ok = kl < seq_len;
      const int64_t row = seq_start + (ok ? kl : 0);
      cp_async16(smem_addr(&s_k[buf][r][c * 8]), k + row * ldk + head * DH + c * 8, ok);
      cp_async16(smem_addr(&s_v[buf][r][c * 8]), v + row * ldv + head * DH + c * 8, ok);
    }
  };
  for (int i = threadIdx.x; i < kMQT * CPR; i += 128) {
    const int r = i / CPR, c = i % CPR;
    const bool ok = q0 + r < seq_len;
    const int64_t row = seq_start + (ok ? q0 + r : 0);
    cp_async16(smem_addr(&s_q[r][c * 8]), q + row * ldq + head * DH + c * 8, ok);
  }
  load_kv(0, 0);
  cp_async_commit();

  uint32_t qa[KS][4];
  float o[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.0f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
  // softmax in base 2: exp(x*scale - m*scale) = exp2((x - m) * scale*log2(e))
  const float sl2 = rsqrtf(static_cast<float>(DH)) * 1.4426950408889634f;

  for (int kt = 0; kt < n_kt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < n_kt) {
      load_kv(kt + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kt == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        ldmatrix_x4(smem_addr(&s_q[warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][ks * 16 + (lane >> 4) * 8]), qa[ks]);
    }
    // ---- S = Q K^T for 64 keys: 8 n-tiles of 8 keys ----
    float s[kMKT / 8][4];
#pragma unroll
    for (int j = 0; j < kMKT / 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
#pragma unroll
      for (int kp = 0; kp < KS / 2; ++kp) {  // one ldmatrix.x4 covers 32 head dims = 2 k-steps
        uint32_t kb[4];
        ldmatrix_x4(smem_addr(&s_k[buf][j * 8 + (lane & 7)][kp * 32 + (lane >> 3) * 8]), kb);
        mma_bf16_16816(s[j], qa[2 * kp], kb[0], kb[1]);
        mma_bf16_16816(s[j], qa[2 * kp + 1], kb[2], kb[3]);
      }
    }
    // ---- online softmax (rows g and g+8 of the warp's 16 queries) ----
    const int kbase = kt * kMKT;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < kMKT / 8; ++j) {
      const int key = kbase + j * 8 + 2 * t;
      if (key >= seq_len) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
      if (key + 1 >= seq_len) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);  // finite: every tile holds >= 1 valid key
    const float c0 = exp2f((m0 - mn0) * sl2), c1 = exp2f((m1 - mn1) * sl2);
    m0 = mn0;
    m1 = mn1;
    float rs0 = 0.0f, rs1 = 0.0f;
    uint32_t pa[kMKT / 16][4];
#pragma unroll
    for (int j = 0; j < kMKT / 8; ++j) {
      const float p0 = exp2f((s[j][0] - mn0) * sl2), p1 = exp2f((s[j][1] - mn0) * sl2);
      const float p2 = exp2f((s[j][2] - mn1) * sl2), p3 = exp2f((s[j][3] - mn1) * sl2);
      rs0 += p0 + p1;
      rs1 += p2 + p3;
      pa[j >> 1][(j & 1) * 2 + 0] = float2_to_bf16x2(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = float2_to_bf16x2(p2, p3);
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
#pragma unroll
    for (int n = 0; n < NT; ++n) { o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1; }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < kMKT / 16; ++kk) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t vb[4];
        ldmatrix_x4_trans(smem_addr(&s_v[buf][kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][np * 16 + (lane >> 4) * 8]), vb);
        mma_bf16_16816(o[2 * np], pa[kk], vb[0], vb[1]);
        mma_bf16_16816(o[2 * np + 1], pa[kk], vb[2], vb[3]);
      }
    }
    __syncthreads();  // all warps done with `buf` before the next iteration's prefetch overwrites it
  }

  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    if (r0 < seq_len)
      *reinterpret_cast<uint32_t*>(out + static_cast<int64_t>(seq_start + r0) * ldo + head * DH + n * 8 + 2 * t) =
          float2_to_bf16x2(o[n][0] * i0, o[n][1] * i0);
    if (r1 < seq_len)
      *reinterpret_cast<uint32_t*>(out + static_cast<int64_t>(seq_start + r1) * ldo + head * DH + n * 8 + 2 * t) =
          float2_to_bf16x2(o[n][2] * i1, o[n][3] * i1);
  }
}


// ------------------------------------------------------------------------------------------------
// Split-key variant for head_dim 32 (the decoder's configuration): one CTA = (sequence, head, 16
// queries); its four warps hold the SAME 16 queries and take every fourth 32-key tile each, with a
// private double-buffered cp.async ring (no block-wide barrier inside the loop). The four partial
// (max, sum, O) triples are merged through shared memory at the end. At 300-400 queries this gives
// 4x more CTAs (all SMs busy) and a 3x shorter dependent chain than the 64-query kernel above.
// ------------------------------------------------------------------------------------------------
constexpr int kSQT = 16;  // queries per CTA
constexpr int kSKT = 32;  // keys per tile
constexpr int kSNB = 4;   // K/V ring depth per warp
constexpr size_t kSplitkSmem = sizeof(__nv_bfloat16) * (8 * kSNB * kSKT * (32 + 8) + 4 * kSQT * (32 + 8));

__global__ void __launch_bounds__(128) self_attention_splitk32_kernel(
    const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k, int64_t ldk,
    const __nv_bfloat16* __restrict__ v, int64_t ldv, __nv_bfloat16* __restrict__ out, int64_t ldo, int batch,
    const int32_t* __restrict__ row_offsets, const int32_t* __restrict__ seg_len) {
  pdl_trigger();
  pdl_wait();
  constexpr int DH = 32;
  constexpr int PITCH = DH + 8;
  constexpr int KS = DH / 16;
  constexpr int NT = DH / 8;
  constexpr int CPR = DH / 8;
  // dynamic shared memory: per warp a kSNB-deep K/V ring (all of a warp's tiles are in flight at once
  // for sequences up to 4*kSNB*32 = 512 keys) and a private copy of the Q tile
  extern __shared__ __align__(16) uint8_t att_smem[];
  typedef __nv_bfloat16 (*KvRing)[kSNB][kSKT][PITCH];
  typedef __nv_bfloat16 (*QTile)[kSQT][PITCH];
  KvRing s_k = reinterpret_cast<KvRing>(att_smem);  // [4][kSNB][kSKT][PITCH]; re-used as the fp32 partial-O buffer
  KvRing s_v = reinterpret_cast<KvRing>(att_smem + sizeof(__nv_bfloat16) * 4 * kSNB * kSKT * PITCH);
  QTile s_q = reinterpret_cast<QTile>(att_smem + sizeof(__nv_bfloat16) * 8 * kSNB * kSKT * PITCH);
  __shared__ float s_m[4][kSQT], s_l[4][kSQT];

  int b = 0, tile = blockIdx.x, seq_start = 0, seq_len = 0;
  for (; b < batch; ++b) {
    seq_start = row_offsets[b];
    seq_len = seg_len ? seg_len[b] : row_offsets[b + 1] - seq_start;
    const int nt = (seq_len + kSQT - 1) / kSQT;
    if (tile < nt) break;
    tile -= nt;
  }
  if (b == batch) return;
  const int head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = tile * kSQT;
  const int n_kt = (seq_len + kSKT - 1) / kSKT;

  auto load_kv = [&](int kt, int buf) {
    for (int i = lane; i < kSKT * CPR; i += 32) {
      const int r = i / CPR, c = i % CPR;
      const int kl = kt * kSKT + r;
      const bool ok = kl < seq_len;
      const int64_t row = seq_start + (ok ? kl : 0);
      cp_async16(smem_addr(&s_k[warp][buf][r][c * 8]), k + row * ldk + head * DH + c * 8, ok);
      cp_async16(smem_addr(&s_v[warp][buf][r][c * 8]), v + row * ldv + head * DH + c * 8, ok);
    }
  };
  for (int i = lane; i < kSQT * CPR; i += 32) {
    const int r = i / CPR, c = i % CPR;
    const bool ok = q0 + r < seq_len;
    const int64_t row = seq_start + (ok ? q0 + r : 0);
    cp_async16(smem_addr(&s_q[warp][r][c * 8]), q + row * ldq + head * DH + c * 8, ok);
  }
  // prologue: Q + the first kSNB-1 tiles of this warp, one commit group per tile
#pragma unroll
  for (int p = 0; p < kSNB - 1; ++p) {
    if (warp + 4 * p < n_kt) load_kv(warp + 4 * p, p);
    cp_async_commit();
  }

  uint32_t qa[KS][4];
  float o[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.0f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
  const float sl2 = rsqrtf(static_cast<float>(DH)) * 1.4426950408889634f;

  int it = 0;
  for (int kt = warp; kt < n_kt; kt += 4, ++it) {
    const int buf = it % kSNB;
    if (kt + 4 * (kSNB - 1) < n_kt) load_kv(kt + 4 * (kSNB - 1), (it + kSNB - 1) % kSNB);
    cp_async_commit();            // uniform group count: group `it` holds tile `it` of this warp
    cp_async_wait<kSNB - 1>();
    __syncwarp();
    if (it == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        ldmatrix_x4(smem_addr(&s_q[warp][(lane & 7) + ((lane >> 3) & 1) * 8][ks * 16 + (lane >> 4) * 8]), qa[ks]);
    }
    float s[kSKT / 8][4];
#pragma unroll
    for (int j = 0; j < kSKT / 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
      uint32_t kb[4];
      ldmatrix_x4(smem_addr(&s_k[warp][buf][j * 8 + (lane & 7)][(lane >> 3) * 8]), kb);
      mma_bf16_16816(s[j], qa[0], kb[0], kb[1]);
      mma_bf16_16816(s[j], qa[1], kb[2], kb[3]);
    }
    const int kbase = kt * kSKT;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < kSKT / 8; ++j) {
      const int key = kbase + j * 8 + 2 * t;
      if (key >= seq_len) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
      if (key + 1 >= seq_len) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);  // finite: every tile holds >= 1 valid key
    const float c0 = exp2f((m0 - mn0) * sl2), c1 = exp2f((m1 - mn1) * sl2);
    m0 = mn0;
    m1 = mn1;
    float rs0 = 0.0f, rs1 = 0.0f;
    uint32_t pa[kSKT / 16][4];
#pragma unroll
    for (int j = 0; j < kSKT / 8; ++j) {
      const float p0 = exp2f((s[j][0] - mn0) * sl2), p1 = exp2f((s[j][1] - mn0) * sl2);
      const float p2 = exp2f((s[j][2] - mn1) * sl2), p3 = exp2f((s[j][3] - mn1) * sl2);
      rs0 += p0 + p1;
      rs1 += p2 + p3;
      pa[j >> 1][(j & 1) * 2 + 0] = float2_to_bf16x2(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = float2_to_bf16x2(p2, p3);
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
#pragma unroll
    for (int n = 0; n < NT; ++n) { o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1; }
#pragma unroll
    for (int kk = 0; kk < kSKT / 16; ++kk) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t vb[4];
        ldmatrix_x4_trans(smem_addr(&s_v[warp][buf][kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][np * 16 + (lane >> 4) * 8]), vb);
        mma_bf16_16816(o[2 * np], pa[kk], vb[0], vb[1]);
        mma_bf16_16816(o[2 * np + 1], pa[kk], vb[2], vb[3]);
      }
    }
    __syncwarp();  // every lane is done with `buf` before the next prefetch overwrites it
  }
  cp_async_wait<0>();  // warps without a tile still own the Q copy group
  __syncwarp();

  // ---- merge the four key-range partials ----
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  float* po = reinterpret_cast<float*>(&s_k[warp][0][0][0]);  // [16][DH] fp32 = 2 KiB of this warp's own ring
  if (t == 0) {
    s_m[warp][g] = m0; s_m[warp][g + 8] = m1;
    s_l[warp][g] = l0; s_l[warp][g + 8] = l1;
  }
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    *reinterpret_cast<float2*>(po + g * DH + n * 8 + 2 * t) = make_float2(o[n][0], o[n][1]);
    *reinterpret_cast<float2*>(po + (g + 8) * DH + n * 8 + 2 * t) = make_float2(o[n][2], o[n][3]);
  }
  __syncthreads();
  const int r = threadIdx.x >> 3, c4 = (threadIdx.x & 7) * 4;
  if (q0 + r < seq_len) {
    float M = fmaxf(fmaxf(s_m[0][r], s_m[1][r]), fmaxf(s_m[2][r], s_m[3][r]));
    float L = 0.0f, acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float mw = s_m[w][r];
      const float a = mw == -INFINITY ? 0.0f : exp2f((mw - M) * sl2);
      L = fmaf(a, s_l[w][r], L);
      const float4 ov = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(&s_k[w][0][0][0]) + r * DH + c4);
      acc[0] = fmaf(a, ov.x, acc[0]);
      acc[1] = fmaf(a, ov.y, acc[1]);
      acc[2] = fmaf(a, ov.z, acc[2]);
      acc[3] = fmaf(a, ov.w, acc[3]);
    }
    const float inv = 1.0f / L;
    uint2 pk;
    pk.x = float2_to_bf16x2(acc[0] * inv, acc[1] * inv);
    pk.y = float2_to_bf16x2(acc[2] * inv, acc[3] * inv);
    *reinterpret_cast<uint2*>(out + static_cast<int64_t>(seq_start + q0 + r) * ldo + head * DH + c4) = pk;
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
  // tensor-core path: bf16, no additive mask, 4-byte aligned rows (eval mode of the decoder and QIM)
  const bool mma = dtype == MOYOLO_BF16 && attn_mask == nullptr && aligned16(q) && aligned16(k) && aligned16(v) &&
                   (ldq % 8 == 0) && (ldk % 8 == 0) && (ldv % 8 == 0) && (ldo % 2 == 0) &&
                   (reinterpret_cast<uintptr_t>(out) & 3u) == 0;
  static const bool allow_splitk = [] { const char* e = getenv("MOYOLO_ATT_SPLITK"); return !(e && e[0] == '0'); }();
  const bool splitk = allow_splitk && mma && head_dim == 32 && (ldo % 4 == 0) && (reinterpret_cast<uintptr_t>(out) & 7u) == 0;
  const int qt = splitk ? kSQT : (mma ? kMQT : kQT);
  int64_t tiles = 0;
  for (int b = 0; b < batch; ++b) {
    const int n = row_offsets_host[b + 1] - row_offsets_host[b];
    MOYOLO_REQUIRE(n >= 0, MOYOLO_ERR_BAD_SHAPE, "self_attention: row_offsets must be non-decreasing");
    tiles += (n + qt - 1) / qt;
  }
  if (tiles == 0) return MOYOLO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mma) {
    dim3 mgrid(static_cast<unsigned>(tiles), n_heads);
    const __nv_bfloat16 *qq = static_cast<const __nv_bfloat16*>(q), *kk = static_cast<const __nv_bfloat16*>(k),
                        *vv = static_cast<const __nv_bfloat16*>(v);
    if (splitk) {
      static DeviceOnce once;
      const int dev_ = DeviceOnce::current();
      if (!once.done(dev_)) {
        cudaError_t e = cudaFuncSetAttribute(self_attention_splitk32_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSplitkSmem));
        if (e != cudaSuccess) return fail(MOYOLO_ERR_CUDA, "cudaFuncSetAttribute(attention): %s", cudaGetErrorString(e));
        once.set(dev_);
      }
      launch_k(self_attention_splitk32_kernel, dim3(mgrid), dim3(128), kSplitkSmem, st, qq, ldq, kk, ldk, vv, ldv,
               static_cast<__nv_bfloat16*>(out), ldo, batch, row_offsets, seg_len);
    } else if (head_dim == 32)
      launch_k(self_attention_mma_kernel<32>, dim3(mgrid), dim3(128), 0, st, qq, ldq, kk, ldk, vv, ldv, static_cast<__nv_bfloat16*>(out),
                                                          ldo, batch, row_offsets, seg_len);
    else
      launch_k(self_attention_mma_kernel<64>, dim3(mgrid), dim3(128), 0, st, qq, ldq, kk, ldk, vv, ldv, static_cast<__nv_bfloat16*>(out),
                                                          ldo, batch, row_offsets, seg_len);
    return check_launch("self_attention_mma_kernel");
  }
  dim3 grid(static_cast<unsigned>(tiles), n_heads);
#define LAUNCH(T, DH)                                                                              \
  launch_k(self_attention_kernel<T, DH>, dim3(grid), dim3(kAttThreads), 0, st,                                        \
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
