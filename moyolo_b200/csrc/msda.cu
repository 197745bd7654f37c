// Multi-scale deformable attention gather for sm_100a.
//
// Replaces: multi_scale_deformable_attn_pytorch (ultralytics/nn/modules/utils.py:41-78), the
// softmax + sampling-location arithmetic of MSDeformAttn.forward (ultralytics/nn/modules/
// transformer.py:268-285) and the legacy im2col kernel (MOTR/models/ops/src/cuda/
// ms_deform_im2col_cuda.cuh:237-299). Written from the arithmetic, not from that kernel:
//
//  * work item = one (query row, head), owned by ONE WARP. The head's value row (head_dim channels,
//    64..256 bytes) is covered by G = head_dim*sizeof(T)/16 adjacent lanes holding one 128-bit slice
//    each, so the warp splits into 32/G sub-groups that gather different bilinear corners at the same
//    time: all 4*L*P corner rows of the item (48 at L=3, P=4) are in flight after ONE round of
//    128-bit loads (6 per lane), which is what bounds the latency of the small per-frame launches.
//  * phase 1 (setup): lane p owns sampling point p: softmax over the L*P logits with warp shuffles,
//    sampling location, floor/fractions; per corner {spatial index | -1, bilinear*attention weight}
//    is staged in shared memory (8 bytes per corner).
//  * phase 2 (gather): sub-group s walks corners s, s+32/G, ...; fp32 FMAs into per-lane partial sums.
//  * phase 3: recursive-halving shuffle reduction across the sub-groups (7 shuffles for bf16/32
//    instead of 24 for a plain butterfly); afterwards the 32 lanes hold the head's output channels
//    and issue one coalesced 64..256-byte store.
//
// The value tensor is read channel-last ([B, Lv, heads, head_dim], arbitrary position stride), i.e.
// straight out of the value_proj GEMM — no NCHW transposes, no [B*H, Dh, Q, L*P] temporaries.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "gather_common.cuh"

namespace moyolo {

struct MsdaParams {
  const void* value;
  int64_t value_batch_stride;  // elements
  int64_t value_pos_stride;    // elements
  int64_t value_head_stride;   // elements between heads of one position (head_dim = channel-last; Lv*head_dim = head-major)
  LevelTable lv;
  int batch;
  int n_heads;
  int n_points;
  // fused inputs
  const float* offsets;
  int64_t offsets_row_stride;
  const float* logits;
  int64_t logits_row_stride;
  const float* refer;
  int ref_levels;
  int ref_dim;
  int softmax_mode;
  // projection fused into the gather (msda_gather_proj_kernel): offsets|logits = xq . w^T + bias
  const void* xq;            // bf16 [rows, n_heads*head_dim], row stride xq_row_stride elements
  int64_t xq_row_stride;
  const void* w_offlog;      // bf16 [n_heads*L*P*3, n_heads*head_dim] = [sampling_offsets ; attention_weights]
  const float* b_offlog;     // fp32 [n_heads*L*P*3]
  // pre-normalised inputs
  const void* loc;
  const void* weights;
  // rows
  int64_t rows;
  int64_t rows_per_batch;
  const int32_t* row_offsets;
  void* out;
  int64_t out_row_stride;
};

constexpr int kGatherWarps = 8;  // items (warps) per CTA
constexpr int kGatherThreads = kGatherWarps * 32;
constexpr int kMaxPointsPerLane = 2;  // L*P <= 64

// Sampling location of one point in fused mode (transformer.py:276-282).
__device__ __forceinline__ void fused_location(const MsdaParams& p, int64_t row, int head, int pt,
                                               int level, float* loc_x, float* loc_y) {
  const int LP = p.lv.n * p.n_points;
  const float* off = p.offsets + row * p.offsets_row_stride + (static_cast<int64_t>(head) * LP + pt) * 2;
  const float ox = off[0], oy = off[1];
  const float* r = p.refer + (row * p.ref_levels + (p.ref_levels == 1 ? 0 : level)) * p.ref_dim;
  if (p.ref_dim == 4) {
    *loc_x = r[0] + ox / static_cast<float>(p.n_points) * r[2] * 0.5f;
    *loc_y = r[1] + oy / static_cast<float>(p.n_points) * r[3] * 0.5f;
  } else {
    *loc_x = r[0] + ox / static_cast<float>(p.lv.w[level]);
    *loc_y = r[1] + oy / static_cast<float>(p.lv.h[level]);
  }
}

// Phases 2 and 3 of a (row, head) item, shared by the gather kernels: sub-group sg of the warp walks the staged
// corners sg, sg + NSG, ... (U of them in flight per trip), then the sub-groups' partial sums are merged.
template <typename VT, int DH, int U>
__device__ __forceinline__ void gather_reduce_store(const MsdaParams& p, const int* s_off, const float* s_w, int NUp,
                                                    int b, int row, int head, int lane) {
  constexpr int G = DH * static_cast<int>(sizeof(VT)) / 16;  // lanes per value row
  constexpr int CH = 16 / static_cast<int>(sizeof(VT));      // channels per lane
  constexpr int NSG = 32 / G;                                // sub-groups per warp
  // ---------------- phase 2: gather, sub-group sg takes corners sg, sg + NSG, ... ----------------
  float acc[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) acc[k] = 0.0f;

  const int sg = lane / G, sub = lane % G;
  const VT* base = static_cast<const VT*>(p.value) + static_cast<int64_t>(b) * p.value_batch_stride +
                   head * DH + sub * CH;
  for (int u0 = sg; u0 < NUp; u0 += NSG * U) {
    uint4 v[U];
    float w[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int eo = s_off[u0 + i * NSG];
      w[i] = s_w[u0 + i * NSG];
      v[i] = ldg128_if(base + eo, eo >= 0);
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if constexpr (sizeof(VT) == 2) fma_bf16x8(acc, v[i], w[i]);
      else fma_f32x4(acc, v[i], w[i]);
    }
  }

  // ---------------- phase 3: recursive-halving reduction across the sub-groups ----------------
  // Each round the lane keeps one half of its channels and receives the partner's partial sums for
  // that half, so after log2(NSG) rounds every lane holds CH/NSG finished channels.
  int ch = 0;
  {
    int n = CH;
#pragma unroll
    for (int off = 16; off >= G; off >>= 1) {
      n >>= 1;
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < CH / 2; ++i) {
        if (i < n) {
          const float send = up ? acc[i] : acc[i + n];
          const float recv = __shfl_xor_sync(0xffffffffu, send, off);
          acc[i] = (up ? acc[i + n] : acc[i]) + recv;
        }
      }
      ch += up ? n : 0;
    }
  }
  constexpr int NF = CH / NSG;  // finished channels per lane: 1 or 2
  VT* o = static_cast<VT*>(p.out) + static_cast<int64_t>(row) * p.out_row_stride + head * DH + sub * CH + ch;
  if constexpr (sizeof(VT) == 2) {
    if constexpr (NF == 1) {
      const float hi = __shfl_xor_sync(0xffffffffu, acc[0], G);  // odd-channel partner (last round's bit)
      if ((lane & G) == 0) *reinterpret_cast<uint32_t*>(o) = float2_to_bf16x2(acc[0], hi);
    } else {
      *reinterpret_cast<uint32_t*>(o) = float2_to_bf16x2(acc[0], acc[1]);
    }
  } else {
    if constexpr (NF == 1) {
      *o = acc[0];
    } else {
      *reinterpret_cast<float2*>(o) = make_float2(acc[0], acc[1]);
    }
  }
}

// U = corner rounds kept in flight per loop trip (all of them when 4*L*P == U*32/G).
// NH / NP / NL: n_heads / n_points / n_levels as compile-time constants (0 = read from the parameters):
// the index arithmetic of the hot configurations then has no integer or float division and each lane
// handles exactly ceil(L*P/32) sampling points. All element offsets are 32-bit
// (the host falls back to the generic kernel when a tensor has >= 2^31 elements).
template <typename VT, int DH, bool FUSED, int U, int NH, int NP, int NL>
__global__ void __launch_bounds__(kGatherThreads) msda_gather_kernel(const MsdaParams p) {
  constexpr int G = DH * static_cast<int>(sizeof(VT)) / 16;  // lanes per value row
  constexpr int CH = 16 / static_cast<int>(sizeof(VT));      // channels per lane
  constexpr int NSG = 32 / G;                                // sub-groups per warp
  static_assert(G >= 4 && G <= 16 && CH >= NSG, "head row must be 64..256 bytes");
  pdl_trigger();

  extern __shared__ __align__(16) int smem_i[];
  constexpr int kPts = NL ? (NL * NP + 31) / 32 : kMaxPointsPerLane;  // sampling points per lane
  const int n_heads = NH ? NH : p.n_heads;
  const int n_points = NP ? NP : p.n_points;
  const int LP = (NL ? NL : p.lv.n) * n_points;
  const int NU = LP * 4;
  const int NUp = (NU + NSG * U - 1) / (NSG * U) * (NSG * U);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = static_cast<int>(blockIdx.x) * kGatherWarps + warp;
  const int n_rows = static_cast<int>(p.rows);
  if (item >= n_rows * n_heads) return;  // warp-uniform
  const int row = item / n_heads;
  const int head = item - row * n_heads;

  int* s_off = smem_i + warp * NUp * 2;                   // element offset of the corner row | -1
  float* s_w = reinterpret_cast<float*>(s_off + NUp);     // bilinear x attention weight
  const int ps = static_cast<int>(p.value_pos_stride);
  pdl_wait();

  // ---------------- phase 1: lane `l` owns sampling points l and l + 32 ----------------
  // Every global operand of the phase (logit, offset pair, reference box | location, weight) is
  // requested up front so the softmax shuffles overlap ONE memory round trip instead of following it.
  float aw[kPts], lx[kPts], ly[kPts];
  if (FUSED) {
    float lg[kPts];
    float2 off[kPts];
    float rx[kPts], ry[kPts], rw[kPts], rh[kPts];
    const float* lrow = p.logits + static_cast<int64_t>(row) * p.logits_row_stride + head * LP;
    const float2* orow = reinterpret_cast<const float2*>(p.offsets + static_cast<int64_t>(row) * p.offsets_row_stride) +
                         head * LP;
    const float* rrow = p.refer + static_cast<int64_t>(row) * (p.ref_levels * p.ref_dim);
#pragma unroll
    for (int i = 0; i < kPts; ++i) {
      const int pt = lane + i * 32;
      const bool ok = pt < LP;
      const int level = ok ? pt / n_points : 0;
      lg[i] = ok ? __ldg(lrow + pt) : -INFINITY;
      off[i] = ok ? __ldg(orow + pt) : make_float2(0.0f, 0.0f);
      const float* r = rrow + (p.ref_levels == 1 ? 0 : level) * p.ref_dim;
      if (p.ref_dim == 4) {
        const float4 r4 = __ldg(reinterpret_cast<const float4*>(r));  // 16-byte aligned: checked by the host
        rx[i] = r4.x; ry[i] = r4.y; rw[i] = r4.z; rh[i] = r4.w;
      } else {
        const float2 r2 = __ldg(reinterpret_cast<const float2*>(r));
        rx[i] = r2.x; ry[i] = r2.y; rw[i] = rh[i] = 0.0f;
      }
    }
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < kPts; ++i) m = fmaxf(m, lg[i]);
    m = warp_max(m);
    if (p.softmax_mode == MOYOLO_SOFTMAX_PLUS1) m = 0.0f;  // exp(x)/(1+sum exp(x)), no shift
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < kPts; ++i) {
      const int pt = lane + i * 32;
      aw[i] = pt < LP ? expf(lg[i] - m) : 0.0f;
      sum += aw[i];
    }
    sum = warp_sum(sum);
    if (p.softmax_mode == MOYOLO_SOFTMAX_PLUS1) sum += 1.0f;
    const float inv = 1.0f / sum;
    const float fnp = static_cast<float>(n_points);
#pragma unroll
    for (int i = 0; i < kPts; ++i) {
      aw[i] *= inv;
      const int pt = lane + i * 32;
      const int level = pt < LP ? pt / n_points : 0;
      // sampling location (transformer.py:276-282)
      if (p.ref_dim == 4) {
        lx[i] = rx[i] + off[i].x / fnp * rw[i] * 0.5f;
        ly[i] = ry[i] + off[i].y / fnp * rh[i] * 0.5f;
      } else {
        lx[i] = rx[i] + off[i].x / static_cast<float>(p.lv.w[level]);
        ly[i] = ry[i] + off[i].y / static_cast<float>(p.lv.h[level]);
      }
    }
  } else {
    const int64_t ibase = static_cast<int64_t>(item) * LP;
#pragma unroll
    for (int i = 0; i < kPts; ++i) {
      const int pt = lane + i * 32;
      if (pt < LP) {
        const float2 l2 = __ldg(reinterpret_cast<const float2*>(p.loc) + ibase + pt);
        lx[i] = l2.x;
        ly[i] = l2.y;
        aw[i] = __ldg(static_cast<const float*>(p.weights) + ibase + pt);
      } else {
        lx[i] = ly[i] = aw[i] = 0.0f;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kPts; ++i) {
    const int pt = lane + i * 32;
    if (pt < LP) {
      const int level = pt / n_points;
      const Corners c = make_corners(lx[i], ly[i], p.lv.h[level], p.lv.w[level], p.lv.start[level], aw[i]);
      *reinterpret_cast<int4*>(s_off + pt * 4) =
          make_int4(c.pos[0] < 0 ? -1 : c.pos[0] * ps, c.pos[1] < 0 ? -1 : c.pos[1] * ps,
                    c.pos[2] < 0 ? -1 : c.pos[2] * ps, c.pos[3] < 0 ? -1 : c.pos[3] * ps);
      *reinterpret_cast<float4*>(s_w + pt * 4) = make_float4(c.w[0], c.w[1], c.w[2], c.w[3]);
    }
  }
  for (int u = NU + lane; u < NUp; u += 32) { s_off[u] = -1; s_w[u] = 0.0f; }
  __syncwarp();

  const int b = p.batch == 1 ? 0 : batch_of_row(row, p.row_offsets, p.batch, p.rows_per_batch);
  gather_reduce_store<VT, DH, U>(p, s_off, s_w, NUp, b, row, head, lane);
}

// ------------------------------------------------------------------------------------------------
// Two work items per warp for L*P <= 16 (the decoder's 3 levels x 4 points): heads 2k and 2k+1 of one
// query row. Phase 1 then runs ONCE for both items with lanes 0..15 / 16..31 owning the points of the
// two heads (segmented 16-lane softmax), which removes ~40 % of the per-item instructions of the
// one-item kernel above (where only 12 of 32 lanes had a point) — the kernel is issue-bound at large
// batch — and keeps 12 corner rows per lane in flight. FUSED mode, 8 heads, head_dim 32.
// ------------------------------------------------------------------------------------------------
template <typename VT, int NP, int NL>
__global__ void __launch_bounds__(kGatherThreads) msda_gather_pair_kernel(const MsdaParams p) {
  constexpr int DH = 32, NH = 8;
  constexpr int G = DH * static_cast<int>(sizeof(VT)) / 16;
  constexpr int CH = 16 / static_cast<int>(sizeof(VT));
  constexpr int NSG = 32 / G;
  constexpr int LP = NP * NL;
  static_assert(LP <= 16, "one point per lane of a 16-lane half");
  constexpr int NU = LP * 4;
  constexpr int R = (NU + NSG - 1) / NSG;  // corner rounds per sub-group and item
  constexpr int NUp = R * NSG;
  pdl_trigger();

  extern __shared__ __align__(16) int smem_i[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pairi = static_cast<int>(blockIdx.x) * kGatherWarps + warp;
  if (pairi >= static_cast<int>(p.rows) * (NH / 2)) return;  // warp-uniform
  const int row = pairi / (NH / 2);
  const int hp = pairi % (NH / 2);
  int* s_off = smem_i + warp * (4 * NUp);                    // [2][NUp] element offsets | -1
  float* s_w = reinterpret_cast<float*>(s_off + 2 * NUp);    // [2][NUp] bilinear x attention weights
  const int ps = static_cast<int>(p.value_pos_stride);
  pdl_wait();

  // ---------------- phase 1: lane (half, pl) owns point pl of head 2*hp + half ----------------
  {
    const int half = lane >> 4, pl = lane & 15;
    const int head = hp * 2 + half;
    const bool ok = pl < LP;
    const int level = ok ? pl / NP : 0;
    const float* lrow = p.logits + static_cast<int64_t>(row) * p.logits_row_stride + head * LP;
    const float2* orow = reinterpret_cast<const float2*>(p.offsets + static_cast<int64_t>(row) * p.offsets_row_stride) +
                         head * LP;
    const float* r = p.refer + static_cast<int64_t>(row) * (p.ref_levels * p.ref_dim) +
                     (p.ref_levels == 1 ? 0 : level) * p.ref_dim;
    const float lg = ok ? __ldg(lrow + pl) : -INFINITY;
    const float2 off = ok ? __ldg(orow + pl) : make_float2(0.0f, 0.0f);
    float rx, ry, rw = 0.0f, rh = 0.0f;
    if (p.ref_dim == 4) {
      const float4 r4 = __ldg(reinterpret_cast<const float4*>(r));
      rx = r4.x; ry = r4.y; rw = r4.z; rh = r4.w;
    } else {
      const float2 r2 = __ldg(reinterpret_cast<const float2*>(r));
      rx = r2.x; ry = r2.y;
    }
    float m = lg;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (p.softmax_mode == MOYOLO_SOFTMAX_PLUS1) m = 0.0f;
    float aw = ok ? expf(lg - m) : 0.0f;
    float sum = aw;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (p.softmax_mode == MOYOLO_SOFTMAX_PLUS1) sum += 1.0f;
    aw *= 1.0f / sum;
    float lx, ly;
    if (p.ref_dim == 4) {
      lx = rx + off.x / static_cast<float>(NP) * rw * 0.5f;
      ly = ry + off.y / static_cast<float>(NP) * rh * 0.5f;
    } else {
      lx = rx + off.x / static_cast<float>(p.lv.w[level]);
      ly = ry + off.y / static_cast<float>(p.lv.h[level]);
    }
    if (ok) {
      const Corners c = make_corners(lx, ly, p.lv.h[level], p.lv.w[level], p.lv.start[level], aw);
      *reinterpret_cast<int4*>(s_off + half * NUp + pl * 4) =
          make_int4(c.pos[0] < 0 ? -1 : c.pos[0] * ps, c.pos[1] < 0 ? -1 : c.pos[1] * ps,
                    c.pos[2] < 0 ? -1 : c.pos[2] * ps, c.pos[3] < 0 ? -1 : c.pos[3] * ps);
      *reinterpret_cast<float4*>(s_w + half * NUp + pl * 4) = make_float4(c.w[0], c.w[1], c.w[2], c.w[3]);
    }
    if constexpr (NUp > NU) {
      for (int u = NU + pl; u < NUp; u += 16) { s_off[half * NUp + u] = -1; s_w[half * NUp + u] = 0.0f; }
    }
  }
  __syncwarp();

  // ---------------- phase 2: both items' corner rows in flight, then the FMAs ----------------
  const int sg = lane / G, sub = lane % G;
  const int b = p.batch == 1 ? 0 : batch_of_row(row, p.row_offsets, p.batch, p.rows_per_batch);
  const int64_t hs = p.value_head_stride;
  const VT* base = static_cast<const VT*>(p.value) + static_cast<int64_t>(b) * p.value_batch_stride +
                   hp * 2 * hs + sub * CH;
  float acc[2][CH];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[h][k] = 0.0f;

  auto accumulate = [&](float (&a)[CH], const uint4& v, float w) {
    if constexpr (sizeof(VT) == 2) fma_bf16x8(a, v, w);
    else fma_f32x4(a, v, w);
  };
  if constexpr (R <= 6) {
    uint4 v[2][R];
    float w[2][R];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int eo = s_off[h * NUp + sg + i * NSG];
        w[h][i] = s_w[h * NUp + sg + i * NSG];
        v[h][i] = ldg128_if(base + h * hs + eo, eo >= 0);
      }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < R; ++i) accumulate(acc[h], v[h][i], w[h][i]);
  } else {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint4 v[R];
      float w[R];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int eo = s_off[h * NUp + sg + i * NSG];
        w[i] = s_w[h * NUp + sg + i * NSG];
        v[i] = ldg128_if(base + h * hs + eo, eo >= 0);
      }
#pragma unroll
      for (int i = 0; i < R; ++i) accumulate(acc[h], v[i], w[i]);
    }
  }

  // ---------------- phase 3: recursive-halving reduction per item, coalesced stores ----------------
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    int ch = 0, n = CH;
#pragma unroll
    for (int off = 16; off >= G; off >>= 1) {
      n >>= 1;
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < CH / 2; ++i) {
        if (i < n) {
          const float send = up ? acc[h][i] : acc[h][i + n];
          const float recv = __shfl_xor_sync(0xffffffffu, send, off);
          acc[h][i] = (up ? acc[h][i + n] : acc[h][i]) + recv;
        }
      }
      ch += up ? n : 0;
    }
    constexpr int NF = CH / NSG;  // finished channels per lane: 1
    static_assert(NF == 1, "head_dim 32: one finished channel per lane");
    VT* o = static_cast<VT*>(p.out) + static_cast<int64_t>(row) * p.out_row_stride + (hp * 2 + h) * DH + sub * CH + ch;
    if constexpr (sizeof(VT) == 2) {
      const float hi = __shfl_xor_sync(0xffffffffu, acc[h][0], G);
      if ((lane & G) == 0) *reinterpret_cast<uint32_t*>(o) = float2_to_bf16x2(acc[h][0], hi);
    } else {
      *o = acc[h][0];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Gather with the sampling_offsets | attention_weights projection fused in (transformer.py:268-269 + 271-285):
// one CTA = 8 query rows x ONE head. The head's 3*L*P projection rows (24 offset + 12 logit rows of the
// [288, 256] weight at L=3, P=4; 18 KB of bf16) are immutable, so they are requested with cp.async BEFORE the
// programmatic-dependency wait; after it the 8 activation rows (4 KB) arrive in one round trip, five warps
// run the [8(16) x 256] x [256 x 40] product on mma.sync.m16n8k16 (tiles far below a tcgen05 128-row tile),
// and the 8 warps then run the usual phases 1-3 for (row, head) items with offsets / logits read from shared
// memory. Removes one dependent launch (and the fp32 [R, 288] round trip through L2) per decoder layer.
// bf16 value, head_dim 32, 8 heads, d_model 256.
// ------------------------------------------------------------------------------------------------
constexpr int kProjRows = 8;

template <int NP, int NL>
__global__ void __launch_bounds__(kGatherThreads) msda_gather_proj_kernel(const MsdaParams p) {
  using VT = __nv_bfloat16;
  constexpr int DH = 32, NH = 8, K = NH * DH, PITCH = K + 8;
  constexpr int LP = NP * NL, NO = LP * 3, NT = (NO + 7) / 8, NOp = NT * 8;
  constexpr int G = 4, NSG = 8, NU = LP * 4, U = 6;
  constexpr int NUp = (NU + NSG * U - 1) / (NSG * U) * (NSG * U);
  static_assert(LP <= 32 && K / 32 == kGatherWarps && kProjRows == kGatherWarps, "one point per lane, one K slice per warp");
  pdl_trigger();
  extern __shared__ __align__(16) int smem_i[];
  __nv_bfloat16* sW = reinterpret_cast<__nv_bfloat16*>(smem_i);      // [NOp][PITCH]
  __nv_bfloat16* sX = sW + NOp * PITCH;                               // [8][PITCH]
  float* sB = reinterpret_cast<float*>(sX + kProjRows * PITCH);       // [NOp]
  float* sOL = sB + NOp;                                              // [8][NOp]
  float* sP = sOL + kProjRows * NOp;                                  // [8 warps][8][NOp] K-slice partial sums
  int* sStage = reinterpret_cast<int*>(sP + kGatherWarps * kProjRows * NOp);  // per warp: [NUp] offsets, [NUp] weights

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int row0 = static_cast<int>(blockIdx.x) * kProjRows;
  const int n_rows = static_cast<int>(p.rows);
  // projection row n of this head -> row of the stacked [offsets ; logits] weight
  auto src_row = [&](int n) { return n < 2 * LP ? head * 2 * LP + n : NH * 2 * LP + head * LP + (n - 2 * LP); };
  {
    const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(p.w_offlog);
    for (int i = threadIdx.x; i < NO * (K / 8); i += kGatherThreads) {
      const int n = i / (K / 8), c = i % (K / 8);
      cp_async16(smem_addr(sW + n * PITCH + c * 8), w + static_cast<int64_t>(src_row(n)) * K + c * 8, true);
    }
    for (int i = threadIdx.x; i < (NOp - NO) * (K / 8); i += kGatherThreads) {
      const int n = NO + i / (K / 8), c = i % (K / 8);
      *reinterpret_cast<uint4*>(sW + n * PITCH + c * 8) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (threadIdx.x < NOp) sB[threadIdx.x] = threadIdx.x < NO ? __ldg(p.b_offlog + src_row(threadIdx.x)) : 0.0f;
    cp_async_commit();
  }
  pdl_wait();
  {
    const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(p.xq);
    for (int i = threadIdx.x; i < kProjRows * (K / 8); i += kGatherThreads) {
      const int r = i / (K / 8), c = i % (K / 8);
      const bool ok = row0 + r < n_rows;
      cp_async16(smem_addr(sX + r * PITCH + c * 8), x + static_cast<int64_t>(ok ? row0 + r : 0) * p.xq_row_stride + c * 8, ok);
    }
    cp_async_commit();
  }
  // this warp's item: reference box requested now, consumed after the projection
  const int row = row0 + warp;
  const bool item_ok = row < n_rows;
  const int level = lane < LP ? lane / NP : 0;
  float rx = 0.0f, ry = 0.0f, rw = 0.0f, rh = 0.0f;
  if (item_ok) {
    const float* r = p.refer + static_cast<int64_t>(row) * (p.ref_levels * p.ref_dim) + (p.ref_levels == 1 ? 0 : level) * p.ref_dim;
    if (p.ref_dim == 4) {
      const float4 r4 = __ldg(reinterpret_cast<const float4*>(r));
      rx = r4.x; ry = r4.y; rw = r4.z; rh = r4.w;
    } else {
      const float2 r2 = __ldg(reinterpret_cast<const float2*>(r));
      rx = r2.x; ry = r2.y;
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  // ---- projection: warp w owns the K slice [32w, 32w+32) for all NT n-tiles (independent accumulators, two
  // MMAs each); rows 8..15 of the m16 tile mirror rows 0..7. The 8 partial [8 x NOp] tiles are summed through
  // shared memory together with the bias.
  {
    uint32_t a0[4], a1[4];
    ldmatrix_x4(smem_addr(sX + (lane & 7) * PITCH + warp * 32 + (lane >> 4) * 8), a0);
    ldmatrix_x4(smem_addr(sX + (lane & 7) * PITCH + warp * 32 + 16 + (lane >> 4) * 8), a1);
    float c[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      uint32_t kb[4];
      ldmatrix_x4(smem_addr(sW + (j * 8 + (lane & 7)) * PITCH + warp * 32 + (lane >> 3) * 8), kb);
      c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.0f;
      mma_bf16_16816(c[j], a0, kb[0], kb[1]);
      mma_bf16_16816(c[j], a1, kb[2], kb[3]);
    }
    const int g = lane >> 2, t = lane & 3;
    float* part = sP + (warp * kProjRows + g) * NOp + 2 * t;
#pragma unroll
    for (int j = 0; j < NT; ++j) *reinterpret_cast<float2*>(part + j * 8) = make_float2(c[j][0], c[j][1]);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < kProjRows * NOp; o += kGatherThreads) {
    float acc = sB[o % NOp];
#pragma unroll
    for (int w8 = 0; w8 < kGatherWarps; ++w8) acc += sP[w8 * kProjRows * NOp + o];
    sOL[o] = acc;
  }
  __syncthreads();
  if (!item_ok) return;  // warp-uniform, after the last block-wide barrier

  // ---------------- phase 1: lane pt owns sampling point pt ----------------
  int* s_off = sStage + warp * NUp * 2;
  float* s_w = reinterpret_cast<float*>(s_off + NUp);
  const int ps = static_cast<int>(p.value_pos_stride);
  {
    const bool ok = lane < LP;
    const float* ol = sOL + warp * NOp;
    const float lg = ok ? ol[2 * LP + lane] : -INFINITY;
    const float2 off = ok ? *reinterpret_cast<const float2*>(ol + 2 * lane) : make_float2(0.0f, 0.0f);
    float m = warp_max(lg);
    if (p.softmax_mode == MOYOLO_SOFTMAX_PLUS1) m = 0.0f;
    float aw = ok ? expf(lg - m) : 0.0f;
    float sum = warp_sum(aw);
    if (p.softmax_mode == MOYOLO_SOFTMAX_PLUS1) sum += 1.0f;
    aw *= 1.0f / sum;
    float lx, ly;
    if (p.ref_dim == 4) {
      lx = rx + off.x / static_cast<float>(NP) * rw * 0.5f;
      ly = ry + off.y / static_cast<float>(NP) * rh * 0.5f;
    } else {
      lx = rx + off.x / static_cast<float>(p.lv.w[level]);
      ly = ry + off.y / static_cast<float>(p.lv.h[level]);
    }
    if (ok) {
      const Corners c = make_corners(lx, ly, p.lv.h[level], p.lv.w[level], p.lv.start[level], aw);
      *reinterpret_cast<int4*>(s_off + lane * 4) =
          make_int4(c.pos[0] < 0 ? -1 : c.pos[0] * ps, c.pos[1] < 0 ? -1 : c.pos[1] * ps,
                    c.pos[2] < 0 ? -1 : c.pos[2] * ps, c.pos[3] < 0 ? -1 : c.pos[3] * ps);
      *reinterpret_cast<float4*>(s_w + lane * 4) = make_float4(c.w[0], c.w[1], c.w[2], c.w[3]);
    }
    for (int u = NU + lane; u < NUp; u += 32) { s_off[u] = -1; s_w[u] = 0.0f; }
  }
  __syncwarp();
  const int b = p.batch == 1 ? 0 : batch_of_row(row, p.row_offsets, p.batch, p.rows_per_batch);
  gather_reduce_store<VT, DH, U>(p, s_off, s_w, NUp, b, row, head, lane);
  (void)G;
}

// Generic kernel: any head_dim / dtype (incl. fp64 for the legacy FFI known-answer test,
// MOTR/models/ops/test.py:21-60). One thread per output scalar; correctness path, not a fast path.
template <typename VT, typename AT, bool FUSED>
__global__ void msda_generic_kernel(const MsdaParams p, int head_dim) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = p.rows * p.n_heads * head_dim;
  const int LP = p.lv.n * p.n_points;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % head_dim);
    const int head = static_cast<int>((idx / head_dim) % p.n_heads);
    const int64_t row = idx / (static_cast<int64_t>(head_dim) * p.n_heads);
    const int b = batch_of_row(row, p.row_offsets, p.batch, p.rows_per_batch);
    const VT* base = static_cast<const VT*>(p.value) + static_cast<int64_t>(b) * p.value_batch_stride +
                     head * head_dim + c;
    AT m = 0, s = 1;
    if (FUSED) {
      float mx = -INFINITY;
      for (int pt = 0; pt < LP; ++pt)
        mx = fmaxf(mx, p.logits[row * p.logits_row_stride + head * LP + pt]);
      if (p.softmax_mode == MOYOLO_SOFTMAX_PLUS1) mx = 0.0f;
      float sum = 0.0f;
      for (int pt = 0; pt < LP; ++pt)
        sum += expf(p.logits[row * p.logits_row_stride + head * LP + pt] - mx);
      if (p.softmax_mode == MOYOLO_SOFTMAX_PLUS1) sum += 1.0f;
      m = mx;
      s = sum;
    }
    AT acc = 0;
    for (int pt = 0; pt < LP; ++pt) {
      const int level = pt / p.n_points;
      const int H = p.lv.h[level], W = p.lv.w[level];
      AT lx, ly, a;
      if (FUSED) {
        float fx, fy;
        fused_location(p, row, head, pt, level, &fx, &fy);
        lx = fx;
        ly = fy;
        a = expf(p.logits[row * p.logits_row_stride + head * LP + pt] - static_cast<float>(m)) /
            static_cast<float>(s);
      } else {
        const int64_t i = (row * p.n_heads + head) * LP + pt;
        lx = static_cast<const AT*>(p.loc)[i * 2];
        ly = static_cast<const AT*>(p.loc)[i * 2 + 1];
        a = static_cast<const AT*>(p.weights)[i];
      }
      const AT x = lx * W - AT(0.5), y = ly * H - AT(0.5);
      if (!(x > AT(-1) && y > AT(-1) && x < AT(W) && y < AT(H))) continue;
      const AT xf = floor(x), yf = floor(y);
      const AT fx = x - xf, fy = y - yf;
      const int x0 = static_cast<int>(xf), y0 = static_cast<int>(yf);
      AT v00 = 0, v01 = 0, v10 = 0, v11 = 0;
      auto fetch = [&](int yy, int xx) -> AT {
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) return AT(0);
        const VT raw = base[static_cast<int64_t>(p.lv.start[level] + yy * W + xx) * p.value_pos_stride];
        return static_cast<AT>(raw);
      };
      v00 = fetch(y0, x0);
      v01 = fetch(y0, x0 + 1);
      v10 = fetch(y0 + 1, x0);
      v11 = fetch(y0 + 1, x0 + 1);
      const AT sample = (AT(1) - fy) * (AT(1) - fx) * v00 + (AT(1) - fy) * fx * v01 +
                        fy * (AT(1) - fx) * v10 + fy * fx * v11;
      acc += a * sample;
    }
    static_cast<VT*>(p.out)[row * p.out_row_stride + head * head_dim + c] = static_cast<VT>(acc);
  }
}


template <typename VT, int DH, bool FUSED, int U, int NH, int NP, int NL>
static int launch_fast_u(const MsdaParams& p, cudaStream_t st) {
  constexpr int G = DH * static_cast<int>(sizeof(VT)) / 16;
  constexpr int NSG = 32 / G;
  const int NU = p.lv.n * p.n_points * 4;
  const int NUp = (NU + NSG * U - 1) / (NSG * U) * (NSG * U);
  const size_t smem = static_cast<size_t>(kGatherWarps) * NUp * 8;
  const int64_t items = p.rows * p.n_heads;
  const int64_t blocks = (items + kGatherWarps - 1) / kGatherWarps;
  if (blocks == 0) return MOYOLO_OK;
  launch_k(msda_gather_kernel<VT, DH, FUSED, U, NH, NP, NL>, dim3(static_cast<unsigned>(blocks)), dim3(kGatherThreads), smem,
           st, p);
  return check_launch("msda_gather_kernel");
}

template <typename VT, int DH, bool FUSED, int U>
static int launch_fast_hp(const MsdaParams& p, cudaStream_t st) {
  // compile-time n_heads / n_points / n_levels for the decoder's configurations:
  // yolo_track.yaml (8 heads, 3 levels x 4 points) and the L=4, P=8 point of the C5 sweep
  if (p.n_heads == 8 && p.n_points == 4 && p.lv.n == 3) return launch_fast_u<VT, DH, FUSED, U, 8, 4, 3>(p, st);
  if (p.n_heads == 8 && p.n_points == 8 && p.lv.n == 4) return launch_fast_u<VT, DH, FUSED, U, 8, 8, 4>(p, st);
  return launch_fast_u<VT, DH, FUSED, U, 0, 0, 0>(p, st);
}

template <typename VT, int NP, int NL>
static int launch_pair(const MsdaParams& p, cudaStream_t st) {
  constexpr int NSG = 32 / (32 * static_cast<int>(sizeof(VT)) / 16);
  constexpr int NUp = (NP * NL * 4 + NSG - 1) / NSG * NSG;
  const size_t smem = static_cast<size_t>(kGatherWarps) * 4 * NUp * 4;
  const int64_t pairs = p.rows * 4;
  const int64_t blocks = (pairs + kGatherWarps - 1) / kGatherWarps;
  if (blocks == 0) return MOYOLO_OK;
  launch_k(msda_gather_pair_kernel<VT, NP, NL>, dim3(static_cast<unsigned>(blocks)), dim3(kGatherThreads), smem, st, p);
  return check_launch("msda_gather_pair_kernel");
}

static bool pair_enabled() {
  static const bool on = [] { const char* e = getenv("MOYOLO_GATHER_PAIR"); return !(e && e[0] == '0'); }();
  return on;
}

template <typename VT, int DH, bool FUSED>
static int launch_fast(const MsdaParams& p, cudaStream_t st) {
  if constexpr (FUSED && DH == 32) {
    // two items per warp pay off once the launch is throughput- rather than latency-bound
    // (measured cross-over between 382 and 3056 rows on B200; see profiles/)
    if (pair_enabled() && p.rows >= 1024 && p.n_heads == 8 && p.n_points == 4 && p.lv.n == 3)
      return launch_pair<VT, 4, 3>(p, st);
  }
  constexpr int NSG = 32 / (DH * static_cast<int>(sizeof(VT)) / 16);
  const int rounds = (p.lv.n * p.n_points * 4 + NSG - 1) / NSG;  // corner rounds per sub-group
  // keep every round in flight when there are few (6 at L=3,P=4 bf16); otherwise trips of 8
  if (rounds % 6 == 0 && rounds <= 12) return launch_fast_hp<VT, DH, FUSED, 6>(p, st);
  return launch_fast_hp<VT, DH, FUSED, 8>(p, st);
}

template <bool FUSED>
static int dispatch(const MsdaParams& p, int value_dtype, int aux_dtype, int head_dim, cudaStream_t st) {
  const int LP = p.lv.n * p.n_points;
  const int esz = value_dtype == MOYOLO_BF16 ? 2 : (value_dtype == MOYOLO_F32 ? 4 : 8);
  const bool fast_ok = (value_dtype != MOYOLO_F64) && (aux_dtype == MOYOLO_F32) &&
                       (head_dim == 32 || head_dim == 64) && aligned16(p.value) && aligned16(p.out) &&
                       (p.value_pos_stride * esz) % 16 == 0 && (p.value_batch_stride * esz) % 16 == 0 &&
                       (p.out_row_stride * esz) % 16 == 0 && LP <= 32 * kMaxPointsPerLane &&
                       (FUSED ? ((reinterpret_cast<uintptr_t>(p.offsets) & 7u) == 0 && p.offsets_row_stride % 2 == 0 &&
                                 (reinterpret_cast<uintptr_t>(p.refer) & (p.ref_dim == 4 ? 15u : 7u)) == 0)
                              : (reinterpret_cast<uintptr_t>(p.loc) & 7u) == 0) &&
                       // 32-bit element offsets inside one batch slice and 32-bit item indices
                       p.lv.start[p.lv.n - 1] + static_cast<int64_t>(p.lv.h[p.lv.n - 1]) * p.lv.w[p.lv.n - 1] <=
                           (INT32_MAX - 4096) / (p.value_pos_stride > 0 ? p.value_pos_stride : 1) &&
                       p.rows * p.n_heads * LP < INT32_MAX;
  if (fast_ok) {
    if (value_dtype == MOYOLO_BF16) {
      return head_dim == 32 ? launch_fast<__nv_bfloat16, 32, FUSED>(p, st)
                            : launch_fast<__nv_bfloat16, 64, FUSED>(p, st);
    }
    return head_dim == 32 ? launch_fast<float, 32, FUSED>(p, st) : launch_fast<float, 64, FUSED>(p, st);
  }
  const int64_t total = p.rows * p.n_heads * head_dim;
  if (total == 0) return MOYOLO_OK;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((total + 255) / 256, 148 * 16));
  if (value_dtype == MOYOLO_F64) {
    MOYOLO_REQUIRE(!FUSED && aux_dtype == MOYOLO_F64, MOYOLO_ERR_UNSUPPORTED,
                   "fp64 value needs fp64 loc/weights in pre-normalised mode");
    launch_k(msda_generic_kernel<double, double, false>, dim3(blocks), dim3(256), 0, st, p, head_dim);
  } else {
    MOYOLO_REQUIRE(aux_dtype == MOYOLO_F32, MOYOLO_ERR_UNSUPPORTED,
                   "loc/weights must be fp32 for fp32/bf16 value");
    if (value_dtype == MOYOLO_F32)
      launch_k(msda_generic_kernel<float, float, FUSED>, dim3(blocks), dim3(256), 0, st, p, head_dim);
    else
      launch_k(msda_generic_kernel<__nv_bfloat16, float, FUSED>, dim3(blocks), dim3(256), 0, st, p, head_dim);
  }
  return check_launch("msda_generic_kernel");
}

static int fill_common(MsdaParams* p, const void* value, int value_dtype, int64_t value_batch_stride,
                       int64_t value_pos_stride, const int32_t* shapes_hw_host, int n_levels, int batch,
                       int64_t len_v, int n_heads, int head_dim, int n_points, int64_t rows,
                       const int32_t* row_offsets, void* out, int64_t out_row_stride) {
  MOYOLO_REQUIRE(value && out && shapes_hw_host, MOYOLO_ERR_BAD_ARG, "null value/out/shapes pointer");
  MOYOLO_REQUIRE(value_dtype == MOYOLO_F32 || value_dtype == MOYOLO_BF16 || value_dtype == MOYOLO_F64,
                 MOYOLO_ERR_UNSUPPORTED, "unsupported value dtype %d", value_dtype);
  MOYOLO_REQUIRE(batch > 0 && n_heads > 0 && head_dim > 0 && n_points > 0 && rows >= 0,
                 MOYOLO_ERR_BAD_ARG, "batch/n_heads/head_dim/n_points must be positive");
  MOYOLO_REQUIRE(row_offsets != nullptr || rows % batch == 0, MOYOLO_ERR_BAD_SHAPE,
                 "dense rows (%lld) must be divisible by batch (%d)", (long long)rows, batch);
  MOYOLO_REQUIRE(out_row_stride >= static_cast<int64_t>(n_heads) * head_dim, MOYOLO_ERR_BAD_SHAPE,
                 "out_row_stride smaller than n_heads*head_dim");
  MOYOLO_REQUIRE(value_pos_stride >= static_cast<int64_t>(n_heads) * head_dim, MOYOLO_ERR_BAD_SHAPE,
                 "value_pos_stride smaller than n_heads*head_dim");
  int rc = make_levels(shapes_hw_host, n_levels, len_v, &p->lv);
  if (rc != MOYOLO_OK) return rc;
  p->value = value;
  p->value_batch_stride = value_batch_stride;
  p->value_pos_stride = value_pos_stride;
  p->value_head_stride = head_dim;
  p->batch = batch;
  p->n_heads = n_heads;
  p->n_points = n_points;
  p->rows = rows;
  p->rows_per_batch = rows / batch > 0 ? rows / batch : 1;
  p->row_offsets = row_offsets;
  p->out = out;
  p->out_row_stride = out_row_stride;
  return MOYOLO_OK;
}

}  // namespace moyolo

using namespace moyolo;

extern "C" int moyolo_msda_sampled_forward(const void* value, int value_dtype, int64_t value_batch_stride,
                                           int64_t value_pos_stride, const int32_t* shapes_hw_host,
                                           int n_levels, int batch, int64_t len_v, int n_heads,
                                           int head_dim, int n_points, const void* loc,
                                           const void* weights, int aux_dtype, int64_t rows,
                                           const int32_t* row_offsets, void* out, int64_t out_row_stride,
                                           moyolo_stream_t stream) {
  MsdaParams p{};
  int rc = fill_common(&p, value, value_dtype, value_batch_stride, value_pos_stride, shapes_hw_host,
                       n_levels, batch, len_v, n_heads, head_dim, n_points, rows, row_offsets, out,
                       out_row_stride);
  if (rc != MOYOLO_OK) return rc;
  MOYOLO_REQUIRE(loc && weights, MOYOLO_ERR_BAD_ARG, "null loc/weights pointer");
  p.loc = loc;
  p.weights = weights;
  return dispatch<false>(p, value_dtype, aux_dtype, head_dim, static_cast<cudaStream_t>(stream));
}

extern "C" int moyolo_msda_fused_forward(const void* value, int value_dtype, int64_t value_batch_stride,
                                         int64_t value_pos_stride, const int32_t* shapes_hw_host,
                                         int n_levels, int batch, int64_t len_v, int n_heads, int head_dim,
                                         int n_points, const float* offsets, int64_t offsets_row_stride,
                                         const float* logits, int64_t logits_row_stride, const float* refer,
                                         int ref_levels, int ref_dim, int softmax_mode, int64_t rows,
                                         const int32_t* row_offsets, void* out, int64_t out_row_stride,
                                         moyolo_stream_t stream) {
  MsdaParams p{};
  int rc = fill_common(&p, value, value_dtype, value_batch_stride, value_pos_stride, shapes_hw_host,
                       n_levels, batch, len_v, n_heads, head_dim, n_points, rows, row_offsets, out,
                       out_row_stride);
  if (rc != MOYOLO_OK) return rc;
  MOYOLO_REQUIRE(offsets && logits && refer, MOYOLO_ERR_BAD_ARG, "null offsets/logits/refer pointer");
  MOYOLO_REQUIRE(ref_dim == 2 || ref_dim == 4, MOYOLO_ERR_BAD_SHAPE,
                 "Last dim of reference_points must be 2 or 4, but got %d.", ref_dim);
  MOYOLO_REQUIRE(ref_levels == 1 || ref_levels == n_levels, MOYOLO_ERR_BAD_SHAPE,
                 "refer_bbox level dim must be 1 or n_levels (%d), got %d", n_levels, ref_levels);
  MOYOLO_REQUIRE(softmax_mode == MOYOLO_SOFTMAX || softmax_mode == MOYOLO_SOFTMAX_PLUS1,
                 MOYOLO_ERR_BAD_ARG, "bad softmax_mode %d", softmax_mode);
  MOYOLO_REQUIRE(value_dtype != MOYOLO_F64, MOYOLO_ERR_UNSUPPORTED, "fused mode is fp32/bf16 only");
  p.offsets = offsets;
  p.offsets_row_stride = offsets_row_stride;
  p.logits = logits;
  p.logits_row_stride = logits_row_stride;
  p.refer = refer;
  p.ref_levels = ref_levels;
  p.ref_dim = ref_dim;
  p.softmax_mode = softmax_mode;
  return dispatch<true>(p, value_dtype, MOYOLO_F32, head_dim, static_cast<cudaStream_t>(stream));
}

// Head-major value layout [B, H, Lv, head_dim] (value_head_stride = Lv * head_dim, value_pos_stride = head_dim): the
// x0 / x1 bilinear corners of a sampling point are then one contiguous 2 * head_dim segment. Experimental entry point
// (benchmarks/headmajor_probe.py): served by the two-items-per-warp kernel only.
extern "C" int moyolo_msda_fused_forward_headmajor(const void* value, int value_dtype, int64_t value_batch_stride,
                                                   int64_t value_head_stride, int64_t value_pos_stride,
                                                   const int32_t* shapes_hw_host, int n_levels, int batch, int64_t len_v,
                                                   int n_heads, int head_dim, int n_points, const float* offsets,
                                                   int64_t offsets_row_stride, const float* logits,
                                                   int64_t logits_row_stride, const float* refer, int ref_levels,
                                                   int ref_dim, int softmax_mode, int64_t rows,
                                                   const int32_t* row_offsets, void* out, int64_t out_row_stride,
                                                   moyolo_stream_t stream) {
  MsdaParams p{};
  MOYOLO_REQUIRE(value_dtype == MOYOLO_BF16 && n_heads == 8 && head_dim == 32 && n_points == 4 && n_levels == 3,
                 MOYOLO_ERR_UNSUPPORTED, "msda_fused_forward_headmajor serves bf16, 8 heads x 32, 3 levels x 4 points");
  int rc = fill_common(&p, value, value_dtype, value_batch_stride, static_cast<int64_t>(n_heads) * head_dim, shapes_hw_host,
                       n_levels, batch, len_v, n_heads, head_dim, n_points, rows, row_offsets, out, out_row_stride);
  if (rc != MOYOLO_OK) return rc;
  MOYOLO_REQUIRE(offsets && logits && refer && (ref_dim == 2 || ref_dim == 4) && (ref_levels == 1 || ref_levels == n_levels),
                 MOYOLO_ERR_BAD_ARG, "msda_fused_forward_headmajor: bad offsets/logits/refer arguments");
  MOYOLO_REQUIRE(aligned16(value) && aligned16(out) && (value_pos_stride * 2) % 16 == 0 && (value_head_stride * 2) % 16 == 0 &&
                     (value_batch_stride * 2) % 16 == 0 && (out_row_stride * 2) % 16 == 0 &&
                     (reinterpret_cast<uintptr_t>(offsets) & 7u) == 0 && offsets_row_stride % 2 == 0 &&
                     (reinterpret_cast<uintptr_t>(refer) & (ref_dim == 4 ? 15u : 7u)) == 0,
                 MOYOLO_ERR_ALIGNMENT, "msda_fused_forward_headmajor: operands must be 16-byte aligned");
  MOYOLO_REQUIRE(len_v * value_pos_stride < INT32_MAX - 4096, MOYOLO_ERR_UNSUPPORTED, "tensor too large for 32-bit offsets");
  p.value_pos_stride = value_pos_stride;
  p.value_head_stride = value_head_stride;
  p.offsets = offsets; p.offsets_row_stride = offsets_row_stride; p.logits = logits; p.logits_row_stride = logits_row_stride;
  p.refer = refer; p.ref_levels = ref_levels; p.ref_dim = ref_dim; p.softmax_mode = softmax_mode;
  if (rows == 0) return MOYOLO_OK;
  return launch_pair<__nv_bfloat16, 4, 3>(p, static_cast<cudaStream_t>(stream));
}

extern "C" int moyolo_msda_proj_fused_forward(const void* value, int value_dtype, int64_t value_batch_stride,
                                              int64_t value_pos_stride, const int32_t* shapes_hw_host, int n_levels,
                                              int batch, int64_t len_v, int n_heads, int head_dim, int n_points,
                                              const void* xq, int64_t xq_row_stride, const void* w_offlog,
                                              const float* b_offlog, const float* refer, int ref_levels, int ref_dim,
                                              int softmax_mode, int64_t rows, const int32_t* row_offsets, void* out,
                                              int64_t out_row_stride, moyolo_stream_t stream) {
  MsdaParams p{};
  int rc = fill_common(&p, value, value_dtype, value_batch_stride, value_pos_stride, shapes_hw_host,
                       n_levels, batch, len_v, n_heads, head_dim, n_points, rows, row_offsets, out,
                       out_row_stride);
  if (rc != MOYOLO_OK) return rc;
  MOYOLO_REQUIRE(xq && w_offlog && b_offlog && refer, MOYOLO_ERR_BAD_ARG, "null xq/w_offlog/b_offlog/refer pointer");
  MOYOLO_REQUIRE(value_dtype == MOYOLO_BF16 && n_heads == 8 && head_dim == 32 && n_points == 4 && n_levels == 3,
                 MOYOLO_ERR_UNSUPPORTED,
                 "msda_proj_fused_forward serves bf16, 8 heads x 32, 3 levels x 4 points (use moyolo_linear + "
                 "moyolo_msda_fused_forward otherwise)");
  MOYOLO_REQUIRE(ref_dim == 2 || ref_dim == 4, MOYOLO_ERR_BAD_SHAPE,
                 "Last dim of reference_points must be 2 or 4, but got %d.", ref_dim);
  MOYOLO_REQUIRE(ref_levels == 1 || ref_levels == n_levels, MOYOLO_ERR_BAD_SHAPE,
                 "refer_bbox level dim must be 1 or n_levels (%d), got %d", n_levels, ref_levels);
  MOYOLO_REQUIRE(softmax_mode == MOYOLO_SOFTMAX || softmax_mode == MOYOLO_SOFTMAX_PLUS1, MOYOLO_ERR_BAD_ARG,
                 "bad softmax_mode %d", softmax_mode);
  MOYOLO_REQUIRE(aligned16(value) && aligned16(out) && aligned16(xq) && aligned16(w_offlog) &&
                     (value_pos_stride * 2) % 16 == 0 && (value_batch_stride * 2) % 16 == 0 &&
                     (out_row_stride * 2) % 16 == 0 && (xq_row_stride * 2) % 16 == 0 &&
                     (reinterpret_cast<uintptr_t>(refer) & (ref_dim == 4 ? 15u : 7u)) == 0,
                 MOYOLO_ERR_ALIGNMENT, "msda_proj_fused_forward: operands must be 16-byte aligned");
  MOYOLO_REQUIRE(p.lv.start[p.lv.n - 1] + static_cast<int64_t>(p.lv.h[p.lv.n - 1]) * p.lv.w[p.lv.n - 1] <=
                         (INT32_MAX - 4096) / (p.value_pos_stride > 0 ? p.value_pos_stride : 1) &&
                     rows < (1 << 24),
                 MOYOLO_ERR_UNSUPPORTED, "msda_proj_fused_forward: tensor too large for 32-bit offsets");
  p.xq = xq; p.xq_row_stride = xq_row_stride; p.w_offlog = w_offlog; p.b_offlog = b_offlog;
  p.refer = refer; p.ref_levels = ref_levels; p.ref_dim = ref_dim; p.softmax_mode = softmax_mode;
  if (rows == 0) return MOYOLO_OK;
  constexpr int K = 256, PITCH = K + 8, NOp = 40, NUp = 48;
  constexpr size_t smem = static_cast<size_t>(NOp + kProjRows) * PITCH * 2 + NOp * 4 + kProjRows * NOp * 4 +
                          static_cast<size_t>(kGatherWarps) * kProjRows * NOp * 4 + static_cast<size_t>(kGatherWarps) * NUp * 8;
  static_assert(smem <= 48 * 1024, "fits the default dynamic shared memory limit");
  const dim3 grid(static_cast<unsigned>((rows + kProjRows - 1) / kProjRows), 8);
  launch_k(msda_gather_proj_kernel<4, 3>, grid, dim3(kGatherThreads), smem, static_cast<cudaStream_t>(stream), p);
  return check_launch("msda_gather_proj_kernel");
}
