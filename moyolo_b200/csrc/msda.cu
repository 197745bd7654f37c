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
#include <algorithm>

#include "common.cuh"

namespace moyolo {

struct MsdaParams {
  const void* value;
  int64_t value_batch_stride;  // elements
  int64_t value_pos_stride;    // elements
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

// U = corner rounds kept in flight per loop trip (all of them when 4*L*P == U*32/G).
template <typename VT, int DH, bool FUSED, int U>
__global__ void __launch_bounds__(kGatherThreads) msda_gather_kernel(const MsdaParams p) {
  pdl_trigger();
  pdl_wait();
  constexpr int G = DH * static_cast<int>(sizeof(VT)) / 16;  // lanes per value row
  constexpr int CH = 16 / static_cast<int>(sizeof(VT));      // channels per lane
  constexpr int NSG = 32 / G;                                // sub-groups per warp
  static_assert(G >= 4 && G <= 16 && CH >= NSG, "head row must be 64..256 bytes");

  extern __shared__ __align__(16) int smem_i[];
  const int LP = p.lv.n * p.n_points;
  const int NU = LP * 4;
  const int NUp = (NU + NSG * U - 1) / (NSG * U) * (NSG * U);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = static_cast<int64_t>(blockIdx.x) * kGatherWarps + warp;
  if (item >= p.rows * p.n_heads) return;  // warp-uniform
  const int64_t row = item / p.n_heads;
  const int head = static_cast<int>(item % p.n_heads);

  int* s_pos = smem_i + warp * NUp * 2;
  float* s_w = reinterpret_cast<float*>(s_pos + NUp);

  // ---------------- phase 1: lane `l` owns sampling points l and l + 32 ----------------
  float aw[kMaxPointsPerLane];
  if (FUSED) {
    float lg[kMaxPointsPerLane];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < kMaxPointsPerLane; ++i) {
      const int pt = lane + i * 32;
      lg[i] = pt < LP ? __ldg(p.logits + row * p.logits_row_stride + head * LP + pt) : -INFINITY;
      m = fmaxf(m, lg[i]);
    }
    m = warp_max(m);
    if (p.softmax_mode == MOYOLO_SOFTMAX_PLUS1) m = 0.0f;  // exp(x)/(1+sum exp(x)), no shift
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < kMaxPointsPerLane; ++i) {
      const int pt = lane + i * 32;
      aw[i] = pt < LP ? expf(lg[i] - m) : 0.0f;
      s += aw[i];
    }
    s = warp_sum(s);
    if (p.softmax_mode == MOYOLO_SOFTMAX_PLUS1) s += 1.0f;
    const float inv = 1.0f / s;
#pragma unroll
    for (int i = 0; i < kMaxPointsPerLane; ++i) aw[i] *= inv;
  }
#pragma unroll
  for (int i = 0; i < kMaxPointsPerLane; ++i) {
    const int pt = lane + i * 32;
    if (pt < LP) {
      const int level = pt / p.n_points;
      float lx, ly, a;
      if (FUSED) {
        fused_location(p, row, head, pt, level, &lx, &ly);
        a = aw[i];
      } else {
        const int64_t idx = (row * p.n_heads + head) * LP + pt;
        const float2 l2 = __ldg(reinterpret_cast<const float2*>(p.loc) + idx);
        lx = l2.x;
        ly = l2.y;
        a = __ldg(static_cast<const float*>(p.weights) + idx);
      }
      const Corners c = make_corners(lx, ly, p.lv.h[level], p.lv.w[level], p.lv.start[level], a);
      *reinterpret_cast<int4*>(s_pos + pt * 4) = make_int4(c.pos[0], c.pos[1], c.pos[2], c.pos[3]);
      *reinterpret_cast<float4*>(s_w + pt * 4) = make_float4(c.w[0], c.w[1], c.w[2], c.w[3]);
    }
  }
  for (int u = NU + lane; u < NUp; u += 32) { s_pos[u] = -1; s_w[u] = 0.0f; }
  __syncwarp();

  // ---------------- phase 2: gather, sub-group sg takes corners sg, sg + NSG, ... ----------------
  float acc[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) acc[k] = 0.0f;

  const int sg = lane / G, sub = lane % G;
  const int b = batch_of_row(row, p.row_offsets, p.batch, p.rows_per_batch);
  const VT* base = static_cast<const VT*>(p.value) + static_cast<int64_t>(b) * p.value_batch_stride +
                   head * DH + sub * CH;
  const int64_t ps = p.value_pos_stride;

  for (int u0 = sg; u0 < NUp; u0 += NSG * U) {
    uint4 v[U];
    float w[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int pos = s_pos[u0 + i * NSG];
      w[i] = s_w[u0 + i * NSG];
      v[i] = (pos >= 0) ? ldg128(base + pos * ps) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if constexpr (sizeof(VT) == 2) {
        const float2 a0 = bf16x2_to_float2(v[i].x), a1 = bf16x2_to_float2(v[i].y);
        const float2 a2 = bf16x2_to_float2(v[i].z), a3 = bf16x2_to_float2(v[i].w);
        acc[0] = fmaf(w[i], a0.x, acc[0]);
        acc[1] = fmaf(w[i], a0.y, acc[1]);
        acc[2] = fmaf(w[i], a1.x, acc[2]);
        acc[3] = fmaf(w[i], a1.y, acc[3]);
        acc[4] = fmaf(w[i], a2.x, acc[4]);
        acc[5] = fmaf(w[i], a2.y, acc[5]);
        acc[6] = fmaf(w[i], a3.x, acc[6]);
        acc[7] = fmaf(w[i], a3.y, acc[7]);
      } else {
        acc[0] = fmaf(w[i], __uint_as_float(v[i].x), acc[0]);
        acc[1] = fmaf(w[i], __uint_as_float(v[i].y), acc[1]);
        acc[2] = fmaf(w[i], __uint_as_float(v[i].z), acc[2]);
        acc[3] = fmaf(w[i], __uint_as_float(v[i].w), acc[3]);
      }
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
  VT* o = static_cast<VT*>(p.out) + row * p.out_row_stride + head * DH + sub * CH + ch;
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


template <typename VT, int DH, bool FUSED, int U>
static int launch_fast_u(const MsdaParams& p, cudaStream_t st) {
  constexpr int G = DH * static_cast<int>(sizeof(VT)) / 16;
  constexpr int NSG = 32 / G;
  const int NU = p.lv.n * p.n_points * 4;
  const int NUp = (NU + NSG * U - 1) / (NSG * U) * (NSG * U);
  const size_t smem = static_cast<size_t>(kGatherWarps) * NUp * 8;
  const int64_t items = p.rows * p.n_heads;
  const int64_t blocks = (items + kGatherWarps - 1) / kGatherWarps;
  if (blocks == 0) return MOYOLO_OK;
  launch_k(msda_gather_kernel<VT, DH, FUSED, U>, dim3(static_cast<unsigned>(blocks)), dim3(kGatherThreads), smem, st, p);
  return check_launch("msda_gather_kernel");
}

template <typename VT, int DH, bool FUSED>
static int launch_fast(const MsdaParams& p, cudaStream_t st) {
  constexpr int NSG = 32 / (DH * static_cast<int>(sizeof(VT)) / 16);
  const int rounds = (p.lv.n * p.n_points * 4 + NSG - 1) / NSG;  // corner rounds per sub-group
  // keep every round in flight when there are few (6 at L=3,P=4 bf16); otherwise trips of 8
  if (rounds % 6 == 0 && rounds <= 12) return launch_fast_u<VT, DH, FUSED, 6>(p, st);
  return launch_fast_u<VT, DH, FUSED, 8>(p, st);
}

template <bool FUSED>
static int dispatch(const MsdaParams& p, int value_dtype, int aux_dtype, int head_dim, cudaStream_t st) {
  const int LP = p.lv.n * p.n_points;
  const int esz = value_dtype == MOYOLO_BF16 ? 2 : (value_dtype == MOYOLO_F32 ? 4 : 8);
  const bool fast_ok = (value_dtype != MOYOLO_F64) && (aux_dtype == MOYOLO_F32) &&
                       (head_dim == 32 || head_dim == 64) && aligned16(p.value) && aligned16(p.out) &&
                       (p.value_pos_stride * esz) % 16 == 0 && (p.value_batch_stride * esz) % 16 == 0 &&
                       (p.out_row_stride * esz) % 16 == 0 && LP <= 32 * kMaxPointsPerLane &&
                       (FUSED || (reinterpret_cast<uintptr_t>(p.loc) & 7u) == 0);
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
