// Row-tile-persistent decoder: all layers of the MOTR / deformable decoder in ONE launch.
//
// Replaces, for one frame: MOTRTransformerDecoder.forward (ultralytics/nn/modules/transformer.py:676-728) = 6 x
// MOTRDecoderLayer.forward (:627-652: nn.MultiheadAttention self-attention + LayerNorm, MSDeformAttn :246-287 +
// LayerNorm, FFN :576-580 + LayerNorm) + the per-layer box refinement sigmoid(bbox_head(x) + inverse_sigmoid(ref))
// (:709) + the class-score head of the last layer (:717-721). The 7 + 3 launches per layer of the launch-chained
// schedule (executor.run_layer_ws) become one kernel for the whole decoder; the value projection stays a separate
// tcgen05 GEMM (it depends on the frame's feature maps only and runs ahead of the frame).
//
// Decomposition. A thread-block CLUSTER of 8 CTAs owns a tile of M = 32 (or 64) query rows through every row-local
// operation of every layer; CTA rank r of the cluster owns
//   * attention head r (self-attention and deformable gather are per head), and
//   * the 32-column slab r of every 256-wide GEMM output (128-column slab of the FFN hidden layer).
// The activation tile [M, 256] (bf16 GEMM operand) is REPLICATED in the shared memory of all 8 CTAs; every stage
// computes its slab and writes it into the 8 replicas through distributed shared memory (st.shared::cluster),
// followed by one cluster barrier. LayerNorm statistics (mean, centred sum of squares per slab) are exchanged the
// same way and merged with the parallel-variance formula. The FFN's second GEMM is split along K (each CTA
// multiplies its own 128 hidden columns: no exchange of the hidden activations) and reduce-scattered through
// DSMEM. Weights are never staged in shared memory: each warp loads its mma.sync B fragments straight from
// global memory / L2 with 128-bit read-only loads (8 consecutive k per lane, a k-permutation that the A
// fragments mirror), each weight byte is read once per cluster from L2.
// The only data another cluster needs are the keys / values of the self-attention: they go through global
// memory and a grid-wide barrier per layer (release/acquire on a global counter; all clusters are co-resident,
// the host checks the occupancy). Tensor work is mma.sync m16n8k16 (bf16 in, fp32 accumulate): the tiles are
// 32 rows x 8..32 columns per warp task, far below a 128-row tcgen05 tile, and the kernel is latency-bound.
//
// Numerics follow the launch-chained bf16 path: bf16 GEMM operands and value tensor; fp32 residual stream,
// LayerNorm, softmax, sampling locations, accumulation; class scores from the bf16-rounded output row.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "gather_common.cuh"

namespace moyolo {
namespace dc {

constexpr int kThreads = 256, kWarps = 8, kCluster = 8;
constexpr int kC = 256, kDh = 32;
constexpr int kFs = 128;  // FFN hidden columns per CTA (d_ffn 1024 / 8)
constexpr int kPA = 288;  // bf16 pitch of the activation tiles: 576 B = 64 (mod 128) -> conflict-free 128-bit A loads
constexpr int kPH = 160;  // bf16 pitch of the FFN hidden slab [M][128]: 320 B = 64 (mod 128)
constexpr int kPQ = 40;   // bf16 pitch of Q / K / V head rows (ldmatrix layout of attention.cu)
constexpr int kPY = 33;   // fp32 pitch of the pre-LayerNorm slab
constexpr int kPO = 40;   // fp32 pitch of one row's offsets | logits (36 used)
constexpr int kNP = 4, kNL = 3, kLP = kNP * kNL;  // sampling points x levels (the model's configuration)
constexpr int kNU = kLP * 4;                       // bilinear corners per (row, head)
constexpr int kMaxLayers = 8;
constexpr int kMaxScoreNc = 8;

struct LayerW {
  const __nv_bfloat16 *wqkv, *wo, *woff, *wout, *w1, *w2, *wb1, *wb2;
  const float *bqkv, *bo, *boff, *bout, *b1, *b2, *bb1, *bb2, *wb3, *bb3;
  const float *g1, *be1, *g2, *be2, *g3, *be3;
};

struct Params {
  LayerW L[kMaxLayers];
  int n_layers;
  const float* x_in;      // [rows_pad, 256] residual stream in (frame_assemble)
  const float* pos;       // [rows_pad, 256] query_pos (fixed for all layers, transformer.py:705-707)
  const float* refer0;    // [rows_pad, 4] sigmoid(refer) (transformer.py:690)
  float* x_out;           // [rows_pad, 256] last layer's output embedding (fp32)
  __nv_bfloat16* x_lp_out;  // optional bf16 copy
  float* refer_out[kMaxLayers];  // [rows_pad, 4] refined boxes after layer i (any may be NULL)
  __nv_bfloat16* kv;      // [2, rows_pad, 512] scratch: K | V of the self-attention, double-buffered by layer parity
                          // (a cluster that is one layer ahead must not overwrite keys another one still reads)
  const __nv_bfloat16* values;  // [B, Lv, n_layers*256] all layers' value projections
  int64_t v_batch_stride, v_pos_stride;
  LevelTable lv;
  int softmax_mode;
  const int32_t* ro;      // [n_seq + 1] row offsets (device)
  int n_seq, rows_pad;
  unsigned* grid_bar;     // zeroed before the launch
  int* status;            // optional: set to 1 when the tiles do not fit the launched clusters / key staging
  const float *score_w, *score_b;
  int nc;
  float* logits;
  float* scores;
  int32_t* labels;
  float eps;
  int kv_cap;             // keys (rounded to 32) that fit the K/V staging area
  long long* profile;     // optional [n_layers][16] globaltimer stamps of cluster 0 / rank 0 (benchmarks/dc_stages.py)
};

// ---------------------------------------------------------------------------------------------------------
// cluster / DSMEM / grid-barrier primitives
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_peer(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_peer_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_peer_v2f(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Every CTA of the cluster has finished its global writes of this phase (cluster_sync), then ONE thread of the
// cluster publishes them at gpu scope and counts the cluster in.
__device__ __forceinline__ void grid_arrive(unsigned* bar, uint32_t rank) {
  cluster_sync();
  if (rank == 0 && threadIdx.x == 0) {
    __threadfence();
    red_release_gpu_add(bar, 1u);
  }
}
__device__ __forceinline__ void grid_wait(const unsigned* bar, unsigned target, uint32_t rank) {
  if (rank == 0 && threadIdx.x == 0) {
    while (ld_acquire_gpu(bar) < target) __nanosleep(32);
    __threadfence();
  }
  cluster_sync();
}

// ---------------------------------------------------------------------------------------------------------
// mma.sync building blocks. B fragments come straight from global memory: lane (g = lane/4, t = lane%4) loads
// the 8 consecutive k  [32 kb + 8 t, +8)  of weight row n0 + g with ONE 128-bit load per 32-wide k block; the
// two m16n8k16 steps of a block then use the k-permutation  virtual k (2t, 2t+1 | 2t+8, 2t+9)  <->  real k
// (8t, 8t+1 | 8t+2, 8t+3)  resp. (8t+4, 8t+5 | 8t+6, 8t+7), which the A fragments (128-bit shared-memory loads of
// the same 8 k of rows g and g + 8) mirror. The sum over k does not care about the order.
// ---------------------------------------------------------------------------------------------------------
template <int KB>
__device__ __forceinline__ void load_b(uint4 (&b)[KB], const __nv_bfloat16* lane_row_ptr, bool valid) {
#pragma unroll
  for (int kb = 0; kb < KB; ++kb) b[kb] = valid ? ldg128(lane_row_ptr + kb * 32) : make_uint4(0u, 0u, 0u, 0u);
}

// acc[i] += A[m-tile i] . B^T for NM m-tiles that share the B fragments. a = &A[first row of m-tile 0][k0].
template <int KB, int NM>
__device__ __forceinline__ void mma_tiles(float (&acc)[NM][4], const __nv_bfloat16* a, int pitch, const uint4 (&b)[KB],
                                          int lane) {
  const __nv_bfloat16* ap = a + (lane >> 2) * pitch + (lane & 3) * 8;
#pragma unroll
  for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
    for (int i = 0; i < NM; ++i) {
      const uint4 lo = *reinterpret_cast<const uint4*>(ap + (i * 16) * pitch + kb * 32);
      const uint4 hi = *reinterpret_cast<const uint4*>(ap + (i * 16 + 8) * pitch + kb * 32);
      const uint32_t a0[4] = {lo.x, hi.x, lo.y, hi.y};
      mma_bf16_16816(acc[i], a0, b[kb].x, b[kb].y);
      const uint32_t a1[4] = {lo.z, hi.z, lo.w, hi.w};
      mma_bf16_16816(acc[i], a1, b[kb].z, b[kb].w);
    }
  }
}

__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float inverse_sigmoidf_(float x) {  // utils.py:34-38
  x = fminf(fmaxf(x, 0.0f), 1.0f);
  return logf(fmaxf(x, 1e-5f) / fmaxf(1.0f - x, 1e-5f));
}

// ---------------------------------------------------------------------------------------------------------
// Shared-memory plan of one CTA
// ---------------------------------------------------------------------------------------------------------
template <int MT>
struct Smem {
  static constexpr int M = 16 * MT;
  static constexpr size_t oA0 = 0;
  static constexpr size_t oA1 = oA0 + size_t(M) * kPA * 2;
  static constexpr size_t oH = oA1 + size_t(M) * kPA * 2;
  static constexpr size_t oY = oH + size_t(M) * kPH * 2;
  static constexpr size_t oRes = oY + size_t(M) * kPY * 4;
  static constexpr size_t oPos = oRes + size_t(M) * 32 * 4;
  static constexpr size_t oSlab0 = oPos + size_t(M) * 32 * 4;
  static constexpr size_t oSlab1 = oSlab0 + size_t(M) * 32 * 2;
  static constexpr size_t oStat = oSlab1 + size_t(M) * 32 * 2;
  static constexpr size_t oRef = oStat + size_t(kCluster) * M * 8;
  static constexpr size_t oOL = oRef + size_t(M) * 16;
  static constexpr size_t oQ = oOL + size_t(M) * kPO * 4;
  static constexpr size_t oStage = oQ + size_t(M) * kPQ * 2;
  static constexpr size_t oMl = oStage + size_t(kWarps) * 2 * kNU * 8;
  static constexpr size_t oScratch = ((oMl + size_t(kWarps) * 16 * 2 * 4) + 127) / 128 * 128;
  // scratch: attention = partial O [8 warps][16][32] fp32, then K [kv_cap][kPQ], V [kv_cap][kPQ];
  //          FFN2      = split-K partial sums [8 sources][M][32] fp32 (aliases the attention area)
  static constexpr size_t kPoBytes = size_t(kWarps) * 16 * 32 * 4;
  static constexpr size_t kPartBytes = size_t(kCluster) * M * 32 * 4;
  static constexpr size_t kTotalMax = 232448 - 1024;  // 227 KiB minus alignment slack
  static constexpr int kv_cap() {
    return static_cast<int>((kTotalMax - oScratch - kPoBytes) / (2 * kPQ * 2)) / 32 * 32;
  }
  static constexpr size_t total() {
    const size_t att = kPoBytes + size_t(kv_cap()) * 2 * kPQ * 2;
    return oScratch + (att > kPartBytes ? att : kPartBytes);
  }
};

struct Tile {
  int seq, row0, n, seq_start, seq_len;
};

// tile `idx` of the frame: tiles are enumerated per sequence (a tile never spans two sequences).
template <int M>
__device__ __forceinline__ bool find_tile(const Params& p, int idx, Tile* t, int* total_tiles, int* max_len) {
  int acc = 0, ml = 0;
  bool found = false;
  for (int s = 0; s < p.n_seq; ++s) {
    const int a = p.ro[s], len = p.ro[s + 1] - a;
    const int nt = (len + M - 1) / M;
    if (!found && idx < acc + nt) {
      const int j = idx - acc;
      t->seq = s;
      t->seq_start = a;
      t->seq_len = len;
      t->row0 = a + j * M;
      t->n = min(M, len - j * M);
      found = true;
    }
    acc += nt;
    ml = max(ml, len);
  }
  *total_tiles = acc;
  *max_len = ml;
  return found;
}

// Write this CTA's bf16 slab [M][32] into columns [col0, col0 + 32) of tile `dst` in ALL 8 CTAs of the cluster
// (warp w -> peer w, 16-byte chunks).
template <int M>
__device__ __forceinline__ void broadcast_slab(const __nv_bfloat16* slab, __nv_bfloat16* dst, int col0, int warp, int lane) {
  const uint32_t base = map_peer(smem_addr(dst), static_cast<uint32_t>(warp));
  for (int c = lane; c < M * 4; c += 32) {
    const int row = c >> 2, part = c & 3;
    const uint4 v = *reinterpret_cast<const uint4*>(slab + row * 32 + part * 8);
    st_peer_v4(base + static_cast<uint32_t>((row * kPA + col0 + part * 8) * 2), v);
  }
}

// LayerNorm of rows whose 256 columns are spread over the 8 CTAs (32 each). In: sY[row][lane] = pre-norm value of
// this CTA's slab (all M rows written, block-synchronised). Exchanges (mean, centred sum of squares) of the slab
// with all peers, merges the eight slabs (parallel-variance formula), and calls emit(row, lane, normalised value).
template <int MT, typename Emit>
__device__ __forceinline__ void cluster_layernorm(const float* sY, float* sStat, const float* __restrict__ gamma,
                                                  const float* __restrict__ beta, float eps, uint32_t rank, int warp,
                                                  int lane, Emit emit) {
  constexpr int M = 16 * MT, RPW = M / kWarps;
  float v[RPW];
  const uint32_t stat_peer = map_peer(smem_addr(sStat), static_cast<uint32_t>(lane & 7));
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int row = warp + i * kWarps;
    v[i] = sY[row * kPY + lane];
    const float mean = warp_sum(v[i]) * (1.0f / 32.0f);
    const float d = v[i] - mean;
    const float m2 = warp_sum(d * d);
    if (lane < kCluster) st_peer_v2f(stat_peer + static_cast<uint32_t>((rank * M + row) * 8), mean, m2);
  }
  const float g = __ldg(gamma + lane), b = __ldg(beta + lane);
  cluster_sync();
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int row = warp + i * kWarps;
    const float2 st = *reinterpret_cast<const float2*>(sStat + ((lane & 7) * M + row) * 2);
    float ms = st.x;
    ms += __shfl_xor_sync(0xffffffffu, ms, 1);
    ms += __shfl_xor_sync(0xffffffffu, ms, 2);
    ms += __shfl_xor_sync(0xffffffffu, ms, 4);
    const float mean = ms * (1.0f / kCluster);
    const float dm = st.x - mean;
    float m2 = st.y + 32.0f * dm * dm;
    m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
    m2 += __shfl_xor_sync(0xffffffffu, m2, 2);
    m2 += __shfl_xor_sync(0xffffffffu, m2, 4);
    const float rstd = rsqrtf(m2 * (1.0f / kC) + eps);
    emit(row, lane, (v[i] - mean) * rstd * g + b);
  }
}

// One 32-column-slab GEMM over K = 256: y[row][col] = A[row] . W[n0 + col] + bias[n0 + col] (+ residual) -> sY.
// Warp task = (n8-tile w & 3, m-group w >> 2) with MT / 2 m-tiles per group.
template <int MT, bool RELU_TO_SLAB>
__device__ __forceinline__ void gemm_slab32(const __nv_bfloat16* sA, const __nv_bfloat16* __restrict__ w, int n0,
                                            const float* __restrict__ bias, const float* sRes, float* sY,
                                            __nv_bfloat16* slab, int warp, int lane) {
  constexpr int NM = MT / 2;
  const int j = warp & 3, mg = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  uint4 b[8];
  load_b<8>(b, w + static_cast<int64_t>(n0 + j * 8 + g) * kC + t * 8, true);
  float acc[NM][4];
#pragma unroll
  for (int i = 0; i < NM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
  mma_tiles<8, NM>(acc, sA + (mg * NM * 16) * kPA, kPA, b, lane);
  const int col = j * 8 + 2 * t;
  const float b0 = __ldg(bias + n0 + col), b1 = __ldg(bias + n0 + col + 1);
#pragma unroll
  for (int i = 0; i < NM; ++i) {
    const int r0 = (mg * NM + i) * 16 + g, r1 = r0 + 8;
    if (RELU_TO_SLAB) {
      *reinterpret_cast<uint32_t*>(slab + r0 * 32 + col) = float2_to_bf16x2(fmaxf(acc[i][0] + b0, 0.0f), fmaxf(acc[i][1] + b1, 0.0f));
      *reinterpret_cast<uint32_t*>(slab + r1 * 32 + col) = float2_to_bf16x2(fmaxf(acc[i][2] + b0, 0.0f), fmaxf(acc[i][3] + b1, 0.0f));
    } else {
      sY[r0 * kPY + col] = acc[i][0] + b0 + sRes[r0 * 32 + col];
      sY[r0 * kPY + col + 1] = acc[i][1] + b1 + sRes[r0 * 32 + col + 1];
      sY[r1 * kPY + col] = acc[i][2] + b0 + sRes[r1 * 32 + col];
      sY[r1 * kPY + col + 1] = acc[i][3] + b1 + sRes[r1 * 32 + col + 1];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------------------
template <int MT>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
    decoder_cluster_kernel(const __grid_constant__ Params p) {
  using S = Smem<MT>;
  constexpr int M = S::M;
  extern __shared__ __align__(128) uint8_t dc_smem[];
  __nv_bfloat16* sA0 = reinterpret_cast<__nv_bfloat16*>(dc_smem + S::oA0);
  __nv_bfloat16* sA1 = reinterpret_cast<__nv_bfloat16*>(dc_smem + S::oA1);
  __nv_bfloat16* sH = reinterpret_cast<__nv_bfloat16*>(dc_smem + S::oH);
  float* sY = reinterpret_cast<float*>(dc_smem + S::oY);
  float* sRes = reinterpret_cast<float*>(dc_smem + S::oRes);
  float* sPos = reinterpret_cast<float*>(dc_smem + S::oPos);
  __nv_bfloat16* sSlab0 = reinterpret_cast<__nv_bfloat16*>(dc_smem + S::oSlab0);
  __nv_bfloat16* sSlab1 = reinterpret_cast<__nv_bfloat16*>(dc_smem + S::oSlab1);
  float* sStat = reinterpret_cast<float*>(dc_smem + S::oStat);
  float* sRef = reinterpret_cast<float*>(dc_smem + S::oRef);
  float* sOL = reinterpret_cast<float*>(dc_smem + S::oOL);
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(dc_smem + S::oQ);
  int* sStage = reinterpret_cast<int*>(dc_smem + S::oStage);
  float* sMl = reinterpret_cast<float*>(dc_smem + S::oMl);
  float* sPO = reinterpret_cast<float*>(dc_smem + S::oScratch);
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(dc_smem + S::oScratch + S::kPoBytes);
  float* sPart = reinterpret_cast<float*>(dc_smem + S::oScratch);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t rank = cluster_rank();
  const int cl = static_cast<int>(blockIdx.x) / kCluster;
  const int n0 = static_cast<int>(rank) * 32;  // this CTA's column slab / head
  pdl_wait();

  Tile tile;
  int total_tiles, max_len;
  const bool have = find_tile<M>(p, cl, &tile, &total_tiles, &max_len);
  const int n_clusters = static_cast<int>(gridDim.x) / kCluster;
  if (total_tiles > n_clusters || max_len > p.kv_cap) {  // cannot run: the host sized the launch wrongly
    if (p.status != nullptr && blockIdx.x == 0 && tid == 0) *p.status = 1;
    return;
  }
  if (!have) return;  // whole cluster idle (uniform over its 8 CTAs)
  __nv_bfloat16* sV = sK + static_cast<size_t>(p.kv_cap) * kPQ;
  const unsigned n_tiles_u = static_cast<unsigned>(total_tiles);

  // ---------------- prologue: tile of x / pos / refer, operands of the first in-projection ----------------
  for (int i = tid; i < M * (kC / 4); i += kThreads) {
    const int row = i / (kC / 4), c4 = (i % (kC / 4)) * 4;
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), pv = xv;
    if (row < tile.n) {
      xv = *reinterpret_cast<const float4*>(p.x_in + static_cast<int64_t>(tile.row0 + row) * kC + c4);
      pv = *reinterpret_cast<const float4*>(p.pos + static_cast<int64_t>(tile.row0 + row) * kC + c4);
    }
    uint2 a, b;
    a.x = float2_to_bf16x2(xv.x, xv.y);
    a.y = float2_to_bf16x2(xv.z, xv.w);
    b.x = float2_to_bf16x2(xv.x + pv.x, xv.y + pv.y);
    b.y = float2_to_bf16x2(xv.z + pv.z, xv.w + pv.w);
    *reinterpret_cast<uint2*>(sA0 + row * kPA + c4) = a;
    *reinterpret_cast<uint2*>(sA1 + row * kPA + c4) = b;
    if (c4 >= n0 && c4 < n0 + 32) {
      *reinterpret_cast<float4*>(sRes + row * 32 + (c4 - n0)) = xv;
      *reinterpret_cast<float4*>(sPos + row * 32 + (c4 - n0)) = pv;
    }
  }
  for (int i = tid; i < M * 4; i += kThreads) {
    const int row = i >> 2;
    sRef[i] = row < tile.n ? p.refer0[static_cast<int64_t>(tile.row0 + row) * 4 + (i & 3)] : 0.5f;
  }
  __syncthreads();

  // q, k (from x + pos) and v (from x) of THIS head for the tile's rows (transformer.py:637-638): q stays in
  // shared memory, k / v go to global memory for the other clusters.
  auto qkv_proj = [&](const LayerW& W, int for_layer) {
    __nv_bfloat16* kv_w = p.kv + static_cast<size_t>(for_layer & 1) * p.rows_pad * (2 * kC);
    for (int task = warp; task < 12 * MT; task += kWarps) {
      const int nt = task % 12, mt = task / 12;
      const int sec = nt >> 2, j = nt & 3;  // section 0 = q, 1 = k, 2 = v
      const int wrow = sec * kC + n0 + j * 8;
      uint4 b[8];
      load_b<8>(b, W.wqkv + static_cast<int64_t>(wrow + g) * kC + t * 8, true);
      float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
      mma_tiles<8, 1>(acc, (sec == 2 ? sA0 : sA1) + (mt * 16) * kPA, kPA, b, lane);
      const int col = j * 8 + 2 * t;
      const float b0 = __ldg(W.bqkv + wrow + 2 * t), b1 = __ldg(W.bqkv + wrow + 2 * t + 1);
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      const uint32_t v0 = float2_to_bf16x2(acc[0][0] + b0, acc[0][1] + b1);
      const uint32_t v1 = float2_to_bf16x2(acc[0][2] + b0, acc[0][3] + b1);
      if (sec == 0) {
        *reinterpret_cast<uint32_t*>(sQ + r0 * kPQ + col) = v0;
        *reinterpret_cast<uint32_t*>(sQ + r1 * kPQ + col) = v1;
      } else {
        __nv_bfloat16* dst = kv_w + (sec == 2 ? kC : 0) + n0 + col;
        if (r0 < tile.n) *reinterpret_cast<uint32_t*>(dst + static_cast<int64_t>(tile.row0 + r0) * (2 * kC)) = v0;
        if (r1 < tile.n) *reinterpret_cast<uint32_t*>(dst + static_cast<int64_t>(tile.row0 + r1) * (2 * kC)) = v1;
      }
    }
  };
  qkv_proj(p.L[0], 0);
  grid_arrive(p.grid_bar, rank);

  const float sl2 = rsqrtf(static_cast<float>(kDh)) * 1.4426950408889634f;  // softmax in base 2 (attention.cu)
  const int ps = static_cast<int>(p.v_pos_stride);

  const bool prof = p.profile != nullptr && blockIdx.x == 0 && tid == 0;
  int mark_i = 0;
  auto mark = [&](int layer) {
    if (prof && mark_i < 16) p.profile[layer * 16 + mark_i] = globaltimer_ns();
    ++mark_i;
  };
  for (int layer = 0; layer < p.n_layers; ++layer) {
    const LayerW& W = p.L[layer];
    const bool last = layer + 1 == p.n_layers;
    mark_i = 0;
    mark(layer);   // 0: layer start (before the grid barrier)
    grid_wait(p.grid_bar, n_tiles_u * static_cast<unsigned>(layer + 1), rank);
    if (last) pdl_trigger();
    mark(layer);   // 1: grid barrier passed

    // ======================= self-attention of head `rank` for the tile's M queries =======================
    {
      const int n_kt = (tile.seq_len + 31) / 32;
      const __nv_bfloat16* kg = p.kv + static_cast<size_t>(layer & 1) * p.rows_pad * (2 * kC) +
                                static_cast<int64_t>(tile.seq_start) * (2 * kC) + n0;
      for (int i = tid; i < n_kt * 32 * 4; i += kThreads) {
        const int key = i >> 2, c = i & 3;
        const bool ok = key < tile.seq_len;
        const __nv_bfloat16* src = kg + static_cast<int64_t>(ok ? key : 0) * (2 * kC) + c * 8;
        cp_async16(smem_addr(sK + key * kPQ + c * 8), src, ok);
        cp_async16(smem_addr(sV + key * kPQ + c * 8), src + kC, ok);
      }
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
      mark(layer);   // 2: K/V staged
      constexpr int NS = kWarps / MT;  // key splits
      const int mt = warp % MT, split = warp / MT;
      uint32_t qa[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
        ldmatrix_x4(smem_addr(sQ + (mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kPQ + ks * 16 + (lane >> 4) * 8), qa[ks]);
      float o[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.0f;
      float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
      for (int kt = split; kt < n_kt; kt += NS) {
        const __nv_bfloat16* kt_k = sK + kt * 32 * kPQ;
        const __nv_bfloat16* kt_v = sV + kt * 32 * kPQ;
        float s[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
          uint32_t kb[4];
          ldmatrix_x4(smem_addr(kt_k + (j * 8 + (lane & 7)) * kPQ + (lane >> 3) * 8), kb);
          mma_bf16_16816(s[j], qa[0], kb[0], kb[1]);
          mma_bf16_16816(s[j], qa[1], kb[2], kb[3]);
        }
        const int kbase = kt * 32;
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int key = kbase + j * 8 + 2 * t;
          if (key >= tile.seq_len) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
          if (key + 1 >= tile.seq_len) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
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
        uint32_t pa[2][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
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
        for (int n = 0; n < 4; ++n) { o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1; }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
          for (int np = 0; np < 2; ++np) {
            uint32_t vb[4];
            ldmatrix_x4_trans(smem_addr(kt_v + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kPQ + np * 16 + (lane >> 4) * 8), vb);
            mma_bf16_16816(o[2 * np], pa[kk], vb[0], vb[1]);
            mma_bf16_16816(o[2 * np + 1], pa[kk], vb[2], vb[3]);
          }
        }
      }
      // merge the NS key-range partials of every m-tile
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      float* po = sPO + warp * 16 * 32;
      if (t == 0) {
        sMl[(warp * 16 + g) * 2] = m0;
        sMl[(warp * 16 + g) * 2 + 1] = l0;
        sMl[(warp * 16 + g + 8) * 2] = m1;
        sMl[(warp * 16 + g + 8) * 2 + 1] = l1;
      }
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        *reinterpret_cast<float2*>(po + g * 32 + n * 8 + 2 * t) = make_float2(o[n][0], o[n][1]);
        *reinterpret_cast<float2*>(po + (g + 8) * 32 + n * 8 + 2 * t) = make_float2(o[n][2], o[n][3]);
      }
      __syncthreads();
      for (int idx = tid; idx < M * 8; idx += kThreads) {
        const int row = idx >> 3, c4 = (idx & 7) * 4;
        const int mtr = row >> 4, r16 = row & 15;
        float Mx = -INFINITY;
#pragma unroll
        for (int s2 = 0; s2 < NS; ++s2) Mx = fmaxf(Mx, sMl[((s2 * MT + mtr) * 16 + r16) * 2]);
        float Lsum = 0.0f, acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int s2 = 0; s2 < NS; ++s2) {
          const int w2 = s2 * MT + mtr;
          const float mw = sMl[(w2 * 16 + r16) * 2];
          const float a = mw == -INFINITY ? 0.0f : exp2f((mw - Mx) * sl2);
          Lsum = fmaf(a, sMl[(w2 * 16 + r16) * 2 + 1], Lsum);
          const float4 ov = *reinterpret_cast<const float4*>(sPO + (w2 * 16 + r16) * 32 + c4);
          acc[0] = fmaf(a, ov.x, acc[0]);
          acc[1] = fmaf(a, ov.y, acc[1]);
          acc[2] = fmaf(a, ov.z, acc[2]);
          acc[3] = fmaf(a, ov.w, acc[3]);
        }
        const float inv = 1.0f / Lsum;
        uint2 pk;
        pk.x = float2_to_bf16x2(acc[0] * inv, acc[1] * inv);
        pk.y = float2_to_bf16x2(acc[2] * inv, acc[3] * inv);
        *reinterpret_cast<uint2*>(sSlab0 + row * 32 + c4) = pk;
      }
      __syncthreads();
      mark(layer);   // 3: attention computed
      broadcast_slab<M>(sSlab0, sA0, n0, warp, lane);
      cluster_sync();
      mark(layer);   // 4: attention tile exchanged
    }

    // ======================= out_proj + residual + LayerNorm1 (transformer.py:638-641) =======================
    gemm_slab32<MT, false>(sA0, W.wo, n0, W.bo, sRes, sY, nullptr, warp, lane);
    __syncthreads();
    mark(layer);   // 5: out_proj GEMM
    cluster_layernorm<MT>(sY, sStat, W.g1 + n0, W.be1 + n0, p.eps, rank, warp, lane, [&](int row, int c, float v) {
      sRes[row * 32 + c] = v;
      sSlab0[row * 32 + c] = __float2bfloat16_rn(v + sPos[row * 32 + c]);  // cross-attention query = x + pos (:644)
    });
    __syncthreads();
    mark(layer);   // 6: LayerNorm1 (stats exchange + normalise)
    broadcast_slab<M>(sSlab0, sA1, n0, warp, lane);
    cluster_sync();
    mark(layer);   // 7: x1 + pos exchanged

    // ======================= MSDeformAttn of head `rank` (transformer.py:268-285) =======================
    // offsets | logits of this head: 24 + 12 = 36 projection rows -> 5 n8-tiles x 2 m-groups = 10 warp tasks
    {
      constexpr int NM = MT / 2;
      for (int task = warp; task < 10; task += kWarps) {
        const int j = task % 5, mg = task / 5;
        const int n = j * 8 + g;  // projection row of this head owned by the lane
        const bool ok = n < 3 * kLP;
        const int src = n < 2 * kLP ? static_cast<int>(rank) * 2 * kLP + n : kCluster * 2 * kLP + static_cast<int>(rank) * kLP + (n - 2 * kLP);
        uint4 b[8];
        load_b<8>(b, W.woff + static_cast<int64_t>(ok ? src : 0) * kC + t * 8, ok);
        float acc[NM][4];
#pragma unroll
        for (int i = 0; i < NM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
        mma_tiles<8, NM>(acc, sA1 + (mg * NM * 16) * kPA, kPA, b, lane);
        const int c0 = j * 8 + 2 * t;  // output column pair; bias of columns c0, c0 + 1
        auto bias_of = [&](int c) {
          if (c >= 3 * kLP) return 0.0f;
          const int s2 = c < 2 * kLP ? static_cast<int>(rank) * 2 * kLP + c : kCluster * 2 * kLP + static_cast<int>(rank) * kLP + (c - 2 * kLP);
          return __ldg(W.boff + s2);
        };
        const float b0 = bias_of(c0), b1 = bias_of(c0 + 1);
#pragma unroll
        for (int i = 0; i < NM; ++i) {
          const int r0 = (mg * NM + i) * 16 + g, r1 = r0 + 8;
          *reinterpret_cast<float2*>(sOL + r0 * kPO + c0) = make_float2(acc[i][0] + b0, acc[i][1] + b1);
          *reinterpret_cast<float2*>(sOL + r1 * kPO + c0) = make_float2(acc[i][2] + b0, acc[i][3] + b1);
        }
      }
      __syncthreads();
      mark(layer);   // 8: offsets | logits projection
      // gather: every warp takes M/8 rows, two at a time (lanes 0..15 / 16..31 own the 12 sampling points of the
      // two rows; then 8 sub-groups of 4 lanes fetch 6 corner rows each per item: 12 128-bit loads in flight)
      const __nv_bfloat16* vbase = p.values + static_cast<int64_t>(tile.seq) * p.v_batch_stride + layer * kC + n0;
      int* st_off = sStage + warp * (2 * kNU * 2);
      float* st_w = reinterpret_cast<float*>(st_off + 2 * kNU);
      for (int pr = 0; pr < MT; ++pr) {
        const int half = lane >> 4, pl = lane & 15;
        const int row = warp + kWarps * (2 * pr + half);
        const bool okp = pl < kLP;
        const float* ol = sOL + row * kPO;
        const float lg = okp ? ol[2 * kLP + pl] : -INFINITY;
        const float2 off = okp ? *reinterpret_cast<const float2*>(ol + 2 * pl) : make_float2(0.0f, 0.0f);
        float mx = lg;
#pragma unroll
        for (int o2 = 8; o2 > 0; o2 >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
        if (p.softmax_mode == MOYOLO_SOFTMAX_PLUS1) mx = 0.0f;
        float aw = okp ? expf(lg - mx) : 0.0f;
        float sum = aw;
#pragma unroll
        for (int o2 = 8; o2 > 0; o2 >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o2);
        if (p.softmax_mode == MOYOLO_SOFTMAX_PLUS1) sum += 1.0f;
        aw *= 1.0f / sum;
        const float4 rf = *reinterpret_cast<const float4*>(sRef + row * 4);
        const float lx = rf.x + off.x / static_cast<float>(kNP) * rf.z * 0.5f;  // transformer.py:281-282
        const float ly = rf.y + off.y / static_cast<float>(kNP) * rf.w * 0.5f;
        if (okp) {
          const int level = pl / kNP;
          const Corners c = make_corners(lx, ly, p.lv.h[level], p.lv.w[level], p.lv.start[level], aw);
          *reinterpret_cast<int4*>(st_off + half * kNU + pl * 4) =
              make_int4(c.pos[0] < 0 ? -1 : c.pos[0] * ps, c.pos[1] < 0 ? -1 : c.pos[1] * ps,
                        c.pos[2] < 0 ? -1 : c.pos[2] * ps, c.pos[3] < 0 ? -1 : c.pos[3] * ps);
          *reinterpret_cast<float4*>(st_w + half * kNU + pl * 4) = make_float4(c.w[0], c.w[1], c.w[2], c.w[3]);
        }
        __syncwarp();
        const int sg = lane >> 2, sub = lane & 3;
        uint4 v[2][6];
        float wv[2][6];
#pragma unroll
        for (int it = 0; it < 2; ++it)
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            const int eo = st_off[it * kNU + sg + 8 * u];
            wv[it][u] = st_w[it * kNU + sg + 8 * u];
            v[it][u] = ldg128_if(vbase + eo + sub * 8, eo >= 0);
          }
        float acc[2][8];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[it][k] = 0.0f;
#pragma unroll
          for (int u = 0; u < 6; ++u) fma_bf16x8(acc[it], v[it][u], wv[it][u]);
        }
        // recursive halving across the 8 sub-groups (msda.cu phase 3): 8 -> 4 -> 2 -> 1 channels per lane
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          int ch = 0, n = 8;
#pragma unroll
          for (int o2 = 16; o2 >= 4; o2 >>= 1) {
            n >>= 1;
            const bool up = (lane & o2) != 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (i < n) {
                const float send = up ? acc[it][i] : acc[it][i + n];
                const float recv = __shfl_xor_sync(0xffffffffu, send, o2);
                acc[it][i] = (up ? acc[it][i + n] : acc[it][i]) + recv;
              }
            }
            ch += up ? n : 0;
          }
          const float hi = __shfl_xor_sync(0xffffffffu, acc[it][0], 4);  // odd-channel partner
          const int orow = warp + kWarps * (2 * pr + it);
          if ((lane & 4) == 0)
            *reinterpret_cast<uint32_t*>(sSlab0 + orow * 32 + sub * 8 + ch) = float2_to_bf16x2(acc[it][0], hi);
        }
        __syncwarp();
      }
      __syncthreads();
      mark(layer);   // 9: gather
      broadcast_slab<M>(sSlab0, sA0, n0, warp, lane);
      cluster_sync();
      mark(layer);   // 10: gathered tile exchanged
    }

    // ======================= output_proj + residual + LayerNorm2 (transformer.py:286, 646-647) =======================
    gemm_slab32<MT, false>(sA0, W.wout, n0, W.bout, sRes, sY, nullptr, warp, lane);
    __syncthreads();
    cluster_layernorm<MT>(sY, sStat, W.g2 + n0, W.be2 + n0, p.eps, rank, warp, lane, [&](int row, int c, float v) {
      sRes[row * 32 + c] = v;
      sSlab0[row * 32 + c] = __float2bfloat16_rn(v);
    });
    __syncthreads();
    broadcast_slab<M>(sSlab0, sA1, n0, warp, lane);
    cluster_sync();
    mark(layer);   // 11: output_proj + LayerNorm2 + exchange

    // ======================= FFN (transformer.py:576-580) =======================
    // linear1: this CTA's 128 hidden columns (16 n8-tiles, two per warp), ReLU, kept in shared memory
    {
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int nt = warp * 2 + jj;
        const int hrow = static_cast<int>(rank) * kFs + nt * 8;
        uint4 b[8];
        load_b<8>(b, W.w1 + static_cast<int64_t>(hrow + g) * kC + t * 8, true);
        float acc[MT][4];
#pragma unroll
        for (int i = 0; i < MT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
        mma_tiles<8, MT>(acc, sA1, kPA, b, lane);
        const float b0 = __ldg(W.b1 + hrow + 2 * t), b1 = __ldg(W.b1 + hrow + 2 * t + 1);
        const int col = nt * 8 + 2 * t;
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          *reinterpret_cast<uint32_t*>(sH + (i * 16 + g) * kPH + col) =
              float2_to_bf16x2(fmaxf(acc[i][0] + b0, 0.0f), fmaxf(acc[i][1] + b1, 0.0f));
          *reinterpret_cast<uint32_t*>(sH + (i * 16 + g + 8) * kPH + col) =
              float2_to_bf16x2(fmaxf(acc[i][2] + b0, 0.0f), fmaxf(acc[i][3] + b1, 0.0f));
        }
      }
      __syncthreads();
      mark(layer);   // 12: FFN linear1
      // linear2, split along K: partial[M, 256] = h[:, own 128] . W2[:, own 128]^T; warp w computes output columns
      // [32 w, 32 w + 32) = the slab of cluster rank w and sends them there (reduce-scatter through DSMEM)
      const uint32_t part_peer = map_peer(smem_addr(sPart), static_cast<uint32_t>(warp));
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int ncol = warp * 32 + jj * 8;
        uint4 b[4];
        load_b<4>(b, W.w2 + static_cast<int64_t>(ncol + g) * (kFs * kCluster) + static_cast<int>(rank) * kFs + t * 8, true);
        float acc[MT][4];
#pragma unroll
        for (int i = 0; i < MT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
        mma_tiles<4, MT>(acc, sH, kPH, b, lane);
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          const int r0 = i * 16 + g, r1 = r0 + 8;
          const int c = jj * 8 + 2 * t;
          st_peer_v2f(part_peer + static_cast<uint32_t>(((rank * M + r0) * 32 + c) * 4), acc[i][0], acc[i][1]);
          st_peer_v2f(part_peer + static_cast<uint32_t>(((rank * M + r1) * 32 + c) * 4), acc[i][2], acc[i][3]);
        }
      }
      cluster_sync();
      mark(layer);   // 13: FFN linear2 (split-K) + reduce-scatter
      for (int i = tid; i < M * 32; i += kThreads) {
        const int row = i >> 5, c = i & 31;
        float s2 = __ldg(W.b2 + n0 + c) + sRes[row * 32 + c];
#pragma unroll
        for (int src = 0; src < kCluster; ++src) s2 += sPart[(src * M + row) * 32 + c];  // fixed order: deterministic
        sY[row * kPY + c] = s2;
      }
      __syncthreads();
      cluster_layernorm<MT>(sY, sStat, W.g3 + n0, W.be3 + n0, p.eps, rank, warp, lane, [&](int row, int c, float v) {
        sRes[row * 32 + c] = v;
        sSlab0[row * 32 + c] = __float2bfloat16_rn(v);
        sSlab1[row * 32 + c] = __float2bfloat16_rn(v + sPos[row * 32 + c]);
      });
      __syncthreads();
      broadcast_slab<M>(sSlab0, sA0, n0, warp, lane);
      if (!last) broadcast_slab<M>(sSlab1, sA1, n0, warp, lane);
      cluster_sync();
      mark(layer);   // 14: LayerNorm3 + exchange
    }

    if (last) {
      // output embedding (fp32 + optional bf16) and the class-score head on the bf16-rounded row
      for (int row = warp; row < tile.n; row += kWarps) {
        const int64_t gr = static_cast<int64_t>(tile.row0 + row) * kC + n0 + lane;
        p.x_out[gr] = sRes[row * 32 + lane];
        if (p.x_lp_out != nullptr) p.x_lp_out[gr] = sSlab0[row * 32 + lane];
      }
      if (p.nc > 0) {
        for (int row = static_cast<int>(rank) + kCluster * warp; row < tile.n; row += kCluster * kWarps) {
          const uint4 xv = *reinterpret_cast<const uint4*>(sA0 + row * kPA + lane * 8);
          const float2 x01 = bf16x2_to_float2(xv.x), x23 = bf16x2_to_float2(xv.y), x45 = bf16x2_to_float2(xv.z),
                       x67 = bf16x2_to_float2(xv.w);
          float best = -INFINITY;
          int best_c = 0;
          for (int c = 0; c < p.nc; ++c) {
            const float4 wa = __ldg(reinterpret_cast<const float4*>(p.score_w + c * kC + lane * 8));
            const float4 wb = __ldg(reinterpret_cast<const float4*>(p.score_w + c * kC + lane * 8 + 4));
            float d = x01.x * wa.x;
            d = fmaf(x01.y, wa.y, d);
            d = fmaf(x23.x, wa.z, d);
            d = fmaf(x23.y, wa.w, d);
            d = fmaf(x45.x, wb.x, d);
            d = fmaf(x45.y, wb.y, d);
            d = fmaf(x67.x, wb.z, d);
            d = fmaf(x67.y, wb.w, d);
            d = warp_sum(d) + __ldg(p.score_b + c);
            if (lane == 0 && p.logits != nullptr) p.logits[static_cast<int64_t>(tile.row0 + row) * p.nc + c] = d;
            if (d > best) { best = d; best_c = c; }  // first maximum wins, as torch.max
          }
          if (lane == 0) {
            if (p.scores != nullptr) p.scores[tile.row0 + row] = sigmoidf_(best);
            if (p.labels != nullptr) p.labels[tile.row0 + row] = best_c;
          }
        }
      }
    } else {
      qkv_proj(p.L[layer + 1], layer + 1);
      grid_arrive(p.grid_bar, rank);  // (its cluster barrier also frees sA1 for the box head below)
    }

    mark(layer);   // 15: next in-projection + grid arrive (or outputs + scores)
    // ======================= box head + refinement (transformer.py:709) =======================
    gemm_slab32<MT, true>(sA0, W.wb1, n0, W.bb1, nullptr, nullptr, sSlab0, warp, lane);
    __syncthreads();
    broadcast_slab<M>(sSlab0, sA1, n0, warp, lane);
    cluster_sync();
    gemm_slab32<MT, true>(sA1, W.wb2, n0, W.bb2, nullptr, nullptr, sSlab0, warp, lane);
    __syncthreads();
    broadcast_slab<M>(sSlab0, sA0, n0, warp, lane);
    cluster_sync();
    {
      float w3[4][8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(W.wb3 + j * kC + lane * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(W.wb3 + j * kC + lane * 8 + 4));
        w3[j][0] = a.x; w3[j][1] = a.y; w3[j][2] = a.z; w3[j][3] = a.w;
        w3[j][4] = b.x; w3[j][5] = b.y; w3[j][6] = b.z; w3[j][7] = b.w;
      }
      const float bias = lane < 4 ? __ldg(W.bb3 + lane) : 0.0f;
      float* rout = p.refer_out[layer];
      for (int row = warp; row < M; row += kWarps) {
        const uint4 hv = *reinterpret_cast<const uint4*>(sA0 + row * kPA + lane * 8);
        const float2 h01 = bf16x2_to_float2(hv.x), h23 = bf16x2_to_float2(hv.y), h45 = bf16x2_to_float2(hv.z),
                     h67 = bf16x2_to_float2(hv.w);
        float d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = h01.x * w3[j][0];
          a = fmaf(h01.y, w3[j][1], a);
          a = fmaf(h23.x, w3[j][2], a);
          a = fmaf(h23.y, w3[j][3], a);
          a = fmaf(h45.x, w3[j][4], a);
          a = fmaf(h45.y, w3[j][5], a);
          a = fmaf(h67.x, w3[j][6], a);
          a = fmaf(h67.y, w3[j][7], a);
          d[j] = warp_sum(a);
        }
        if (lane < 4) {
          const float tt = (lane == 0 ? d[0] : (lane == 1 ? d[1] : (lane == 2 ? d[2] : d[3]))) + bias +
                           inverse_sigmoidf_(sRef[row * 4 + lane]);
          const float nb = sigmoidf_(tt);
          sRef[row * 4 + lane] = nb;
          if (rank == 0 && rout != nullptr && row < tile.n) rout[static_cast<int64_t>(tile.row0 + row) * 4 + lane] = nb;
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace dc
}  // namespace moyolo

using namespace moyolo;

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
namespace {

template <int MT>
int dc_configure(int* max_clusters, size_t* smem_out) {
  static DeviceOnce once;
  static int max_cl[64];
  const int dev = DeviceOnce::current();
  const size_t smem = dc::Smem<MT>::total();
  if (!once.done(dev)) {
    cudaError_t e = cudaFuncSetAttribute(dc::decoder_cluster_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return fail(MOYOLO_ERR_CUDA, "cudaFuncSetAttribute(decoder_cluster smem=%zu): %s", smem, cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(dc::kCluster * 18);
    cfg.blockDim = dim3(dc::kThreads);
    cfg.dynamicSmemBytes = smem;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, dc::decoder_cluster_kernel<MT>, &cfg);
    if (e != cudaSuccess) return fail(MOYOLO_ERR_CUDA, "cudaOccupancyMaxActiveClusters(decoder_cluster): %s", cudaGetErrorString(e));
    max_cl[dev] = n;
    once.set(dev);
  }
  *max_clusters = max_cl[dev];
  *smem_out = smem;
  return MOYOLO_OK;
}

}  // namespace

extern "C" int moyolo_decoder_cluster_limits(int rows_per_tile, int* max_clusters, int* kv_cap) {
  MOYOLO_REQUIRE(rows_per_tile == 32 || rows_per_tile == 64, MOYOLO_ERR_BAD_ARG, "decoder_cluster: rows_per_tile must be 32 or 64");
  size_t smem = 0;
  int mc = 0;
  const int rc = rows_per_tile == 32 ? dc_configure<2>(&mc, &smem) : dc_configure<4>(&mc, &smem);
  if (rc != MOYOLO_OK) return rc;
  if (max_clusters) *max_clusters = mc;
  if (kv_cap) *kv_cap = rows_per_tile == 32 ? dc::Smem<2>::kv_cap() : dc::Smem<4>::kv_cap();
  return MOYOLO_OK;
}

extern "C" int moyolo_decoder_cluster_forward(const moyolo_decoder_cluster_t* a, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(a != nullptr, MOYOLO_ERR_BAD_ARG, "decoder_cluster: null descriptor");
  MOYOLO_REQUIRE(a->n_layers >= 1 && a->n_layers <= dc::kMaxLayers, MOYOLO_ERR_BAD_ARG, "decoder_cluster: n_layers must be in [1, %d]", dc::kMaxLayers);
  MOYOLO_REQUIRE(a->d_model == dc::kC && a->n_heads == 8 && a->d_ffn == dc::kFs * dc::kCluster && a->n_levels == dc::kNL && a->n_points == dc::kNP,
                 MOYOLO_ERR_UNSUPPORTED, "decoder_cluster: built for d_model 256, 8 heads, d_ffn 1024, 3 levels x 4 points");
  MOYOLO_REQUIRE(a->x_in && a->pos && a->refer0 && a->x_out && a->kv && a->values && a->row_offsets && a->grid_barrier,
                 MOYOLO_ERR_BAD_ARG, "decoder_cluster: null pointer");
  MOYOLO_REQUIRE(a->nc >= 0 && a->nc <= dc::kMaxScoreNc, MOYOLO_ERR_UNSUPPORTED, "decoder_cluster: nc must be <= %d", dc::kMaxScoreNc);
  MOYOLO_REQUIRE(a->n_seq >= 1 && a->rows_pad >= 1, MOYOLO_ERR_BAD_SHAPE, "decoder_cluster: bad n_seq / rows_pad");
  MOYOLO_REQUIRE(a->rows_per_tile == 32 || a->rows_per_tile == 64, MOYOLO_ERR_BAD_ARG, "decoder_cluster: rows_per_tile must be 32 or 64");
  dc::Params p = {};
  for (int l = 0; l < a->n_layers; ++l) {
    const moyolo_decoder_layer_weights_t& s = a->layers[l];
    dc::LayerW& d = p.L[l];
    MOYOLO_REQUIRE(s.wqkv && s.wo && s.woff && s.wout && s.w1 && s.w2 && s.wb1 && s.wb2 && s.wb3, MOYOLO_ERR_BAD_ARG,
                   "decoder_cluster: layer %d has a null weight", l);
    d.wqkv = static_cast<const __nv_bfloat16*>(s.wqkv); d.wo = static_cast<const __nv_bfloat16*>(s.wo);
    d.woff = static_cast<const __nv_bfloat16*>(s.woff); d.wout = static_cast<const __nv_bfloat16*>(s.wout);
    d.w1 = static_cast<const __nv_bfloat16*>(s.w1); d.w2 = static_cast<const __nv_bfloat16*>(s.w2);
    d.wb1 = static_cast<const __nv_bfloat16*>(s.wb1); d.wb2 = static_cast<const __nv_bfloat16*>(s.wb2);
    d.bqkv = s.bqkv; d.bo = s.bo; d.boff = s.boff; d.bout = s.bout; d.b1 = s.b1; d.b2 = s.b2; d.bb1 = s.bb1; d.bb2 = s.bb2;
    d.wb3 = s.wb3; d.bb3 = s.bb3;
    d.g1 = s.ln1_w; d.be1 = s.ln1_b; d.g2 = s.ln2_w; d.be2 = s.ln2_b; d.g3 = s.ln3_w; d.be3 = s.ln3_b;
    p.refer_out[l] = a->refer_out[l];
  }
  p.n_layers = a->n_layers;
  p.x_in = a->x_in; p.pos = a->pos; p.refer0 = a->refer0; p.x_out = a->x_out;
  p.x_lp_out = static_cast<__nv_bfloat16*>(a->x_lp_out);
  p.kv = static_cast<__nv_bfloat16*>(a->kv);
  p.values = static_cast<const __nv_bfloat16*>(a->values);
  p.v_batch_stride = a->value_batch_stride; p.v_pos_stride = a->value_pos_stride;
  int64_t len_v = 0;
  for (int l = 0; l < a->n_levels; ++l) len_v += static_cast<int64_t>(a->value_shapes[2 * l]) * a->value_shapes[2 * l + 1];
  int rc = make_levels(a->value_shapes, a->n_levels, len_v, &p.lv);
  if (rc != MOYOLO_OK) return rc;
  p.softmax_mode = a->softmax_mode;
  p.ro = a->row_offsets; p.n_seq = a->n_seq; p.rows_pad = static_cast<int>(a->rows_pad);
  p.grid_bar = static_cast<unsigned*>(a->grid_barrier);
  p.status = a->status;
  p.score_w = a->score_w; p.score_b = a->score_b; p.nc = a->score_w != nullptr ? a->nc : 0;
  p.logits = a->logits; p.scores = a->scores; p.labels = a->labels;
  p.eps = a->eps;
  p.profile = static_cast<long long*>(a->profile);
  size_t smem = 0;
  int max_clusters = 0;
  const bool small = a->rows_per_tile == 32;
  rc = small ? dc_configure<2>(&max_clusters, &smem) : dc_configure<4>(&max_clusters, &smem);
  if (rc != MOYOLO_OK) return rc;
  p.kv_cap = small ? dc::Smem<2>::kv_cap() : dc::Smem<4>::kv_cap();
  // tiles <= rows_pad / M + n_seq (every sequence may end with a partial tile)
  const int64_t tiles_bound = (a->rows_pad + a->rows_per_tile - 1) / a->rows_per_tile + (a->n_seq - 1);
  MOYOLO_REQUIRE(tiles_bound <= max_clusters, MOYOLO_ERR_UNSUPPORTED,
                 "decoder_cluster: %lld row tiles need more than the %d co-resident clusters of this device",
                 (long long)tiles_bound, max_clusters);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->reset_barrier) {
    cudaError_t e = cudaMemsetAsync(a->grid_barrier, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return fail(MOYOLO_ERR_CUDA, "decoder_cluster: memset of the grid barrier: %s", cudaGetErrorString(e));
  }
  const dim3 grid(static_cast<unsigned>(tiles_bound) * dc::kCluster);
  if (small)
    launch_k(dc::decoder_cluster_kernel<2>, grid, dim3(dc::kThreads), smem, st, p);
  else
    launch_k(dc::decoder_cluster_kernel<4>, grid, dim3(dc::kThreads), smem, st, p);
  return check_launch("decoder_cluster_kernel");
}
