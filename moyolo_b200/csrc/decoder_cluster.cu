// Row-tile-persistent decoder: all layers of the MOTR / deformable decoder in ONE launch.
//
// Replaces, for one frame: MOTRTransformerDecoder.forward (ultralytics/nn/modules/transformer.py:676-728) = 6 x
// MOTRDecoderLayer.forward (:627-652: nn.MultiheadAttention self-attention + LayerNorm, MSDeformAttn :246-287 +
// LayerNorm, FFN :576-580 + LayerNorm) + the per-layer box refinement sigmoid(bbox_head(x) + inverse_sigmoid(ref))
// (:709) + the class-score head of the last layer (:717-721). The 7 + 3 launches per layer of the launch-chained
// schedule (executor.run_layer_ws) become one kernel for the whole decoder; the value projection stays a separate
// tcgen05 GEMM (it depends on the frame's feature maps only and runs ahead of the frame).
//
// Decomposition. A thread-block CLUSTER of 8 CTAs owns a tile of M = 32 query rows through every row-local
// operation of every layer; CTA rank r of the cluster owns
//   * attention head r (self-attention and deformable gather are per head), and
//   * the 32-column slab r of every 256-wide GEMM output (128-column slab of the FFN hidden layer).
// GEMM operands (activation tiles [32, 256] bf16) are REPLICATED in the shared memory of the 8 CTAs. A stage
// computes its slab, stages it in its own shared memory and sends it to all 8 replicas with bulk asynchronous
// distributed-shared-memory copies (cp.async.bulk.shared::cluster.shared::cta, one row per thread) that complete a
// transaction barrier in the RECEIVING CTA (mbarrier::complete_tx): a receiver waits on its own mbarrier for the
// expected byte count, there is no cluster-wide barrier in the layer loop. Buffers are
// re-used only along the dependency chain of the exchanges themselves (a peer can send exchange k + 1 only after
// it received this CTA's part of exchange k), which is what makes the re-use race-free without extra handshakes.
//   * LayerNorm: the pre-norm fp32 slab is sent to all peers, every CTA then normalises the full rows locally
//     (two-pass mean / variance in registers) and writes the bf16 operand tile(s) of the next GEMM itself.
//   * FFN: linear1 produces this CTA's 128 hidden columns, linear2 is split along K (no exchange of the hidden
//     activations) and the partial sums are reduce-scattered to the slab owners.
//   * Weights are never staged in shared memory: each warp loads its mma.sync B fragments straight from global
//     memory / L2 with 128-bit read-only loads (8 consecutive k per lane, a k-permutation mirrored by the A
//     fragments) and ISSUES them one stage ahead, so their latency hides behind the preceding exchange.
//   * Biases and LayerNorm parameters of a layer are staged in shared memory one layer ahead (cp.async).
// The only data another cluster needs are the keys / values of the self-attention: they go through global memory
// (double-buffered by layer parity) and a grid-wide barrier per layer (release/acquire on a global counter; all
// clusters are co-resident, the host checks the occupancy). Tensor work is mma.sync m16n8k16 (bf16 in, fp32
// accumulate): the tiles are 32 rows x 8..32 columns per warp task, far below a 128-row tcgen05 tile, and the
// kernel is latency-bound (benchmarks/dc_stages.py prints the stage timeline).
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
constexpr int M = 32;     // query rows per cluster tile
constexpr int kC = 256, kDh = 32;
constexpr int kFs = 128;  // FFN hidden columns per CTA (d_ffn 1024 / 8)
constexpr int kPA = 288;  // bf16 pitch of the activation tiles: 576 B = 64 (mod 128) -> conflict-free 128-bit A loads
constexpr int kPH = 160;  // bf16 pitch of the FFN hidden slab [M][128]: 320 B = 64 (mod 128)
constexpr int kPQ = 40;   // bf16 pitch of Q / K / V head rows (ldmatrix layout of attention.cu)
constexpr int kPY = 260;  // fp32 pitch of the pre-LayerNorm tile [M][256]
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
  long long* profile;     // optional [n_layers][16] globaltimer stamps of cluster 0 / rank 0 (benchmarks/dc_stages.py)
};

// ---------------------------------------------------------------------------------------------------------
// Shared-memory plan of one CTA (bytes). The tail [TB, TC, TD, H, PART] doubles as the key / value staging of the
// self-attention (none of those buffers is live between the grid barrier and the end of the attention).
// ---------------------------------------------------------------------------------------------------------
namespace sm {
constexpr size_t kTile = size_t(M) * kPA * 2;             // one bf16 activation tile
constexpr size_t oTA = 0;                                 // attention output / gathered tile (received)
constexpr size_t oYF = oTA + kTile;                       // pre-LayerNorm fp32 tile (received); attention partial O
constexpr size_t oRes = oYF + size_t(M) * kPY * 4;        // residual stream, own 32 columns, fp32
constexpr size_t oPos = oRes + size_t(M) * 32 * 4;        // query_pos, full rows, fp32
constexpr size_t oOL = oPos + size_t(M) * kC * 4;         // offsets | logits of this head
constexpr size_t oQ = oOL + size_t(M) * kPO * 4;          // q of this head
constexpr size_t oRef = oQ + size_t(M) * kPQ * 2;         // reference boxes of the tile
constexpr size_t oStage = oRef + size_t(M) * 16;          // gather staging: per warp 4 items x 48 corners x (off, w)
constexpr size_t oMl = oStage + size_t(kWarps) * 4 * kNU * 8;   // attention partial (max, sum)
constexpr size_t oSlab = oMl + size_t(kWarps) * 16 * 2 * 4;     // bf16 slab staging [M][32]
constexpr int kParamFloats = 1968;                        // see PF:: below
constexpr size_t oParam = oSlab + size_t(M) * 32 * 2;     // 2 x per-layer parameter block (layer parity)
constexpr size_t oBq0 = oParam + 2 * size_t(kParamFloats) * 4;  // in-projection bias of layer 0 (this head)
constexpr size_t oYs = oBq0 + 96 * 4;                     // fp32 slab staging [M][32] (source of the pre-LN bulk copies)
constexpr size_t oBars = oYs + size_t(M) * 32 * 4;        // 8 mbarriers
constexpr size_t oTB = (oBars + 8 * 8 + 127) / 128 * 128; // x + pos operand tiles (written locally by the LayerNorms)
constexpr size_t oTC = oTB + kTile;                       // x3 tile (local) / box-head hidden 2 (received)
constexpr size_t oTD = oTC + kTile;                       // box-head hidden 1 (received)
constexpr size_t oH = oTD + kTile;                        // FFN hidden slab
constexpr size_t oPart = oH + size_t(M) * kPH * 2;        // split-K partial sums [8 sources][M][32] fp32 (received)
constexpr size_t kTotal = 232448 - 512;                   // everything a CTA can opt in to, minus alignment slack
constexpr size_t kPartBytes = size_t(kCluster) * M * 32 * 4;
static_assert(oPart + kPartBytes <= kTotal, "shared-memory plan exceeds 227 KiB");
constexpr int kKvCap = static_cast<int>((kTotal - oTB) / (2 * kPQ * 2)) / 32 * 32;   // keys the tail can stage
constexpr size_t kPoBytes = size_t(kWarps) * 16 * 32 * 4;
static_assert(kPoBytes <= size_t(M) * kPY * 4, "attention partials must fit the pre-LayerNorm tile they alias");
}  // namespace sm

// per-layer parameter block (float offsets)
namespace PF {
constexpr int bo = 0, bout = 32, b2 = 64, bb1 = 96, bb2 = 128, boff = 160 /* 24 offsets + 12 logits, pad to 40 */,
              b1 = 200, bqkv = 328 /* q | k | v of the NEXT layer's in-projection */, bb3 = 424,
              g1 = 432, be1 = 688, g2 = 944, be2 = 1200, g3 = 1456, be3 = 1712;
static_assert(be3 + 256 == sm::kParamFloats, "parameter block layout");
}  // namespace PF

enum { E_ATT = 0, E_LN1, E_G, E_LN2, E_RS, E_LN3, E_H1, E_H2, kNumBars };
constexpr uint32_t kBf16TileBytes = kCluster * M * 32 * 2;   // 8 senders x [32][32] bf16
constexpr uint32_t kF32TileBytes = kCluster * M * 32 * 4;    // 8 senders x [32][32] fp32

// ---------------------------------------------------------------------------------------------------------
// cluster / DSMEM / mbarrier / grid-barrier primitives
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
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arm(uint32_t bar, uint32_t bytes) {  // one arrival + the bytes this phase expects
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// Bulk asynchronous copy (async proxy / TMA engine) of `bytes` (multiple of 16) from this CTA's shared memory into a
// CTA of the cluster; completes the bytes on that CTA's mbarrier. Generic-proxy writes of the source must be made
// visible to the async proxy first (fence_proxy_async by the writers, then a block barrier).
__device__ __forceinline__ void bulk_to_peer(uint32_t remote_addr, uint32_t local_src, uint32_t bytes, uint32_t remote_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(remote_addr),
               "r"(local_src), "r"(bytes), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Grid barrier: every CTA counts itself in once all its threads' global writes of the phase are done.
__device__ __forceinline__ void grid_arrive(unsigned* bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    red_release_gpu_add(bar, 1u);
  }
}
__device__ __forceinline__ void grid_wait(const unsigned* bar, unsigned target) {
  if (threadIdx.x == 0) {
    while (ld_acquire_gpu(bar) < target) __nanosleep(20);
    __threadfence();
  }
  __syncthreads();
}
__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
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

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float inverse_sigmoidf_(float x) {  // utils.py:34-38
  x = fminf(fmaxf(x, 0.0f), 1.0f);
  return logf(fmaxf(x, 1e-5f) / fmaxf(1.0f - x, 1e-5f));
}

struct Tile {
  int seq, row0, n, seq_start, seq_len;
};

// tile `idx` of the frame: tiles are enumerated per sequence (a tile never spans two sequences).
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

// ---------------------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
    decoder_cluster_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(128) uint8_t dc_smem[];
  __nv_bfloat16* sTA = reinterpret_cast<__nv_bfloat16*>(dc_smem + sm::oTA);
  float* sYF = reinterpret_cast<float*>(dc_smem + sm::oYF);
  float* sPO = sYF;  // attention partial O aliases the pre-LayerNorm tile
  float* sRes = reinterpret_cast<float*>(dc_smem + sm::oRes);
  float* sPos = reinterpret_cast<float*>(dc_smem + sm::oPos);
  float* sOL = reinterpret_cast<float*>(dc_smem + sm::oOL);
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(dc_smem + sm::oQ);
  float* sRef = reinterpret_cast<float*>(dc_smem + sm::oRef);
  int* sStage = reinterpret_cast<int*>(dc_smem + sm::oStage);
  float* sMl = reinterpret_cast<float*>(dc_smem + sm::oMl);
  __nv_bfloat16* sSlab = reinterpret_cast<__nv_bfloat16*>(dc_smem + sm::oSlab);
  float* sParam = reinterpret_cast<float*>(dc_smem + sm::oParam);
  float* sBq0 = reinterpret_cast<float*>(dc_smem + sm::oBq0);
  __nv_bfloat16* sTB = reinterpret_cast<__nv_bfloat16*>(dc_smem + sm::oTB);
  __nv_bfloat16* sTC = reinterpret_cast<__nv_bfloat16*>(dc_smem + sm::oTC);
  __nv_bfloat16* sTD = reinterpret_cast<__nv_bfloat16*>(dc_smem + sm::oTD);
  __nv_bfloat16* sH = reinterpret_cast<__nv_bfloat16*>(dc_smem + sm::oH);
  float* sPart = reinterpret_cast<float*>(dc_smem + sm::oPart);
  __nv_bfloat16* sK = sTB;  // key / value staging aliases [TB, TC, TD, H, PART]
  __nv_bfloat16* sV = sK + static_cast<size_t>(sm::kKvCap) * kPQ;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t rank = cluster_rank();
  const int cl = static_cast<int>(blockIdx.x) / kCluster;
  const int n0 = static_cast<int>(rank) * 32;  // this CTA's column slab / head
  const uint32_t smem_base = smem_addr(dc_smem);
  pdl_wait();

  Tile tile;
  int total_tiles, max_len;
  const bool have = find_tile(p, cl, &tile, &total_tiles, &max_len);
  const int n_clusters = static_cast<int>(gridDim.x) / kCluster;
  if (total_tiles > n_clusters || max_len > sm::kKvCap) {  // cannot run: the host sized the launch wrongly
    if (p.status != nullptr && blockIdx.x == 0 && tid == 0) *p.status = 1;
    return;
  }
  if (!have) return;  // whole cluster idle (uniform over its 8 CTAs)
  const unsigned n_cta_u = static_cast<unsigned>(total_tiles) * kCluster;

  auto bar_local = [&](int e) { return smem_base + static_cast<uint32_t>(sm::oBars + e * 8); };

  if (tid == 0) {
    for (int e = 0; e < kNumBars; ++e) mbar_init(bar_local(e), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto arm_all = [&]() {  // expected bytes of one layer's exchanges (thread 0)
    mbar_arm(bar_local(E_ATT), kBf16TileBytes);
    mbar_arm(bar_local(E_LN1), kF32TileBytes);
    mbar_arm(bar_local(E_G), kBf16TileBytes);
    mbar_arm(bar_local(E_LN2), kF32TileBytes);
    mbar_arm(bar_local(E_RS), kF32TileBytes);
    mbar_arm(bar_local(E_LN3), kF32TileBytes);
    mbar_arm(bar_local(E_H1), kBf16TileBytes);
    mbar_arm(bar_local(E_H2), kBf16TileBytes);
  };
  if (tid == 0) arm_all();
  cluster_sync();  // every CTA's barriers are initialised and armed before any peer may send

  // ---- per-layer parameter block -> shared memory (cp.async; issued one layer ahead) ----
  auto stage_params = [&](int layer) {
    float* dst = sParam + (layer & 1) * sm::kParamFloats;
    const LayerW& W = p.L[layer];
    auto cp = [&](int off, const float* src, int n) {  // n floats, multiple of 4, 16-byte aligned source
      for (int i = tid; i < n / 4; i += kThreads) cp_async16(smem_addr(dst + off + i * 4), src + i * 4, true);
    };
    cp(PF::bo, W.bo + n0, 32);
    cp(PF::bout, W.bout + n0, 32);
    cp(PF::b2, W.b2 + n0, 32);
    cp(PF::bb1, W.bb1 + n0, 32);
    cp(PF::bb2, W.bb2 + n0, 32);
    cp(PF::boff, W.boff + static_cast<int>(rank) * 2 * kLP, 2 * kLP);
    cp(PF::boff + 2 * kLP, W.boff + kCluster * 2 * kLP + static_cast<int>(rank) * kLP, kLP);
    cp(PF::b1, W.b1 + static_cast<int>(rank) * kFs, kFs);
    if (layer + 1 < p.n_layers) {
      const float* bq = p.L[layer + 1].bqkv;
      cp(PF::bqkv, bq + n0, 32);
      cp(PF::bqkv + 32, bq + kC + n0, 32);
      cp(PF::bqkv + 64, bq + 2 * kC + n0, 32);
    }
    cp(PF::bb3, W.bb3, 4);
    cp(PF::g1, W.g1, kC);
    cp(PF::be1, W.be1, kC);
    cp(PF::g2, W.g2, kC);
    cp(PF::be2, W.be2, kC);
    cp(PF::g3, W.g3, kC);
    cp(PF::be3, W.be3, kC);
    cp_async_commit();
  };
  stage_params(0);

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
    *reinterpret_cast<uint2*>(sTC + row * kPA + c4) = a;   // x       (operand of v)
    *reinterpret_cast<uint2*>(sTB + row * kPA + c4) = b;   // x + pos (operand of q, k)
    *reinterpret_cast<float4*>(sPos + row * kC + c4) = pv;
    if (c4 >= n0 && c4 < n0 + 32) *reinterpret_cast<float4*>(sRes + row * 32 + (c4 - n0)) = xv;
  }
  for (int i = tid; i < M * 4; i += kThreads) {
    const int row = i >> 2;
    sRef[i] = row < tile.n ? p.refer0[static_cast<int64_t>(tile.row0 + row) * 4 + (i & 3)] : 0.5f;
  }
  // (the parameter block holds the NEXT layer's in-projection bias; layer 0's is read directly)
  if (tid < 96) sBq0[tid] = __ldg(p.L[0].bqkv + (tid >> 5) * kC + n0 + (tid & 31));
  __syncthreads();

  // In-projection of the next self-attention for the tile's rows (transformer.py:637-638): q, k from x + pos (TB),
  // v from x (TC), this head only; q stays in shared memory, k / v go to global memory for the other clusters.
  // 12 n8-tiles x 2 m-tiles = 24 warp tasks, three per warp.
  uint4 bq[3][8];
  auto qkv_load = [&](const LayerW& W) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int task = warp + kWarps * i, nt = task % 12;
      const int wrow = (nt >> 2) * kC + n0 + (nt & 3) * 8;
      load_b<8>(bq[i], W.wqkv + static_cast<int64_t>(wrow + g) * kC + t * 8, true);
    }
  };
  auto qkv_compute = [&](const float* bias /* q|k|v of this head, 96 floats */, int for_layer) {
    __nv_bfloat16* kv_w = p.kv + static_cast<size_t>(for_layer & 1) * p.rows_pad * (2 * kC);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int task = warp + kWarps * i, nt = task % 12, mt = task / 12;
      const int sec = nt >> 2, j = nt & 3;  // section 0 = q, 1 = k, 2 = v
      float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
      mma_tiles<8, 1>(acc, (sec == 2 ? sTC : sTB) + (mt * 16) * kPA, kPA, bq[i], lane);
      const int col = j * 8 + 2 * t;
      const float b0 = bias[sec * 32 + col], b1 = bias[sec * 32 + col + 1];
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
  // one 32-column bf16 slab GEMM with ReLU (box-head hidden layers): warp task = (n8-tile w & 3, m-tile w >> 2)
  auto relu_slab = [&](const __nv_bfloat16* sA, const uint4 (&b)[8], const float* bias32) {
    float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
    mma_tiles<8, 1>(acc, sA + ((warp >> 2) * 16) * kPA, kPA, b, lane);
    const int col = (warp & 3) * 8 + 2 * t, r0 = (warp >> 2) * 16 + g;
    const float b0 = bias32[col], b1 = bias32[col + 1];
    *reinterpret_cast<uint32_t*>(sSlab + r0 * 32 + col) = float2_to_bf16x2(fmaxf(acc[0][0] + b0, 0.f), fmaxf(acc[0][1] + b1, 0.f));
    *reinterpret_cast<uint32_t*>(sSlab + (r0 + 8) * 32 + col) = float2_to_bf16x2(fmaxf(acc[0][2] + b0, 0.f), fmaxf(acc[0][3] + b1, 0.f));
  };
  // send this CTA's bf16 slab (sSlab [M][32]) into columns [n0, n0+32) of tile `tile_off` of every CTA: warp w ->
  // peer w, 16-byte chunks; completes bytes on exchange barrier e of the receiver
  auto send_slab = [&](size_t tile_off, int e) {   // (callers: fence_proxy_async() + __syncthreads() after writing sSlab)
    const int row = tid >> 3;
    const uint32_t pb = map_peer(smem_base, static_cast<uint32_t>(tid & 7));
    bulk_to_peer(pb + static_cast<uint32_t>(tile_off + (row * kPA + n0) * 2), smem_addr(sSlab + row * 32), 64u,
                 pb + static_cast<uint32_t>(sm::oBars + e * 8));
  };
  // send this CTA's fp32 pre-LayerNorm slab (sYs [M][32], staged in shared memory) to the pre-LN tile of every CTA
  float* sYs = reinterpret_cast<float*>(dc_smem + sm::oYs);
  auto send_preln = [&](int e) {                  // (callers: fence_proxy_async() + __syncthreads() after writing sYs)
    const int row = tid >> 3;
    const uint32_t pb = map_peer(smem_base, static_cast<uint32_t>(tid & 7));
    bulk_to_peer(pb + static_cast<uint32_t>(sm::oYF + (row * kPY + n0) * 4), smem_addr(sYs + row * 32), 128u,
                 pb + static_cast<uint32_t>(sm::oBars + e * 8));
  };
  // 32-column fp32 slab GEMM over K = 256 whose pre-LayerNorm result (+ bias + residual) goes STRAIGHT from the
  // accumulator registers to the pre-LN tile of all 8 CTAs (one 16-byte st.async per peer and lane)
  auto preln_slab = [&](const __nv_bfloat16* sA, const uint4 (&b)[8], const float* bias32, int e) {
    float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
    mma_tiles<8, 1>(acc, sA + ((warp >> 2) * 16) * kPA, kPA, b, lane);
    const int col = (warp & 3) * 8 + 2 * t, r0 = (warp >> 2) * 16 + g;
    const float b0 = bias32[col], b1 = bias32[col + 1];
    const float c[4] = {acc[0][0] + b0 + sRes[r0 * 32 + col], acc[0][1] + b1 + sRes[r0 * 32 + col + 1],
                        acc[0][2] + b0 + sRes[(r0 + 8) * 32 + col], acc[0][3] + b1 + sRes[(r0 + 8) * 32 + col + 1]};
    *reinterpret_cast<float2*>(sYs + r0 * 32 + col) = make_float2(c[0], c[1]);
    *reinterpret_cast<float2*>(sYs + (r0 + 8) * 32 + col) = make_float2(c[2], c[3]);
    fence_proxy_async();
    __syncthreads();
    send_preln(e);
  };
  // LayerNorm of the full rows of the received pre-LN tile (two-pass, fp32): warp w normalises rows w, w+8, ...;
  // lane l owns columns [8 l, 8 l + 8). emit(row, col0, values[8]).
  auto layernorm_rows = [&](const float* gamma, const float* beta, auto emit) {
#pragma unroll
    for (int i = 0; i < M / kWarps; ++i) {
      const int row = warp + i * kWarps;
      const float4 a = *reinterpret_cast<const float4*>(sYF + row * kPY + lane * 8);
      const float4 b = *reinterpret_cast<const float4*>(sYF + row * kPY + lane * 8 + 4);
      float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[k];
      const float mean = warp_sum(s) * (1.0f / kC);
      float q2 = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] -= mean;
        q2 = fmaf(v[k], v[k], q2);
      }
      const float rstd = rsqrtf(warp_sum(q2) * (1.0f / kC) + p.eps);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = v[k] * rstd * gamma[lane * 8 + k] + beta[lane * 8 + k];
      if ((lane >> 2) == static_cast<int>(rank)) {  // the four lanes that hold this CTA's residual slab
        *reinterpret_cast<float4*>(sRes + row * 32 + (lane & 3) * 8) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(sRes + row * 32 + (lane & 3) * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
      emit(row, lane * 8, v);
    }
  };
  auto pack8 = [](const float (&v)[8]) {
    return make_uint4(float2_to_bf16x2(v[0], v[1]), float2_to_bf16x2(v[2], v[3]), float2_to_bf16x2(v[4], v[5]),
                      float2_to_bf16x2(v[6], v[7]));
  };

  qkv_load(p.L[0]);
  qkv_compute(sBq0, 0);
  grid_arrive(p.grid_bar);

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
    const uint32_t par = static_cast<uint32_t>(layer & 1);
    const float* P = sParam + (layer & 1) * sm::kParamFloats;
    mark_i = 0;
    mark(layer);   // 0: layer start
    if (!last) stage_params(layer + 1);   // (this layer's block was requested one layer ago)
    uint4 bo[8];   // out_proj fragments: requested before the grid barrier
    load_b<8>(bo, W.wo + static_cast<int64_t>(n0 + (warp & 3) * 8 + g) * kC + t * 8, true);
    grid_wait(p.grid_bar, n_cta_u * static_cast<unsigned>(layer + 1));
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
      cp_async_wait<0>();   // (also this layer's parameter block, and the next one's if it is already there)
      __syncthreads();
      mark(layer);   // 2: K/V staged
      constexpr int NS = 4;  // key splits: warp = split * 2 + m-tile
      const int mt = warp & 1, split = warp >> 1;
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
      // merge the four key-range partials of each m-tile
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
      {
        const int row = tid >> 3, c4 = (tid & 7) * 4;   // 256 threads = 32 rows x 8 column quads
        const int mtr = row >> 4, r16 = row & 15;
        float Mx = -INFINITY;
#pragma unroll
        for (int s2 = 0; s2 < NS; ++s2) Mx = fmaxf(Mx, sMl[((s2 * 2 + mtr) * 16 + r16) * 2]);
        float Lsum = 0.0f, acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int s2 = 0; s2 < NS; ++s2) {
          const int w2 = s2 * 2 + mtr;
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
        *reinterpret_cast<uint2*>(sSlab + row * 32 + c4) = pk;
      }
      fence_proxy_async();
      __syncthreads();
      mark(layer);   // 3: attention computed
      send_slab(sm::oTA, E_ATT);
      mbar_wait(bar_local(E_ATT), par);
      mark(layer);   // 4: attention tile received
    }

    // ======================= out_proj + residual -> LayerNorm1 (transformer.py:638-641) =======================
    preln_slab(sTA, bo, P + PF::bo, E_LN1);
    mark(layer);   // 5: out_proj GEMM sent
    // offsets | logits fragments of this head (24 + 12 = 36 projection rows -> 5 n8-tiles x 2 m-tiles = 10 warp
    // tasks: warps 0, 1 take a second one), requested before the LayerNorm exchange completes
    uint4 bf[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int task = warp + kWarps * i;
      if (task < 10) {
        const int n = (task % 5) * 8 + g;  // projection row of this head owned by the lane
        const bool ok = n < 3 * kLP;
        const int src = n < 2 * kLP ? static_cast<int>(rank) * 2 * kLP + n : kCluster * 2 * kLP + static_cast<int>(rank) * kLP + (n - 2 * kLP);
        load_b<8>(bf[i], W.woff + static_cast<int64_t>(ok ? src : 0) * kC + t * 8, ok);
      }
    }
    mbar_wait(bar_local(E_LN1), par);
    layernorm_rows(P + PF::g1, P + PF::be1, [&](int row, int c0, float (&v)[8]) {
      const float4 pa = *reinterpret_cast<const float4*>(sPos + row * kC + c0);
      const float4 pb = *reinterpret_cast<const float4*>(sPos + row * kC + c0 + 4);
      const float q[8] = {v[0] + pa.x, v[1] + pa.y, v[2] + pa.z, v[3] + pa.w, v[4] + pb.x, v[5] + pb.y, v[6] + pb.z, v[7] + pb.w};
      *reinterpret_cast<uint4*>(sTB + row * kPA + c0) = pack8(q);   // cross-attention query = x + pos (:644)
    });
    __syncthreads();
    mark(layer);   // 6: LayerNorm1

    // ======================= MSDeformAttn of head `rank` (transformer.py:268-285) =======================
    uint4 bout[8];   // output_proj fragments: in flight during the gather
    {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int task = warp + kWarps * i;
        if (task < 10) {
          const int j = task % 5, mg = task / 5;
          float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
          mma_tiles<8, 1>(acc, sTB + (mg * 16) * kPA, kPA, bf[i], lane);
          const int c0 = j * 8 + 2 * t;
          // (columns 36..39 read the neighbouring parameters: they are never consumed)
          const float b0 = P[PF::boff + c0], b1 = P[PF::boff + c0 + 1];
          const int r0 = mg * 16 + g;
          *reinterpret_cast<float2*>(sOL + r0 * kPO + c0) = make_float2(acc[0][0] + b0, acc[0][1] + b1);
          *reinterpret_cast<float2*>(sOL + (r0 + 8) * kPO + c0) = make_float2(acc[0][2] + b0, acc[0][3] + b1);
        }
      }
      load_b<8>(bout, W.wout + static_cast<int64_t>(n0 + (warp & 3) * 8 + g) * kC + t * 8, true);
      __syncthreads();
      mark(layer);   // 7: offsets | logits projection
      // gather: warp w owns rows w, w+8, w+16, w+24. Phase 1 (softmax, locations, corners) runs twice with lanes
      // 0..15 / 16..31 on two rows each; phase 2 keeps all 4 x 6 corner rows of a lane in flight at once.
      const __nv_bfloat16* vbase = p.values + static_cast<int64_t>(tile.seq) * p.v_batch_stride + layer * kC + n0;
      int* st_off = sStage + warp * (4 * kNU * 2);
      float* st_w = reinterpret_cast<float*>(st_off + 4 * kNU);
#pragma unroll
      for (int pr = 0; pr < 2; ++pr) {
        const int half = lane >> 4, pl = lane & 15;
        const int item = 2 * pr + half;
        const int row = warp + kWarps * item;
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
          *reinterpret_cast<int4*>(st_off + item * kNU + pl * 4) =
              make_int4(c.pos[0] < 0 ? -1 : c.pos[0] * ps, c.pos[1] < 0 ? -1 : c.pos[1] * ps,
                        c.pos[2] < 0 ? -1 : c.pos[2] * ps, c.pos[3] < 0 ? -1 : c.pos[3] * ps);
          *reinterpret_cast<float4*>(st_w + item * kNU + pl * 4) = make_float4(c.w[0], c.w[1], c.w[2], c.w[3]);
        }
      }
      __syncwarp();
      const int sg = lane >> 2, sub = lane & 3;
      uint4 v[4][6];
      float wv[4][6];
#pragma unroll
      for (int it = 0; it < 4; ++it)
#pragma unroll
        for (int u = 0; u < 6; ++u) {
          const int eo = st_off[it * kNU + sg + 8 * u];
          wv[it][u] = st_w[it * kNU + sg + 8 * u];
          v[it][u] = ldg128_if(vbase + eo + sub * 8, eo >= 0);
        }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
#pragma unroll
        for (int u = 0; u < 6; ++u) fma_bf16x8(acc, v[it][u], wv[it][u]);
        // recursive halving across the 8 sub-groups (msda.cu phase 3): 8 -> 4 -> 2 -> 1 channels per lane
        int ch = 0, n = 8;
#pragma unroll
        for (int o2 = 16; o2 >= 4; o2 >>= 1) {
          n >>= 1;
          const bool up = (lane & o2) != 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i < n) {
              const float send = up ? acc[i] : acc[i + n];
              const float recv = __shfl_xor_sync(0xffffffffu, send, o2);
              acc[i] = (up ? acc[i + n] : acc[i]) + recv;
            }
          }
          ch += up ? n : 0;
        }
        const float hi = __shfl_xor_sync(0xffffffffu, acc[0], 4);  // odd-channel partner
        const int orow = warp + kWarps * it;
        if ((lane & 4) == 0) *reinterpret_cast<uint32_t*>(sSlab + orow * 32 + sub * 8 + ch) = float2_to_bf16x2(acc[0], hi);
      }
      fence_proxy_async();
      __syncthreads();
      mark(layer);   // 8: gather
      send_slab(sm::oTA, E_G);
      mbar_wait(bar_local(E_G), par);
      mark(layer);   // 9: gathered tile received
    }

    // ======================= output_proj + residual -> LayerNorm2 (transformer.py:286, 646-647) =======================
    preln_slab(sTA, bout, P + PF::bout, E_LN2);
    // FFN fragments: linear1 = this CTA's 128 hidden columns (16 n8-tiles, two per warp); linear2 = the K slice of
    // those 128 columns for the output columns [32 w, 32 w + 32) of cluster rank w (4 n8-tiles x 4 k-blocks)
    uint4 b1f[2][8], b2f[4][4];
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
      load_b<8>(b1f[jj], W.w1 + static_cast<int64_t>(static_cast<int>(rank) * kFs + (warp * 2 + jj) * 8 + g) * kC + t * 8, true);
    mbar_wait(bar_local(E_LN2), par);
    layernorm_rows(P + PF::g2, P + PF::be2, [&](int row, int c0, float (&v)[8]) {
      *reinterpret_cast<uint4*>(sTB + row * kPA + c0) = pack8(v);
    });
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      load_b<4>(b2f[jj], W.w2 + static_cast<int64_t>(warp * 32 + jj * 8 + g) * (kFs * kCluster) + static_cast<int>(rank) * kFs + t * 8, true);
    __syncthreads();
    mark(layer);   // 10: output_proj + LayerNorm2

    // ======================= FFN (transformer.py:576-580) =======================
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int nt = warp * 2 + jj;
      float acc[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
      mma_tiles<8, 2>(acc, sTB, kPA, b1f[jj], lane);
      const int col = nt * 8 + 2 * t;
      const float b0 = P[PF::b1 + col], b1 = P[PF::b1 + col + 1];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        *reinterpret_cast<uint32_t*>(sH + (i * 16 + g) * kPH + col) =
            float2_to_bf16x2(fmaxf(acc[i][0] + b0, 0.0f), fmaxf(acc[i][1] + b1, 0.0f));
        *reinterpret_cast<uint32_t*>(sH + (i * 16 + g + 8) * kPH + col) =
            float2_to_bf16x2(fmaxf(acc[i][2] + b0, 0.0f), fmaxf(acc[i][3] + b1, 0.0f));
      }
    }
    __syncthreads();
    mark(layer);   // 11: FFN linear1
    // linear2, split along K: partial[M, 256] = h[:, own 128] . W2[:, own 128]^T; warp w's 32 output columns go to
    // cluster rank w (reduce-scatter through DSMEM, 16 bytes per store)
    {
      float* sRs = reinterpret_cast<float*>(sTB);   // [8 destinations][M][32] fp32 staging (TB, TC are dead here)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float acc[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
        mma_tiles<4, 2>(acc, sH, kPH, b2f[jj], lane);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int c = jj * 8 + 2 * t;
          *reinterpret_cast<float2*>(sRs + (warp * M + i * 16 + g) * 32 + c) = make_float2(acc[i][0], acc[i][1]);
          *reinterpret_cast<float2*>(sRs + (warp * M + i * 16 + g + 8) * 32 + c) = make_float2(acc[i][2], acc[i][3]);
        }
      }
      fence_proxy_async();
      __syncthreads();
      const int row = tid >> 3, dst = tid & 7;
      const uint32_t pb = map_peer(smem_base, static_cast<uint32_t>(dst));
      bulk_to_peer(pb + static_cast<uint32_t>(sm::oPart + ((rank * M + row) * 32) * 4), smem_addr(sRs + (dst * M + row) * 32), 128u,
                   pb + static_cast<uint32_t>(sm::oBars + E_RS * 8));
    }
    mark(layer);   // 12: FFN linear2 partials sent
    mbar_wait(bar_local(E_RS), par);
    {
      // sum of the 8 partial slabs (fixed order: deterministic) + bias + residual -> pre-LN slab, sent to all CTAs
      const int row = tid >> 3, c4 = (tid & 7) * 4;
      float4 s4 = *reinterpret_cast<const float4*>(P + PF::b2 + c4);
      const float4 r4 = *reinterpret_cast<const float4*>(sRes + row * 32 + c4);
      s4.x += r4.x; s4.y += r4.y; s4.z += r4.z; s4.w += r4.w;
#pragma unroll
      for (int src = 0; src < kCluster; ++src) {
        const float4 q4 = *reinterpret_cast<const float4*>(sPart + (src * M + row) * 32 + c4);
        s4.x += q4.x; s4.y += q4.y; s4.z += q4.z; s4.w += q4.w;
      }
      *reinterpret_cast<float4*>(sYs + row * 32 + c4) = s4;
    }
    fence_proxy_async();
    __syncthreads();
    send_preln(E_LN3);
    mark(layer);   // 13: reduce-scatter received, pre-LN slab sent
    uint4 bh1[8];
    load_b<8>(bh1, W.wb1 + static_cast<int64_t>(n0 + (warp & 3) * 8 + g) * kC + t * 8, true);
    if (!last) qkv_load(p.L[layer + 1]);   // next in-projection fragments: in flight across the LayerNorm exchange
    mbar_wait(bar_local(E_LN3), par);
    layernorm_rows(P + PF::g3, P + PF::be3, [&](int row, int c0, float (&v)[8]) {
      *reinterpret_cast<uint4*>(sTC + row * kPA + c0) = pack8(v);             // x3: operand of v, box head, scores
      const float4 pa = *reinterpret_cast<const float4*>(sPos + row * kC + c0);
      const float4 pb = *reinterpret_cast<const float4*>(sPos + row * kC + c0 + 4);
      const float q[8] = {v[0] + pa.x, v[1] + pa.y, v[2] + pa.z, v[3] + pa.w, v[4] + pb.x, v[5] + pb.y, v[6] + pb.z, v[7] + pb.w};
      *reinterpret_cast<uint4*>(sTB + row * kPA + c0) = pack8(q);             // x3 + pos: operand of q, k
    });
    __syncthreads();
    mark(layer);   // 14: LayerNorm3

    if (last) {
      // output embedding (fp32 + optional bf16) and the class-score head on the bf16-rounded row
      for (int row = warp; row < tile.n; row += kWarps) {
        const int64_t gr = static_cast<int64_t>(tile.row0 + row) * kC + n0 + lane;
        p.x_out[gr] = sRes[row * 32 + lane];
        if (p.x_lp_out != nullptr) p.x_lp_out[gr] = sTC[row * kPA + n0 + lane];
      }
      if (p.nc > 0) {
        for (int row = static_cast<int>(rank) + kCluster * warp; row < tile.n; row += kCluster * kWarps) {
          const uint4 xv = *reinterpret_cast<const uint4*>(sTC + row * kPA + lane * 8);
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
      qkv_compute(P + PF::bqkv, layer + 1);
    }
    // ======================= box head + refinement (transformer.py:709) =======================
    relu_slab(sTC, bh1, P + PF::bb1);
    fence_proxy_async();
    if (!last) grid_arrive(p.grid_bar); else __syncthreads();   // (the barrier's __syncthreads also completes sSlab)
    mark(layer);   // 15: next in-projection + box-head layer 1 (+ grid arrive)
    if (tid == 0 && !last) {   // next layer's expectations for the six exchanges that have completed in this layer
      mbar_arm(bar_local(E_ATT), kBf16TileBytes);
      mbar_arm(bar_local(E_LN1), kF32TileBytes);
      mbar_arm(bar_local(E_G), kBf16TileBytes);
      mbar_arm(bar_local(E_LN2), kF32TileBytes);
      mbar_arm(bar_local(E_RS), kF32TileBytes);
      mbar_arm(bar_local(E_LN3), kF32TileBytes);
    }
    send_slab(sm::oTD, E_H1);
    uint4 bh2[8];
    load_b<8>(bh2, W.wb2 + static_cast<int64_t>(n0 + (warp & 3) * 8 + g) * kC + t * 8, true);
    mbar_wait(bar_local(E_H1), par);
    __syncthreads();   // every warp has sent its part of sSlab before it is overwritten
    if (tid == 0 && !last) mbar_arm(bar_local(E_H1), kBf16TileBytes);
    relu_slab(sTD, bh2, P + PF::bb2);
    fence_proxy_async();
    __syncthreads();
    send_slab(sm::oTC, E_H2);
    float w3[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(W.wb3 + j * kC + lane * 8));
      const float4 b = __ldg(reinterpret_cast<const float4*>(W.wb3 + j * kC + lane * 8 + 4));
      w3[j][0] = a.x; w3[j][1] = a.y; w3[j][2] = a.z; w3[j][3] = a.w;
      w3[j][4] = b.x; w3[j][5] = b.y; w3[j][6] = b.z; w3[j][7] = b.w;
    }
    mbar_wait(bar_local(E_H2), par);
    {
      const float bias = lane < 4 ? P[PF::bb3 + lane] : 0.0f;
      float* rout = p.refer_out[layer];
      for (int row = warp; row < M; row += kWarps) {
        const uint4 hv = *reinterpret_cast<const uint4*>(sTC + row * kPA + lane * 8);
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
      if (tid == 0 && !last) mbar_arm(bar_local(E_H2), kBf16TileBytes);
    }
  }
  cluster_sync();  // no CTA may exit while a peer can still address its shared memory
}

}  // namespace dc
}  // namespace moyolo

using namespace moyolo;

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
namespace {

int dc_configure(int* max_clusters, size_t* smem_out) {
  static DeviceOnce once;
  static int max_cl[64];
  const int dev = DeviceOnce::current();
  const size_t smem = dc::sm::kTotal;
  if (!once.done(dev)) {
    cudaError_t e = cudaFuncSetAttribute(dc::decoder_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return fail(MOYOLO_ERR_CUDA, "cudaFuncSetAttribute(decoder_cluster smem=%zu): %s", smem, cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(dc::kCluster * 18);
    cfg.blockDim = dim3(dc::kThreads);
    cfg.dynamicSmemBytes = smem;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, dc::decoder_cluster_kernel, &cfg);
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
  MOYOLO_REQUIRE(rows_per_tile == dc::M, MOYOLO_ERR_BAD_ARG, "decoder_cluster: rows_per_tile must be %d", dc::M);
  size_t smem = 0;
  int mc = 0;
  const int rc = dc_configure(&mc, &smem);
  if (rc != MOYOLO_OK) return rc;
  if (max_clusters) *max_clusters = mc;
  if (kv_cap) *kv_cap = dc::sm::kKvCap;
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
  MOYOLO_REQUIRE(a->rows_per_tile == dc::M, MOYOLO_ERR_BAD_ARG, "decoder_cluster: rows_per_tile must be %d", dc::M);
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
  rc = dc_configure(&max_clusters, &smem);
  if (rc != MOYOLO_OK) return rc;
  // tiles <= rows_pad / M + n_seq - 1 (every sequence may end with a partial tile)
  const int64_t tiles_bound = (a->rows_pad + dc::M - 1) / dc::M + (a->n_seq - 1);
  MOYOLO_REQUIRE(tiles_bound <= max_clusters, MOYOLO_ERR_UNSUPPORTED,
                 "decoder_cluster: %lld row tiles need more than the %d co-resident clusters of this device",
                 (long long)tiles_bound, max_clusters);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->reset_barrier) {
    cudaError_t e = cudaMemsetAsync(a->grid_barrier, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return fail(MOYOLO_ERR_CUDA, "decoder_cluster: memset of the grid barrier: %s", cudaGetErrorString(e));
  }
  const dim3 grid(static_cast<unsigned>(tiles_bound) * dc::kCluster);
  launch_k(dc::decoder_cluster_kernel, grid, dim3(dc::kThreads), smem, st, p);
  return check_launch("decoder_cluster_kernel");
}
