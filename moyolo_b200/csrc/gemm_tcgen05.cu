// tcgen05 / TMA / TMEM GEMMs for sm_100a:  y[M,N] = act(x[M,K] . w[N,K]^T + bias), bf16 operands,
// fp32 accumulation in tensor memory. Serves every dense contraction of the decoder hot path
// (value_proj for all layers at once, sampling_offsets|attention_weights, output_proj, MHA in/out
// projections, FFN, bbox-MLP hidden layers; nn.Linear call sites of
// ultralytics/nn/modules/transformer.py:264,268,269,286,576-580,638 and MOTR/models/qim.py:276-290).
//
// Both operands are K-major (x rows and nn.Linear's [out,in] weight rows are contiguous in K), so no
// transposes are needed: TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) stages [128 x 64] x-tiles and
// [BN x 64] w-tiles in shared memory; one elected thread issues tcgen05.mma.cta_group::1.kind::f16
// (M=128, N=BN, K=16); accumulators live in TMEM; epilogue warps read them back with tcgen05.ld
// (one output row per thread).
//
// Two kernels:
//  * gemm_tcgen05_kernel<BN, TO, LN> — one output tile per CTA, tuned for LATENCY (the few-hundred-row
//    query GEMMs of a frame are a chain of short dependent launches): weight tiles are requested
//    before the programmatic-dependency wait, activation tiles right after it, bias/gamma/beta and the
//    residual rows are prefetched while the MMAs run. Options: two A operands split at a column
//    (q,k from x+pos and v from x in ONE launch, transformer.py:637-638), and LN=true: the
//    residual-add + LayerNorm (+ "+pos" operand of the next GEMM) of the post-norm blocks
//    (transformer.py:640-641, 646-647, 578-579) fused into the epilogue. A LayerNorm row spans the
//    256/BN CTAs of a thread-block cluster; row statistics are exchanged through distributed shared
//    memory (per-slab mean / centred sum of squares, sent with st.async stores that complete an mbarrier in the
//    receiving CTA, merged with the parallel-variance formula: one exchange, no cluster-wide barrier after start-up).
//  * gemm_stream_kernel — persistent, weight-resident kernel for the tall value projection
//    (M = S*Lv ~ 10^4..10^5 rows, K = 256): every CTA keeps its [128 x 256] weight slab in shared
//    memory, streams x-tiles through a 6-deep TMA ring, double-buffers the accumulator in TMEM so the
//    epilogue of tile i overlaps the MMAs of tile i+1, and writes bf16 through swizzled shared-memory
//    slabs with TMA stores (full-line writes instead of one row per thread).
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include <atomic>

#include "common.cuh"

namespace moyolo {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int kStages = 4;
constexpr int kGemmThreads = 192;
constexpr int kLnCols = 256;  // LayerNorm width of the fused epilogue (d_model)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4, [16,30) LBO>>4 (unused for swizzled K-major), [32,46) SBO>>4 = 1024B between
// 8-row groups, [46,48) version = 1, [61,64) layout = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6),
// a/b format BF16 (1) at [7,10)/[10,13), a/b K-major (0) at 15/16, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- thread-block cluster helpers (fused LayerNorm epilogue) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_f32(uint32_t local_addr, uint32_t rank, float v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

// Asynchronous store into a peer CTA's shared memory that signals (complete_tx) an mbarrier in that CTA when it has
// landed: the receiver waits on its own mbarrier instead of a cluster-wide barrier.
__device__ __forceinline__ void st_async_cluster_f32(uint32_t local_addr, uint32_t local_bar, uint32_t rank, float v) {
  uint32_t raddr, rbar;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_addr), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(local_bar), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr),
               "r"(__float_as_uint(v)), "r"(rbar)
               : "memory");
}

// Everything the epilogues need besides the tensor maps.
struct GemmEpi {
  const float* bias;
  void* y;
  int64_t ldy;
  int64_t M;
  int N;
  int K;
  int relu;
  const uint8_t* zero_rows;
  int n_split;  // output columns >= n_split take their A operand from the second tensor map
  // LN epilogue (N == 256): out = LayerNorm(acc + bias + residual) * gamma + beta
  const float* residual;
  const float* gamma;
  const float* beta;
  float eps;
  float* out_f32;
  void* out_lp;
  const float* pos;
  void* out_pos_lp;
  // optional class-score head fused behind the LayerNorm (transformer.py:717-721): logits = bf16(out) . score_w^T + b
  const float* score_w;   // fp32 [score_nc, 256] or NULL
  const float* score_b;
  int score_nc;           // 0 = no score head; <= kMaxScoreNc
  float* logits;          // [M, score_nc]
  float* scores;          // [M] sigmoid(max logit)
  int32_t* labels;        // [M] argmax (first maximum)
};

constexpr int kMaxScoreNc = 8;

template <int BN, bool LN>
struct GemmCtl {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full;
  uint64_t ln_bar;  // LN: completes when the row statistics of all cluster peers have landed in `red`
  uint64_t sc_bar;  // LN + score head, cluster rank 0: completes when every peer's partial class scores have landed
  uint32_t tmem_base;
  float bias[BN];
  float gamma[LN ? BN : 1];
  float beta[LN ? BN : 1];
  float red[LN ? 2 : 1][LN ? kLnCols / BN : 1][LN ? kBM : 1];  // [mean | M2][peer CTA][row]
};

template <typename TO>
__device__ __forceinline__ void store16(TO* dst, const float (&v)[16], bool vec_ok) {
  if (vec_ok) {
    if constexpr (sizeof(TO) == 2) {
      uint4 a, b;
      a.x = float2_to_bf16x2(v[0], v[1]);   a.y = float2_to_bf16x2(v[2], v[3]);
      a.z = float2_to_bf16x2(v[4], v[5]);   a.w = float2_to_bf16x2(v[6], v[7]);
      b.x = float2_to_bf16x2(v[8], v[9]);   b.y = float2_to_bf16x2(v[10], v[11]);
      b.z = float2_to_bf16x2(v[12], v[13]); b.w = float2_to_bf16x2(v[14], v[15]);
      reinterpret_cast<uint4*>(dst)[0] = a;
      reinterpret_cast<uint4*>(dst)[1] = b;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) dst[j] = from_float<TO>(v[j]);
  }
}

template <int BN, typename TO, bool LN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_x2,
                    const __grid_constant__ CUtensorMap tmap_w, const GemmEpi e) {
  constexpr uint32_t kABytes = kBM * kBK * 2;
  constexpr uint32_t kBBytes = BN * kBK * 2;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;  // power of two >= 32 (BN in {32,64,128,256})
  constexpr int NC = kLnCols / BN;                   // CTAs per LayerNorm row (cluster size) when LN
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte align the dynamic shared memory window (SWIZZLE_128B atoms)
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = base;
  uint8_t* smem_b = base + kStages * kABytes;
  using Ctl = GemmCtl<BN, LN>;
  Ctl* ctl = reinterpret_cast<Ctl*>(smem_b + kStages * kBBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kBM;
  const int n0 = blockIdx.x * BN;
  const int num_kb = e.K / kBK;
  const int pre = num_kb < kStages ? num_kb : kStages;  // k-blocks whose stage is free at start

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&ctl->full[s]), 2);  // one arrive.expect_tx per operand
      mbar_init(smem_u32(&ctl->empty[s]), 1);
    }
    mbar_init(smem_u32(&ctl->tmem_full), 1);
    if constexpr (LN) {
      mbar_init(smem_u32(&ctl->ln_bar), 1);
      mbar_init(smem_u32(&ctl->sc_bar), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // LN: every peer sends 2 floats per row (its slab's mean and centred sum of squares)
    if constexpr (LN) mbar_expect_tx(smem_u32(&ctl->ln_bar), (kLnCols / BN) * kBM * 2 * 4);
    if constexpr (LN) {
      if (e.score_nc > 0 && cluster_ctarank() == 0)
        mbar_expect_tx(smem_u32(&ctl->sc_bar), static_cast<uint32_t>(e.score_nc) * (kLnCols / BN) * kBM * 4);
    }
    // weights are immutable during a frame: request them before waiting on the previous kernel
    for (int kb = 0; kb < pre; ++kb) {
      const uint32_t full = smem_u32(&ctl->full[kb]);
      mbar_expect_tx(full, kBBytes);
      tma_load_2d(smem_u32(smem_b + kb * kBBytes), &tmap_w, full, kb * kBK, n0);
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {  // per-column epilogue constants (immutable)
    for (int i = threadIdx.x - 64; i < BN; i += kGemmThreads - 64) {
      const bool in = n0 + i < e.N;
      ctl->bias[i] = (e.bias != nullptr && in) ? __ldg(e.bias + n0 + i) : 0.0f;
      if constexpr (LN) {
        ctl->gamma[i] = __ldg(e.gamma + n0 + i);
        ctl->beta[i] = __ldg(e.beta + n0 + i);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();  // TMEM is allocated: the next kernel may start its own prologue
  if constexpr (LN) cluster_arrive();  // #0: peers are running (required before any DSMEM access)
  const uint32_t tmem_acc = ctl->tmem_base;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const CUtensorMap* ta = (n0 >= e.n_split) ? &tmap_x2 : &tmap_x;
      pdl_wait();  // activations come from the previous kernel
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t full = smem_u32(&ctl->full[s]);
        if (kb >= pre) {
          mbar_wait(smem_u32(&ctl->empty[s]), ((kb / kStages) & 1) ^ 1);
          mbar_expect_tx(full, kBBytes);
          tma_load_2d(smem_u32(smem_b + s * kBBytes), &tmap_w, full, kb * kBK, n0);
        }
        mbar_expect_tx(full, kABytes);
        tma_load_2d(smem_u32(smem_a + s * kABytes), ta, full, kb * kBK, m0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kBM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        mbar_wait(smem_u32(&ctl->full[s]), (kb / kStages) & 1);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc(smem_u32(smem_a + s * kABytes));
        const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + s * kBBytes));
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in 16-byte units
          umma_bf16(tmem_acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&ctl->empty[s]));  // frees the smem stage when these MMAs retire
      }
      umma_commit(smem_u32(&ctl->tmem_full));   // accumulator complete
    }
    __syncwarp();
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4, one output row per thread =====
    const int quad = warp & 3;
    const int rl = quad * 32 + lane;  // row inside the tile
    const int64_t row = static_cast<int64_t>(m0) + rl;
    const bool row_ok = row < e.M;
    const uint32_t tbase = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16);
    pdl_wait();
    if constexpr (!LN) {
      const bool zero = row_ok && e.zero_rows != nullptr && e.zero_rows[row] != 0;
      TO* yrow = static_cast<TO*>(e.y) + row * e.ldy + n0;
      const bool vec_ok = (reinterpret_cast<uintptr_t>(e.y) & 15u) == 0 && (e.ldy * sizeof(TO)) % 16 == 0;
      mbar_wait(smem_u32(&ctl->tmem_full), 0);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r0[16], r1[16];
        tmem_ld16(tbase + c0, r0);
        tmem_ld16(tbase + c0 + 16, r1);
        tmem_ld_wait();
        if (!row_ok) continue;
        float v[16];
        if (n0 + c0 < e.N) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float t = __uint_as_float(r0[j]) + ctl->bias[c0 + j];
            if (e.relu) t = fmaxf(t, 0.0f);
            v[j] = zero ? 0.0f : t;
          }
          store16<TO>(yrow + c0, v, vec_ok);
        }
        if (n0 + c0 + 16 < e.N) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float t = __uint_as_float(r1[j]) + ctl->bias[c0 + 16 + j];
            if (e.relu) t = fmaxf(t, 0.0f);
            v[j] = zero ? 0.0f : t;
          }
          store16<TO>(yrow + c0 + 16, v, vec_ok);
        }
      }
    } else {
      static_assert(!LN || BN == 32, "the fused LayerNorm epilogue holds one 32-column slab per thread");
      const uint32_t my_rank = cluster_ctarank();
      // residual row slab (and the +pos operand) prefetched while the MMAs run
      float v[32], pv[32];
      const int64_t goff = row * kLnCols + n0;
      const bool want_pos = e.out_pos_lp != nullptr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f), p4 = r4;
        if (row_ok && e.residual != nullptr) r4 = *reinterpret_cast<const float4*>(e.residual + goff + 4 * j);
        if (row_ok && want_pos) p4 = *reinterpret_cast<const float4*>(e.pos + goff + 4 * j);
        v[4 * j] = r4.x; v[4 * j + 1] = r4.y; v[4 * j + 2] = r4.z; v[4 * j + 3] = r4.w;
        pv[4 * j] = p4.x; pv[4 * j + 1] = p4.y; pv[4 * j + 2] = p4.z; pv[4 * j + 3] = p4.w;
      }
      mbar_wait(smem_u32(&ctl->tmem_full), 0);
      tc_fence_after();
      {
        uint32_t r0[16], r1[16];
        tmem_ld16(tbase, r0);
        tmem_ld16(tbase + 16, r1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          v[j] += __uint_as_float(r0[j]) + ctl->bias[j];
          v[16 + j] += __uint_as_float(r1[j]) + ctl->bias[16 + j];
        }
      }
      // Row statistics over the 256 columns held by the NC CTAs of the cluster in ONE exchange: every CTA
      // computes the mean and the centred sum of squares of its own 32-column slab (two passes, in registers)
      // and the slabs are merged with the parallel-variance formula (Chan et al.):
      //   mean = sum_i m_i / NC,   M2 = sum_i M2_i + BN * sum_i (m_i - mean)^2
      // -- as robust as a global two-pass, one cluster barrier and one DSMEM round fewer.
      float ps = 0.0f;
#pragma unroll
      for (int j = 0; j < 32; ++j) ps += v[j];
      const float m_loc = ps * (1.0f / BN);
      float pq = 0.0f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float d = v[j] - m_loc;
        pq = fmaf(d, d, pq);
      }
      cluster_wait();  // #0: every peer has initialised its barrier
      const uint32_t slot0 = smem_u32(&ctl->red[0][my_rank][rl]);
      const uint32_t slot1 = smem_u32(&ctl->red[1][my_rank][rl]);
      const uint32_t lbar = smem_u32(&ctl->ln_bar);
#pragma unroll
      for (int p = 0; p < NC; ++p) {
        st_async_cluster_f32(slot0, lbar, p, m_loc);
        st_async_cluster_f32(slot1, lbar, p, pq);
      }
      mbar_wait(lbar, 0);  // all NC * 128 * 2 values of the peers (and our own) have landed
      float msum = 0.0f, sq = 0.0f;
#pragma unroll
      for (int p = 0; p < NC; ++p) {
        msum += ctl->red[0][p][rl];
        sq += ctl->red[1][p][rl];
      }
      const float mean = msum * (1.0f / NC);
#pragma unroll
      for (int p = 0; p < NC; ++p) {
        const float d = ctl->red[0][p][rl] - mean;
        sq = fmaf(static_cast<float>(BN) * d, d, sq);
      }
      const float rstd = rsqrtf(sq * (1.0f / kLnCols) + e.eps);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] -= mean;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = v[j] * rstd * ctl->gamma[j] + ctl->beta[j];
      if (row_ok) {
        if (e.out_f32 != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(e.out_f32 + goff + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (e.out_lp != nullptr) {
          TO* d = static_cast<TO*>(e.out_lp) + goff;
          float h0[16], h1[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { h0[j] = v[j]; h1[j] = v[16 + j]; }
          store16<TO>(d, h0, true);
          store16<TO>(d + 16, h1, true);
        }
        if (want_pos) {
          TO* d = static_cast<TO*>(e.out_pos_lp) + goff;
          float h0[16], h1[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { h0[j] = v[j] + pv[j]; h1[j] = v[16 + j] + pv[16 + j]; }
          store16<TO>(d, h0, true);
          store16<TO>(d + 16, h1, true);
        }
      }
      if (e.score_nc > 0) {
        // class-score head on the bf16-rounded output row (what the stand-alone score_head kernel reads): this CTA's
        // partial dot products over its 32 columns go to cluster rank 0, which adds the bias and writes the logits,
        // sigmoid(max logit) and the arg-max label of the row
        float* sred = reinterpret_cast<float*>(ctl + 1);   // [score_nc][NC][kBM], dynamic shared memory behind ctl
        const uint32_t sbar = smem_u32(&ctl->sc_bar);
        for (int c = 0; c < e.score_nc; ++c) {
          const float* wc = e.score_w + c * kLnCols + n0;
          float sdot = 0.0f;
#pragma unroll
          for (int j = 0; j < 32; ++j) sdot = fmaf(__bfloat162float(__float2bfloat16_rn(v[j])), __ldg(wc + j), sdot);
          st_async_cluster_f32(smem_u32(sred + (c * NC + static_cast<int>(my_rank)) * kBM + rl), sbar, 0, row_ok ? sdot : 0.0f);
        }
        if (my_rank == 0) {
          mbar_wait(sbar, 0);
          float best = -INFINITY;
          int best_c = 0;
          for (int c = 0; c < e.score_nc; ++c) {
            float d = 0.0f;
#pragma unroll
            for (int p = 0; p < NC; ++p) d += sred[(c * NC + p) * kBM + rl];
            d += __ldg(e.score_b + c);
            if (row_ok && e.logits != nullptr) e.logits[row * e.score_nc + c] = d;
            if (d > best) { best = d; best_c = c; }  // first maximum wins, as torch.max
          }
          if (row_ok) {
            if (e.scores != nullptr) e.scores[row] = 1.0f / (1.0f + expf(-best));
            if (e.labels != nullptr) e.labels[row] = best_c;
          }
        }
      }
    }
  }
  if constexpr (LN) {
    if (warp < 2) cluster_wait();  // producer / MMA warps: the start-up barrier only
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// The whole post-norm FFN block in ONE launch (transformer.py:576-580; QIM qim.py:280-282, 290-298):
//   out = LayerNorm(residual + W2 . relu(W1 . x + b1) + b2)
// A thread-block cluster of 8 CTAs owns 128 rows. Phase 1: CTA r computes the hidden slab
// H[:, r*BN1 : (r+1)*BN1] = relu(x . W1_r^T + b1_r) (tcgen05, accumulator in TMEM) and writes it as bf16 to the
// global scratch `h`. A cluster barrier (release/acquire, with generic->async proxy fences around it) makes the
// eight slabs visible; phase 2 streams the full H row tile [128 x F] back through TMA as the A operand of the
// second GEMM (CTA r: output columns [32r, 32r+32), its W2 slab resident in shared memory since before the
// dependency wait), and the epilogue is the cluster LayerNorm of gemm_tcgen05_kernel<32, ., true>.
// Saves one dependent launch per FFN block; the arithmetic (operand rounding, K order) is that of the two-launch
// path, so results are bit-identical to it.
// ------------------------------------------------------------------------------------------------
template <int BN1>
struct FfnCtl {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t w2_full;
  uint64_t acc1_full;
  uint64_t acc2_full;
  uint32_t tmem_base;
  float bias1[BN1];
  float bias2[32];
  float gamma[32];
  float beta[32];
  float red[2][kLnCols / 32][kBM];  // [mean | M2][peer CTA][row]
};

struct FfnArgs {
  const float* b1;
  const float* b2;
  void* h;          // bf16 [M, F] scratch (row stride F)
  int F;            // hidden width = 8 * BN1
  int64_t M;
  const float* residual;
  const float* gamma;
  const float* beta;
  float eps;
  float* out_f32;
  void* out_lp;
  const float* pos;
  void* out_pos_lp;
};

template <int BN1>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_ffn_ln_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                   const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_w2,
                   const FfnArgs e) {
  constexpr int NC = kLnCols / 32;                   // 8 CTAs per cluster
  constexpr uint32_t kABytes = kBM * kBK * 2;        // 16 KiB
  constexpr uint32_t kB1Bytes = BN1 * kBK * 2;       // W1 k-block of this CTA's hidden slab
  constexpr uint32_t kB2Bytes = 32 * kBK * 2;        // W2 k-block of this CTA's 32 output columns (4 KiB)
  constexpr uint32_t kTmemCols = BN1 + 32 <= 64 ? 64 : 256;
  constexpr int kKB1 = kLnCols / kBK;                // K of phase 1 = d_model = 256 -> 4 k-blocks
  static_assert(kKB1 == kStages, "phase 1 fills the ring exactly once");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = base;                                   // ring: x tiles (phase 1), then H tiles (phase 2)
  uint8_t* smem_b1 = smem_a + kStages * kABytes;            // W1 slab, 4 k-blocks
  uint8_t* smem_w2 = smem_b1 + kStages * kB1Bytes;          // W2 slab, F/64 k-blocks, resident
  const int num_kb2 = e.F / kBK;
  using Ctl = FfnCtl<BN1>;
  Ctl* ctl = reinterpret_cast<Ctl*>(smem_w2 + static_cast<size_t>(num_kb2) * kB2Bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t my_rank = cluster_ctarank();
  const int m0 = blockIdx.y * kBM;
  const int n1 = static_cast<int>(my_rank) * BN1;   // hidden columns of phase 1
  const int n2 = static_cast<int>(my_rank) * 32;    // output columns of phase 2

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w1)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w2)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_h)) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&ctl->full[s]), 2);
      mbar_init(smem_u32(&ctl->empty[s]), 1);
    }
    mbar_init(smem_u32(&ctl->w2_full), 1);
    mbar_init(smem_u32(&ctl->acc1_full), 1);
    mbar_init(smem_u32(&ctl->acc2_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // weights are immutable during a frame: both slabs are requested before waiting on the previous kernel
    for (int kb = 0; kb < kKB1; ++kb) {
      const uint32_t full = smem_u32(&ctl->full[kb]);
      mbar_expect_tx(full, kB1Bytes);
      tma_load_2d(smem_u32(smem_b1 + kb * kB1Bytes), &tmap_w1, full, kb * kBK, n1);
    }
    const uint32_t wf = smem_u32(&ctl->w2_full);
    mbar_expect_tx(wf, static_cast<uint32_t>(num_kb2) * kB2Bytes);
    for (int kb = 0; kb < num_kb2; ++kb) tma_load_2d(smem_u32(smem_w2 + kb * kB2Bytes), &tmap_w2, wf, kb * kBK, n2);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < BN1; i += kGemmThreads - 64) ctl->bias1[i] = e.b1 ? __ldg(e.b1 + n1 + i) : 0.0f;
    if (threadIdx.x - 64 < 32) {
      const int i = threadIdx.x - 64;
      ctl->bias2[i] = e.b2 ? __ldg(e.b2 + n2 + i) : 0.0f;
      ctl->gamma[i] = __ldg(e.gamma + n2 + i);
      ctl->beta[i] = __ldg(e.beta + n2 + i);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  const uint32_t tmem_acc1 = ctl->tmem_base;         // BN1 columns
  const uint32_t tmem_acc2 = ctl->tmem_base + BN1;   // 32 columns

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      pdl_wait();
      for (int kb = 0; kb < kKB1; ++kb) {  // phase 1: x tiles (the W1 halves of these stages are already in flight)
        const uint32_t full = smem_u32(&ctl->full[kb]);
        mbar_expect_tx(full, kABytes);
        tma_load_2d(smem_u32(smem_a + kb * kABytes), &tmap_x, full, kb * kBK, m0);
      }
    }
    __syncwarp();
    cluster_arrive();  // #H: every CTA of the cluster has written its hidden slab
    cluster_wait();
    if (lane == 0) {
      asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes of H (acquired above) -> async-proxy reads
      for (int kb = 0; kb < num_kb2; ++kb) {  // phase 2: H tiles; W2 is resident
        const int it = kKB1 + kb, s = it % kStages;
        const uint32_t full = smem_u32(&ctl->full[s]);
        mbar_wait(smem_u32(&ctl->empty[s]), ((it / kStages) & 1) ^ 1);
        mbar_arrive(full);  // the stage's second arrival (no B operand to wait for in this phase)
        mbar_expect_tx(full, kABytes);
        tma_load_2d(smem_u32(smem_a + s * kABytes), &tmap_h, full, kb * kBK, m0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(kBM, BN1);
      for (int kb = 0; kb < kKB1; ++kb) {
        mbar_wait(smem_u32(&ctl->full[kb]), 0);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc(smem_u32(smem_a + kb * kABytes));
        const uint64_t bdesc = make_smem_desc(smem_u32(smem_b1 + kb * kB1Bytes));
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) umma_bf16(tmem_acc1, adesc + 2 * k, bdesc + 2 * k, idesc1, (kb | k) != 0 ? 1u : 0u);
        umma_commit(smem_u32(&ctl->empty[kb]));
      }
      umma_commit(smem_u32(&ctl->acc1_full));
    }
    __syncwarp();
    cluster_arrive();  // #H
    cluster_wait();
    if (lane == 0) {
      constexpr uint32_t idesc2 = make_idesc(kBM, 32);
      mbar_wait(smem_u32(&ctl->w2_full), 0);
      for (int kb = 0; kb < num_kb2; ++kb) {
        const int it = kKB1 + kb, s = it % kStages;
        mbar_wait(smem_u32(&ctl->full[s]), (it / kStages) & 1);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc(smem_u32(smem_a + s * kABytes));
        const uint64_t bdesc = make_smem_desc(smem_u32(smem_w2 + kb * kB2Bytes));
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) umma_bf16(tmem_acc2, adesc + 2 * k, bdesc + 2 * k, idesc2, (kb | k) != 0 ? 1u : 0u);
        umma_commit(smem_u32(&ctl->empty[s]));
      }
      umma_commit(smem_u32(&ctl->acc2_full));
    }
    __syncwarp();
  } else {
    // ===== epilogue warps 2..5: TMEM lane quadrant = warp % 4, one row per thread =====
    const int quad = warp & 3;
    const int rl = quad * 32 + lane;
    const int64_t row = static_cast<int64_t>(m0) + rl;
    const bool row_ok = row < e.M;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    pdl_wait();
    // ---- phase 1 epilogue: hidden slab -> bias + ReLU -> bf16 -> global scratch ----
    mbar_wait(smem_u32(&ctl->acc1_full), 0);
    tc_fence_after();
    {
      __nv_bfloat16* hrow = static_cast<__nv_bfloat16*>(e.h) + row * e.F + n1;
#pragma unroll 1
      for (int c0 = 0; c0 < BN1; c0 += 32) {
        uint32_t r0[16], r1[16];
        tmem_ld16(tmem_acc1 + lane_base + c0, r0);
        tmem_ld16(tmem_acc1 + lane_base + c0 + 16, r1);
        tmem_ld_wait();
        if (!row_ok) continue;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(__uint_as_float(r0[j]) + ctl->bias1[c0 + j], 0.0f);
        store16<__nv_bfloat16>(hrow + c0, v, true);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(__uint_as_float(r1[j]) + ctl->bias1[c0 + 16 + j], 0.0f);
        store16<__nv_bfloat16>(hrow + c0 + 16, v, true);
      }
    }
    asm volatile("fence.proxy.async;" ::: "memory");  // H is read back through TMA (async proxy) by every CTA
    tc_fence_before();
    cluster_arrive();  // #H (release: the slab is visible to the cluster)
    // residual row slab (and the +pos operand) prefetched while the peers finish and phase 2 runs
    float v[32], pv[32];
    const int64_t goff = row * kLnCols + n2;
    const bool want_pos = e.out_pos_lp != nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f), p4 = r4;
      if (row_ok && e.residual != nullptr) r4 = *reinterpret_cast<const float4*>(e.residual + goff + 4 * j);
      if (row_ok && want_pos) p4 = *reinterpret_cast<const float4*>(e.pos + goff + 4 * j);
      v[4 * j] = r4.x; v[4 * j + 1] = r4.y; v[4 * j + 2] = r4.z; v[4 * j + 3] = r4.w;
      pv[4 * j] = p4.x; pv[4 * j + 1] = p4.y; pv[4 * j + 2] = p4.z; pv[4 * j + 3] = p4.w;
    }
    cluster_wait();
    // ---- phase 2 epilogue: + bias + residual, cluster LayerNorm ----
    mbar_wait(smem_u32(&ctl->acc2_full), 0);
    tc_fence_after();
    {
      uint32_t r0[16], r1[16];
      tmem_ld16(tmem_acc2 + lane_base, r0);
      tmem_ld16(tmem_acc2 + lane_base + 16, r1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        v[j] += __uint_as_float(r0[j]) + ctl->bias2[j];
        v[16 + j] += __uint_as_float(r1[j]) + ctl->bias2[16 + j];
      }
    }
    float ps = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; ++j) ps += v[j];
    const float m_loc = ps * (1.0f / 32);
    float pq = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float d = v[j] - m_loc;
      pq = fmaf(d, d, pq);
    }
    const uint32_t slot0 = smem_u32(&ctl->red[0][my_rank][rl]);
    const uint32_t slot1 = smem_u32(&ctl->red[1][my_rank][rl]);
#pragma unroll
    for (int p = 0; p < NC; ++p) {
      st_cluster_f32(slot0, p, m_loc);
      st_cluster_f32(slot1, p, pq);
    }
    cluster_arrive();
    cluster_wait();  // #LN
    float msum = 0.0f, sq = 0.0f;
#pragma unroll
    for (int p = 0; p < NC; ++p) {
      msum += ctl->red[0][p][rl];
      sq += ctl->red[1][p][rl];
    }
    const float mean = msum * (1.0f / NC);
#pragma unroll
    for (int p = 0; p < NC; ++p) {
      const float d = ctl->red[0][p][rl] - mean;
      sq = fmaf(32.0f * d, d, sq);
    }
    const float rstd = rsqrtf(sq * (1.0f / kLnCols) + e.eps);
    if (row_ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (v[j] - mean) * rstd * ctl->gamma[j] + ctl->beta[j];
      if (e.out_f32 != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(e.out_f32 + goff + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      if (e.out_lp != nullptr) {
        __nv_bfloat16* d = static_cast<__nv_bfloat16*>(e.out_lp) + goff;
        float h0[16], h1[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { h0[j] = v[j]; h1[j] = v[16 + j]; }
        store16<__nv_bfloat16>(d, h0, true);
        store16<__nv_bfloat16>(d + 16, h1, true);
      }
      if (want_pos) {
        __nv_bfloat16* d = static_cast<__nv_bfloat16*>(e.out_pos_lp) + goff;
        float h0[16], h1[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { h0[j] = v[j] + pv[j]; h1[j] = v[16 + j] + pv[16 + j]; }
        store16<__nv_bfloat16>(d, h0, true);
        store16<__nv_bfloat16>(d + 16, h1, true);
      }
    }
  }
  if (warp < 2) {  // producer / MMA warps take part in the LayerNorm barrier
    cluster_arrive();
    cluster_wait();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(ctl->tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent weight-resident kernel for tall x (value projection). K == 256, N % 128 == 0.
// CTA c owns output-column tile (c % n_tiles) and walks the row tiles group, group + n_groups, ...
// ------------------------------------------------------------------------------------------------
// BN = output columns per CTA: 256 when N allows it (every x-tile is then read by N/256 instead of N/128 CTAs,
// which halves the dominant L2 -> SM traffic; the two 256-column accumulators fill the SM's 512 TMEM columns),
// else 128.
constexpr int kSK = 256;
constexpr int kSKB = kSK / kBK;   // 4 k-blocks
constexpr int kSlabBytes = 32 * 128;  // one epilogue slab: 32 rows x 64 bf16
// The kernel is bound by its EPILOGUE (TMEM -> registers -> bf16 -> swizzled shared memory -> TMA store), not by the
// MMAs: it runs EIGHT epilogue warps -- two per TMEM lane quadrant, each taking every other 64-column slab of a tile --
// next to the producer and the MMA warp (320 threads).
constexpr int kStreamEpiWarps = 8;
constexpr int kStreamThreads = 64 + kStreamEpiWarps * 32;
template <int BN>
struct StreamCfg {
  static constexpr int kStages = BN == 256 ? 3 : 6;  // x-tile ring of 16 KiB stages (shared memory budget)
  static constexpr int kSlabs = BN == 256 ? 1 : 2;   // output slabs per epilogue warp (TMA stores in flight)
};

template <int BN>
struct StreamCtl {
  uint64_t full[StreamCfg<BN>::kStages];
  uint64_t empty[StreamCfg<BN>::kStages];
  uint64_t b_full;
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  float bias[BN];
};

template <int BN>
__global__ void __launch_bounds__(kStreamThreads, 1)
gemm_stream_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                   const __grid_constant__ CUtensorMap tmap_y, const float* __restrict__ bias,
                   const uint8_t* __restrict__ zero_rows, int64_t M, int n_tiles, int n_groups) {
  constexpr int kSBN = BN;
  constexpr int kSStages = StreamCfg<BN>::kStages;
  constexpr uint32_t kABytes = kBM * kBK * 2;   // 16 KiB
  constexpr uint32_t kBBytes = kSBN * kBK * 2;  // 16 / 32 KiB per k-block
  constexpr uint32_t kTmemCols = 2 * kSBN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = base;                                  // 4 k-blocks of the weight slab, resident
  uint8_t* smem_a = smem_b + kSKB * kBBytes;               // ring of 16 KiB x-tile stages
  uint8_t* smem_o = smem_a + kSStages * kABytes;           // 8 warps x 2 slabs x 4 KiB
  constexpr int kSlabs = StreamCfg<BN>::kSlabs;
  StreamCtl<BN>* ctl = reinterpret_cast<StreamCtl<BN>*>(smem_o + kStreamEpiWarps * kSlabs * kSlabBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x % n_tiles;
  const int group = blockIdx.x / n_tiles;
  const int n0 = n_tile * kSBN;
  const int m_tiles = static_cast<int>((M + kBM - 1) / kBM);
  const int my_tiles = group < m_tiles ? (m_tiles - group + n_groups - 1) / n_groups : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_y)) : "memory");
    for (int s = 0; s < kSStages; ++s) {
      mbar_init(smem_u32(&ctl->full[s]), 1);
      mbar_init(smem_u32(&ctl->empty[s]), 1);
    }
    mbar_init(smem_u32(&ctl->b_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&ctl->acc_full[s]), 1);
      mbar_init(smem_u32(&ctl->acc_empty[s]), kStreamEpiWarps);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bf = smem_u32(&ctl->b_full);
    mbar_expect_tx(bf, kSKB * kBBytes);
    for (int kb = 0; kb < kSKB; ++kb) tma_load_2d(smem_u32(smem_b + kb * kBBytes), &tmap_w, bf, kb * kBK, n0);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < kSBN; i += kStreamThreads - 64) ctl->bias[i] = bias ? __ldg(bias + n0 + i) : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      pdl_wait();
      int it = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int m0 = (group + t * n_groups) * kBM;
        for (int kb = 0; kb < kSKB; ++kb, ++it) {
          const int s = it % kSStages;
          mbar_wait(smem_u32(&ctl->empty[s]), ((it / kSStages) & 1) ^ 1);
          const uint32_t full = smem_u32(&ctl->full[s]);
          mbar_expect_tx(full, kABytes);
          tma_load_2d(smem_u32(smem_a + s * kABytes), &tmap_x, full, kb * kBK, m0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kBM, kSBN);
      mbar_wait(smem_u32(&ctl->b_full), 0);
      int it = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int as = t & 1;
        mbar_wait(smem_u32(&ctl->acc_empty[as]), ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * kSBN;
        for (int kb = 0; kb < kSKB; ++kb, ++it) {
          const int s = it % kSStages;
          mbar_wait(smem_u32(&ctl->full[s]), (it / kSStages) & 1);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(smem_u32(smem_a + s * kABytes));
          const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + kb * kBBytes));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) umma_bf16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(smem_u32(&ctl->empty[s]));
        }
        umma_commit(smem_u32(&ctl->acc_full[as]));
      }
    }
    __syncwarp();
  } else {
    // epilogue warp (quad, part): rows [quad*32, quad*32+32) of every tile (its TMEM lane quadrant = warp % 4), the
    // 64-column slabs part, part + 2, ...; two private 4 KiB staging slabs
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;   // 0 or 1
    uint8_t* slab = smem_o + (warp - 2) * kSlabs * kSlabBytes;
    pdl_wait();
    int n_store = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int as = t & 1;
      const int m0 = (group + t * n_groups) * kBM;
      const int64_t row = static_cast<int64_t>(m0) + quad * 32 + lane;
      const bool zero = zero_rows != nullptr && row < M && zero_rows[row] != 0;
      mbar_wait(smem_u32(&ctl->acc_full[as]), (t >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + as * kSBN + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
      for (int half = part; half < kSBN / 64; half += kStreamEpiWarps / 4) {
        uint8_t* sl = slab + (n_store % kSlabs) * kSlabBytes;
        // the TMA store issued kSlabs slabs ago must have finished READING this slab
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kSlabs - 1) : "memory");
        __syncwarp();
        // all 64 columns of the slab are requested from TMEM before the single wait (one TMEM round trip per
        // slab instead of four: the epilogue, not the MMAs, bounds this kernel)
        uint32_t r[4][16];
#pragma unroll
        for (int q = 0; q < 4; ++q) tmem_ld16(tacc + half * 64 + q * 16, r[q]);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = q * 16;
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = zero ? 0.0f : __uint_as_float(r[q][j]) + ctl->bias[half * 64 + c + j];
          uint4 a, b;
          a.x = float2_to_bf16x2(v[0], v[1]);   a.y = float2_to_bf16x2(v[2], v[3]);
          a.z = float2_to_bf16x2(v[4], v[5]);   a.w = float2_to_bf16x2(v[6], v[7]);
          b.x = float2_to_bf16x2(v[8], v[9]);   b.y = float2_to_bf16x2(v[10], v[11]);
          b.z = float2_to_bf16x2(v[12], v[13]); b.w = float2_to_bf16x2(v[14], v[15]);
          // SWIZZLE_128B: 16-byte chunk j of row r lives at chunk (j ^ (r & 7))
          const int j0 = c / 8;
          *reinterpret_cast<uint4*>(sl + lane * 128 + (((j0) ^ (lane & 7)) << 4)) = a;
          *reinterpret_cast<uint4*>(sl + lane * 128 + (((j0 + 1) ^ (lane & 7)) << 4)) = b;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmap_y, smem_u32(sl), n0 + half * 64, m0 + quad * 32);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++n_store;
      }
      // all TMEM reads of this accumulator stage are complete (wait::ld above): hand it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ctl->acc_empty[as]));
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Tall Linear(256 -> 256) + LayerNorm with a fused class-score head, one CTA per 128 rows (BN = 256: the
// whole LayerNorm row sits in one thread's TMEM lane, no cluster exchange). The encoder-side
// `enc_output` + `enc_score_head` of the query selection (ultralytics/nn/modules/head.py:1039-1041):
//   f = LayerNorm((valid ? x : 0) . w^T + b) * gamma + beta;  logits = f . ws^T + bs;  max_logit = max_c logits
// ------------------------------------------------------------------------------------------------
constexpr int kRowLnMaxNc = 8;

struct RowLnCtl {
  uint64_t full[4];
  uint64_t tmem_full;
  uint32_t tmem_base;
  float bias[kLnCols], gamma[kLnCols], beta[kLnCols];
  float ws[kRowLnMaxNc][kLnCols];
  float bs[kRowLnMaxNc];
};


__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_rowln_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  const RowLnArgs a) {
  constexpr uint32_t kABytes = kBM * kBK * 2;       // 16 KiB
  constexpr uint32_t kBBytes = kLnCols * kBK * 2;   // 32 KiB
  constexpr int kKB = 4;                            // K = 256
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = base;
  uint8_t* smem_b = base + kKB * kABytes;
  RowLnCtl* ctl = reinterpret_cast<RowLnCtl*>(smem_b + kKB * kBBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBM;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
    for (int s = 0; s < kKB; ++s) mbar_init(smem_u32(&ctl->full[s]), 2);
    mbar_init(smem_u32(&ctl->tmem_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int kb = 0; kb < kKB; ++kb) {  // weights: immutable, requested before the dependency wait
      const uint32_t full = smem_u32(&ctl->full[kb]);
      mbar_expect_tx(full, kBBytes);
      tma_load_2d(smem_u32(smem_b + kb * kBBytes), &tmap_w, full, kb * kBK, 0);
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)),
                 "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < kLnCols; i += kGemmThreads - 64) {
      ctl->bias[i] = a.bias ? __ldg(a.bias + i) : 0.0f;
      ctl->gamma[i] = __ldg(a.gamma + i);
      ctl->beta[i] = __ldg(a.beta + i);
      for (int c = 0; c < a.nc; ++c) ctl->ws[c][i] = __ldg(a.score_w + c * kLnCols + i);
    }
    if (threadIdx.x - 64 < a.nc) ctl->bs[threadIdx.x - 64] = a.score_b ? __ldg(a.score_b + threadIdx.x - 64) : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  const uint32_t tmem_acc = ctl->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      pdl_wait();
      for (int kb = 0; kb < kKB; ++kb) {
        const uint32_t full = smem_u32(&ctl->full[kb]);
        mbar_expect_tx(full, kABytes);
        tma_load_2d(smem_u32(smem_a + kb * kABytes), &tmap_x, full, kb * kBK, m0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kBM, kLnCols);
      for (int kb = 0; kb < kKB; ++kb) {
        mbar_wait(smem_u32(&ctl->full[kb]), 0);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc(smem_u32(smem_a + kb * kABytes));
        const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + kb * kBBytes));
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) umma_bf16(tmem_acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
      }
      umma_commit(smem_u32(&ctl->tmem_full));
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int64_t row = static_cast<int64_t>(m0) + quad * 32 + lane;
    const bool row_ok = row < a.M;
    const uint32_t tbase = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16);
    pdl_wait();
    const float keep = (row_ok && a.zero_acc_rows != nullptr && a.zero_acc_rows[row] != 0) ? 0.0f : 1.0f;
    mbar_wait(smem_u32(&ctl->tmem_full), 0);
    tc_fence_after();
    // pass 1: mean
    float sum = 0.0f;
#pragma unroll 1
    for (int c0 = 0; c0 < kLnCols; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tbase + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) sum += fmaf(keep, __uint_as_float(r[j]), ctl->bias[c0 + j]);
    }
    const float mean = sum * (1.0f / kLnCols);
    // pass 2: centred variance
    float sq = 0.0f;
#pragma unroll 1
    for (int c0 = 0; c0 < kLnCols; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tbase + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float d = fmaf(keep, __uint_as_float(r[j]), ctl->bias[c0 + j]) - mean;
        sq = fmaf(d, d, sq);
      }
    }
    const float rstd = rsqrtf(sq * (1.0f / kLnCols) + a.eps);
    // pass 3: normalise, store, class scores
    float dot[kRowLnMaxNc];
#pragma unroll
    for (int c = 0; c < kRowLnMaxNc; ++c) dot[c] = 0.0f;
#pragma unroll 1
    for (int c0 = 0; c0 < kLnCols; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tbase + c0, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        v[j] = (fmaf(keep, __uint_as_float(r[j]), ctl->bias[c0 + j]) - mean) * rstd * ctl->gamma[c0 + j] + ctl->beta[c0 + j];
#pragma unroll
      for (int c = 0; c < kRowLnMaxNc; ++c) {
        if (c < a.nc) {
#pragma unroll
          for (int j = 0; j < 16; ++j) dot[c] = fmaf(v[j], ctl->ws[c][c0 + j], dot[c]);
        }
      }
      if (row_ok) {
        if (a.out_f32 != nullptr) store16<float>(a.out_f32 + row * kLnCols + c0, v, true);
        if (a.out_lp != nullptr) store16<__nv_bfloat16>(static_cast<__nv_bfloat16*>(a.out_lp) + row * kLnCols + c0, v, true);
      }
    }
    if (row_ok && a.nc > 0) {
      float best = -INFINITY;
#pragma unroll
      for (int c = 0; c < kRowLnMaxNc; ++c) {
        if (c < a.nc) {
          const float l = dot[c] + ctl->bs[c];
          if (a.logits != nullptr) a.logits[row * a.nc + c] = l;
          best = fmaxf(best, l);
        }
      }
      if (a.max_logit != nullptr) a.max_logit[row] = best;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(256) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  });
  return fn;
}

// 2-D bf16 row-major [rows, cols] tensor with row stride `ld` elements; box = [box_rows, 64 cols].
static int make_tmap(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  MOYOLO_REQUIRE(enc != nullptr, MOYOLO_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MOYOLO_REQUIRE(r == CUDA_SUCCESS, MOYOLO_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return MOYOLO_OK;
}

bool linear_tcgen05_supported(const void* x, int64_t ldx, const void* w, int64_t M, int N, int K) {
  return M > 0 && K % kBK == 0 && N % 32 == 0 && aligned16(x) && aligned16(w) && (ldx * 2) % 16 == 0 &&
         M < (1ll << 31);
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                  unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  note_launch();
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <int BN, typename TO, bool LN>
static int launch_gemm(const CUtensorMap& tx, const CUtensorMap& tx2, const CUtensorMap& tw, const GemmEpi& e,
                       cudaStream_t st) {
  constexpr size_t smem_base = kStages * (kBM * kBK * 2 + BN * kBK * 2) + sizeof(GemmCtl<BN, LN>) + 1024;
  constexpr size_t smem_max = smem_base + (LN ? kMaxScoreNc * (kLnCols / BN) * kBM * 4 : 0);
  const size_t smem = smem_base + (LN ? static_cast<size_t>(e.score_nc) * (kLnCols / BN) * kBM * 4 : 0);
  static DeviceOnce once;  // per template instantiation and device
  const int dev_ = DeviceOnce::current();
  if (!once.done(dev_)) {
    cudaError_t err = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, TO, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem_max));
    if (err != cudaSuccess)
      return fail(MOYOLO_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(err));
    once.set(dev_);
  }
  dim3 grid((e.N + BN - 1) / BN, static_cast<unsigned>((e.M + kBM - 1) / kBM));
  if constexpr (LN)
    launch_cluster(gemm_tcgen05_kernel<BN, TO, LN>, grid, dim3(kGemmThreads), smem, st, kLnCols / BN, tx, tx2, tw, e);
  else
    launch_k(gemm_tcgen05_kernel<BN, TO, LN>, grid, dim3(kGemmThreads), smem, st, tx, tx2, tw, e);
  return check_launch("gemm_tcgen05_kernel");
}

template <int BN>
static int launch_stream_bn(const void* x, int64_t ldx, const void* w, const float* bias, void* y, int64_t ldy, int64_t M,
                            int N, const uint8_t* zero_rows, int max_ctas, cudaStream_t st) {
  constexpr size_t smem = kSKB * BN * kBK * 2 + StreamCfg<BN>::kStages * kBM * kBK * 2 +
                          kStreamEpiWarps * StreamCfg<BN>::kSlabs * kSlabBytes +
                          sizeof(StreamCtl<BN>) + 1024;
  static_assert(smem <= 232448, "exceeds the 227 KiB of shared memory a CTA can opt in to");
  static DeviceOnce once;
  const int dev_ = DeviceOnce::current();
  if (!once.done(dev_)) {
    cudaError_t err = cudaFuncSetAttribute(gemm_stream_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
    if (err != cudaSuccess)
      return fail(MOYOLO_ERR_CUDA, "cudaFuncSetAttribute(stream smem=%zu): %s", smem, cudaGetErrorString(err));
    once.set(dev_);
  }
  const int n_sm = once.n_sm[dev_];
  CUtensorMap tx, tw, ty;
  int rc = make_tmap(&tx, x, M, kSK, ldx, kBM);
  if (rc != MOYOLO_OK) return rc;
  rc = make_tmap(&tw, w, N, kSK, kSK, BN);
  if (rc != MOYOLO_OK) return rc;
  rc = make_tmap(&ty, y, M, N, ldy, 32);
  if (rc != MOYOLO_OK) return rc;
  const int n_tiles = N / BN;
  const int m_tiles = static_cast<int>((M + kBM - 1) / kBM);
  int n_groups = (max_ctas > 0 && max_ctas < n_sm ? max_ctas : n_sm) / n_tiles;
  {
    // MOYOLO_STREAM_GROUPS caps the row-tile groups (CTAs = groups * N/BN): leaving some SMs free lets the short
    // query GEMMs of the frame's main chain run next to the value projection instead of queueing behind it
    static const int cap = [] { const char* e = getenv("MOYOLO_STREAM_GROUPS"); return e ? atoi(e) : 0; }();
    if (cap > 0 && n_groups > cap) n_groups = cap;
  }
  if (n_groups < 1) n_groups = 1;
  if (n_groups > m_tiles) n_groups = m_tiles;
  launch_k(gemm_stream_kernel<BN>, dim3(n_tiles * n_groups), dim3(kStreamThreads), smem, st, tx, tw, ty, bias, zero_rows, M,
           n_tiles, n_groups);
  return check_launch("gemm_stream_kernel");
}

static int launch_stream(const void* x, int64_t ldx, const void* w, const float* bias, void* y, int64_t ldy, int64_t M,
                         int N, const uint8_t* zero_rows, int max_ctas, cudaStream_t st) {
  // 128-column tiles by default (measured equal or faster than 256 on B200 for K = 256: the kernel is bound by
  // its epilogue / store path, not by re-reading x); MOYOLO_STREAM_BN=256 selects the wide variant
  static const bool wide = [] { const char* e = getenv("MOYOLO_STREAM_BN"); return e && atoi(e) == 256; }();
  if (wide && N % 256 == 0) return launch_stream_bn<256>(x, ldx, w, bias, y, ldy, M, N, zero_rows, max_ctas, st);
  return launch_stream_bn<128>(x, ldx, w, bias, y, ldy, M, N, zero_rows, max_ctas, st);
}

bool linear_tall_supported(const void* x, int64_t ldx, const void* w, const void* y, int64_t ldy, int64_t M, int N, int K) {
  return M > 0 && M < (1ll << 31) && K == kSK && N % 128 == 0 && aligned16(x) && aligned16(w) && aligned16(y) &&
         (ldx * 2) % 16 == 0 && (ldy * 2) % 16 == 0;
}

int linear_tall(const void* x, int64_t ldx, const void* w, const float* bias, void* y, int64_t ldy, int64_t M, int N,
                const uint8_t* zero_rows, int max_ctas, cudaStream_t st) {
  return launch_stream(x, ldx, w, bias, y, ldy, M, N, zero_rows, max_ctas, st);
}

// x2 / n_split: optional second A operand for output columns >= n_split (n_split % 64 == 0).
int linear_tcgen05(const void* x, int64_t ldx, const void* x2, int64_t ldx2, int n_split, const void* w,
                   const float* bias, void* y, int64_t ldy, int64_t M, int N, int K, int out_dtype, int relu,
                   const uint8_t* zero_rows, cudaStream_t st) {
  MOYOLO_REQUIRE(out_dtype == MOYOLO_F32 || out_dtype == MOYOLO_BF16, MOYOLO_ERR_UNSUPPORTED,
                 "tcgen05 engine writes fp32 or bf16");
  // tall x, K == 256: persistent weight-resident kernel (value projection)
  if (x2 == nullptr && M >= 4096 && K == kSK && N % 128 == 0 && out_dtype == MOYOLO_BF16 && !relu && aligned16(y) &&
      (ldy * 2) % 16 == 0)
    return launch_stream(x, ldx, w, bias, y, ldy, M, N, zero_rows, 0, st);
  // Tile width: narrow tiles spread the few-hundred-row query GEMMs over more SMs (latency-bound).
  int bn;
  if (M > 2048 && N % 256 == 0) bn = 256;
  else if (M > 2048 && N % 128 == 0) bn = 128;
  else if (N % 64 == 0 && static_cast<int64_t>((M + kBM - 1) / kBM) * (N / 64) >= 96) bn = 64;
  else if (N % 32 == 0) bn = 32;
  else bn = 16;
  MOYOLO_REQUIRE(bn != 16, MOYOLO_ERR_UNSUPPORTED, "tcgen05 engine needs N %% 32 == 0 (N=%d)", N);
  if (x2 != nullptr) {
    MOYOLO_REQUIRE(n_split > 0 && n_split < N && n_split % 64 == 0, MOYOLO_ERR_BAD_SHAPE,
                   "dual-operand GEMM: n_split (%d) must be a multiple of 64 inside (0, N)", n_split);
    if (bn > 64) bn = 64;
  }
  CUtensorMap tx, tx2, tw;
  int rc = make_tmap(&tx, x, M, K, ldx, kBM);
  if (rc != MOYOLO_OK) return rc;
  if (x2 != nullptr) {
    rc = make_tmap(&tx2, x2, M, K, ldx2, kBM);
    if (rc != MOYOLO_OK) return rc;
  } else {
    tx2 = tx;
  }
  rc = make_tmap(&tw, w, N, K, K, bn);
  if (rc != MOYOLO_OK) return rc;
  GemmEpi e{};
  e.bias = bias; e.y = y; e.ldy = ldy; e.M = M; e.N = N; e.K = K; e.relu = relu; e.zero_rows = zero_rows;
  e.n_split = x2 != nullptr ? n_split : N;
#define GO(BN)                                                                        \
  (out_dtype == MOYOLO_F32 ? launch_gemm<BN, float, false>(tx, tx2, tw, e, st)        \
                           : launch_gemm<BN, __nv_bfloat16, false>(tx, tx2, tw, e, st))
  switch (bn) {
    case 256: return GO(256);
    case 128: return GO(128);
    case 64: return GO(64);
    default: return GO(32);
  }
#undef GO
}

bool linear_ln_tcgen05_supported(const void* x, int64_t ldx, const void* w, int64_t M, int N, int K) {
  return N == kLnCols && linear_tcgen05_supported(x, ldx, w, M, N, K);
}

// out = LayerNorm(x . w^T + bias + residual) * gamma + beta, N == 256, all row-major contiguous [M, 256].
int linear_ln_tcgen05(const void* x, int64_t ldx, const void* w, const float* bias, const float* residual,
                      const float* gamma, const float* beta, float eps, int64_t M, int K, float* out_f32, void* out_lp,
                      const float* pos, void* out_pos_lp, const float* score_w, const float* score_b, int score_nc,
                      float* logits, float* scores, int32_t* labels, cudaStream_t st) {
  CUtensorMap tx, tw;
  int rc = make_tmap(&tx, x, M, K, ldx, kBM);
  if (rc != MOYOLO_OK) return rc;
  rc = make_tmap(&tw, w, kLnCols, K, K, 32);
  if (rc != MOYOLO_OK) return rc;
  GemmEpi e{};
  e.bias = bias; e.M = M; e.N = kLnCols; e.K = K; e.n_split = kLnCols;
  e.residual = residual; e.gamma = gamma; e.beta = beta; e.eps = eps;
  e.out_f32 = out_f32; e.out_lp = out_lp; e.pos = pos; e.out_pos_lp = out_pos_lp;
  e.score_w = score_w; e.score_b = score_b; e.score_nc = score_w != nullptr ? score_nc : 0;
  e.logits = logits; e.scores = scores; e.labels = labels;
  return launch_gemm<32, __nv_bfloat16, true>(tx, tx, tw, e, st);
}

bool ffn_ln_tcgen05_supported(const void* x, int64_t ldx, const void* w1, const void* w2, const void* h, int64_t M, int C,
                              int F) {
  return C == kLnCols && (F == 1024 || F == 256) && M > 0 && M < (1ll << 31) && aligned16(x) && aligned16(w1) &&
         aligned16(w2) && aligned16(h) && (ldx * 2) % 16 == 0;
}

template <int BN1>
static int launch_ffn_ln(const CUtensorMap& tx, const CUtensorMap& tw1, const CUtensorMap& th, const CUtensorMap& tw2,
                         const FfnArgs& e, cudaStream_t st) {
  const size_t smem = kStages * (kBM * kBK * 2 + BN1 * kBK * 2) + static_cast<size_t>(e.F / kBK) * 32 * kBK * 2 +
                      sizeof(FfnCtl<BN1>) + 1024;
  static DeviceOnce once;
  const int dev_ = DeviceOnce::current();
  if (!once.done(dev_)) {
    cudaError_t err = cudaFuncSetAttribute(gemm_ffn_ln_kernel<BN1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
    if (err != cudaSuccess)
      return fail(MOYOLO_ERR_CUDA, "cudaFuncSetAttribute(ffn smem=%zu): %s", smem, cudaGetErrorString(err));
    once.set(dev_);
  }
  dim3 grid(kLnCols / 32, static_cast<unsigned>((e.M + kBM - 1) / kBM));
  launch_cluster(gemm_ffn_ln_kernel<BN1>, grid, dim3(kGemmThreads), smem, st, kLnCols / 32, tx, tw1, th, tw2, e);
  return check_launch("gemm_ffn_ln_kernel");
}

// out = LayerNorm(residual + relu(x . w1^T + b1) . w2^T + b2) * gamma + beta; x [M,256], w1 [F,256], w2 [256,F],
// h = bf16 [M,F] scratch.
int ffn_ln_tcgen05(const void* x, int64_t ldx, const void* w1, const float* b1, const void* w2, const float* b2, void* h,
                   int F, const float* residual, const float* gamma, const float* beta, float eps, int64_t M,
                   float* out_f32, void* out_lp, const float* pos, void* out_pos_lp, cudaStream_t st) {
  CUtensorMap tx, tw1, th, tw2;
  const int bn1 = F / (kLnCols / 32);
  int rc = make_tmap(&tx, x, M, kLnCols, ldx, kBM);
  if (rc != MOYOLO_OK) return rc;
  rc = make_tmap(&tw1, w1, F, kLnCols, kLnCols, bn1);
  if (rc != MOYOLO_OK) return rc;
  rc = make_tmap(&th, h, M, F, F, kBM);
  if (rc != MOYOLO_OK) return rc;
  rc = make_tmap(&tw2, w2, kLnCols, F, F, 32);
  if (rc != MOYOLO_OK) return rc;
  FfnArgs e{};
  e.b1 = b1; e.b2 = b2; e.h = h; e.F = F; e.M = M; e.residual = residual; e.gamma = gamma; e.beta = beta; e.eps = eps;
  e.out_f32 = out_f32; e.out_lp = out_lp; e.pos = pos; e.out_pos_lp = out_pos_lp;
  return bn1 == 128 ? launch_ffn_ln<128>(tx, tw1, th, tw2, e, st) : launch_ffn_ln<32>(tx, tw1, th, tw2, e, st);
}

// Tall Linear(256->256) + LayerNorm (+ class scores): see gemm_rowln_kernel.
int linear_rowln_tcgen05(const void* x, int64_t ldx, const void* w, const RowLnArgs& a, cudaStream_t st) {
  constexpr size_t smem = 4 * (kBM * kBK * 2 + kLnCols * kBK * 2) + sizeof(RowLnCtl) + 1024;
  static DeviceOnce once;
  const int dev_ = DeviceOnce::current();
  if (!once.done(dev_)) {
    cudaError_t err = cudaFuncSetAttribute(gemm_rowln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
    if (err != cudaSuccess)
      return fail(MOYOLO_ERR_CUDA, "cudaFuncSetAttribute(rowln smem=%zu): %s", smem, cudaGetErrorString(err));
    once.set(dev_);
  }
  CUtensorMap tx, tw;
  int rc = make_tmap(&tx, x, a.M, kLnCols, ldx, kBM);
  if (rc != MOYOLO_OK) return rc;
  rc = make_tmap(&tw, w, kLnCols, kLnCols, kLnCols, 256);
  if (rc != MOYOLO_OK) return rc;
  launch_k(gemm_rowln_kernel, dim3(static_cast<unsigned>((a.M + kBM - 1) / kBM)), dim3(kGemmThreads), smem, st, tx, tw, a);
  return check_launch("gemm_rowln_kernel");
}

}  // namespace moyolo
