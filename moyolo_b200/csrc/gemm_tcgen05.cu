// tcgen05 / TMA / TMEM GEMM for sm_100a:  y[M,N] = act(x[M,K] . w[N,K]^T + bias), bf16 operands,
// fp32 accumulation in tensor memory. Serves every dense contraction of the decoder hot path
// (value_proj for all layers at once, sampling_offsets|attention_weights, output_proj, MHA in/out
// projections, FFN, bbox-MLP hidden layers; nn.Linear call sites of
// ultralytics/nn/modules/transformer.py:264,268,269,286,576-580,638 and MOTR/models/qim.py:276-290).
//
// Both operands are K-major (x rows and nn.Linear's [out,in] weight rows are contiguous in K), so no
// transposes are needed: TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) stages [128 x 64] x-tiles and
// [BN x 64] w-tiles into a 4-deep shared-memory ring; one elected thread issues
// tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) four times per stage; accumulators live in
// BN TMEM columns; four epilogue warps read them back with tcgen05.ld (one output row per thread),
// add bias / ReLU / row mask and store fp32 or bf16.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2..5
// epilogue (TMEM lane quadrant = warp_id % 4). One output tile per CTA; BN is chosen small (32) when
// M is small so that the few hundred query rows still spread over many SMs.
#include <cuda.h>

#include <mutex>

#include "common.cuh"

namespace moyolo {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int kStages = 4;
constexpr int kGemmThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
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

struct GemmSmem {
  // data ring first: every stage base must be 1024-byte aligned for SWIZZLE_128B
  // (sizes are multiples of 1024: A 16 KiB, B BN*128 B with BN % 8 == 0 ... BN >= 32 -> 4 KiB)
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
};

template <int BN, typename TO>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const float* __restrict__ bias, TO* __restrict__ y, int64_t ldy, int64_t M, int N, int K,
                    int relu, const uint8_t* __restrict__ zero_rows) {
  constexpr uint32_t kABytes = kBM * kBK * 2;
  constexpr uint32_t kBBytes = BN * kBK * 2;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;  // power of two >= 32 (BN in {32,64,128,256})
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte align the dynamic shared memory window
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = base;
  uint8_t* smem_b = base + kStages * kABytes;
  GemmSmem* ctl = reinterpret_cast<GemmSmem*>(smem_b + kStages * kBBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kBM;
  const int n0 = blockIdx.x * BN;
  const int num_kb = K / kBK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w)) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&ctl->full[s]), 1);
      mbar_init(smem_u32(&ctl->empty[s]), 1);
    }
    mbar_init(smem_u32(&ctl->tmem_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = ctl->tmem_base;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(smem_u32(&ctl->empty[s]), ph ^ 1);
        const uint32_t full = smem_u32(&ctl->full[s]);
        mbar_expect_tx(full, kABytes + kBBytes);
        tma_load_2d(smem_u32(smem_a + s * kABytes), &tmap_x, full, kb * kBK, m0);
        tma_load_2d(smem_u32(smem_b + s * kBBytes), &tmap_w, full, kb * kBK, n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kBM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(smem_u32(&ctl->full[s]), ph);
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
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4, one output row per thread =====
    const int quad = warp & 3;
    const int64_t row = static_cast<int64_t>(m0) + quad * 32 + lane;
    mbar_wait(smem_u32(&ctl->tmem_full), 0);
    tc_fence_after();
    const bool row_ok = row < M;
    const bool zero = row_ok && zero_rows != nullptr && zero_rows[row] != 0;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(y) & 15u) == 0 && (ldy * sizeof(TO)) % 16 == 0;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + c0, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (!row_ok || n0 + c0 >= N) continue;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float t = __uint_as_float(r[j]) + (bias ? __ldg(bias + n0 + c0 + j) : 0.0f);
        if (relu) t = fmaxf(t, 0.0f);
        v[j] = zero ? 0.0f : t;
      }
      TO* dst = y + row * ldy + n0 + c0;
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(kTmemCols) : "memory");
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

template <int BN, typename TO>
static int launch_gemm(const CUtensorMap& tx, const CUtensorMap& tw, const float* bias, void* y, int64_t ldy,
                       int64_t M, int N, int K, int relu, const uint8_t* zero_rows, cudaStream_t st) {
  constexpr size_t smem = kStages * (kBM * kBK * 2 + BN * kBK * 2) + sizeof(GemmSmem) + 1024;
  static bool configured = false;  // per template instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return fail(MOYOLO_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
    configured = true;
  }
  dim3 grid((N + BN - 1) / BN, static_cast<unsigned>((M + kBM - 1) / kBM));
  gemm_tcgen05_kernel<BN, TO><<<grid, kGemmThreads, smem, st>>>(tx, tw, bias, static_cast<TO*>(y), ldy, M, N, K, relu,
                                                               zero_rows);
  return check_launch("gemm_tcgen05_kernel");
}

int linear_tcgen05(const void* x, int64_t ldx, const void* w, const float* bias, void* y, int64_t ldy, int64_t M, int N,
                   int K, int out_dtype, int relu, const uint8_t* zero_rows, cudaStream_t st) {
  // Tile width: wide tiles amortise the x-tile for the big value_proj GEMM; narrow tiles spread the
  // few-hundred-row query GEMMs over more SMs (they are latency- not throughput-bound).
  int bn;
  if (M > 2048 && N % 256 == 0) bn = 256;
  else if (M > 2048 && N % 128 == 0) bn = 128;
  else if (N % 64 == 0 && static_cast<int64_t>((M + kBM - 1) / kBM) * (N / 64) >= 96) bn = 64;
  else if (N % 32 == 0) bn = 32;
  else bn = 16;
  MOYOLO_REQUIRE(bn != 16, MOYOLO_ERR_UNSUPPORTED, "tcgen05 engine needs N %% 32 == 0 (N=%d)", N);
  CUtensorMap tx, tw;
  int rc = make_tmap(&tx, x, M, K, ldx, kBM);
  if (rc != MOYOLO_OK) return rc;
  rc = make_tmap(&tw, w, N, K, K, bn);
  if (rc != MOYOLO_OK) return rc;
#define GO(BN)                                                                                              \
  (out_dtype == MOYOLO_F32 ? launch_gemm<BN, float>(tx, tw, bias, y, ldy, M, N, K, relu, zero_rows, st)     \
                           : launch_gemm<BN, __nv_bfloat16>(tx, tw, bias, y, ldy, M, N, K, relu, zero_rows, st))
  MOYOLO_REQUIRE(out_dtype == MOYOLO_F32 || out_dtype == MOYOLO_BF16, MOYOLO_ERR_UNSUPPORTED,
                 "tcgen05 engine writes fp32 or bf16");
  switch (bn) {
    case 256: return GO(256);
    case 128: return GO(128);
    case 64: return GO(64);
    default: return GO(32);
  }
#undef GO
}

}  // namespace moyolo
