// Shared helpers of libmoyolo_b200: status/error plumbing, vector load/store, warp reductions.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/moyolo_b200.h"

namespace moyolo {

// thread-local last-error text (moyolo_last_error)
char* last_error_buf();
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);

#define MOYOLO_REQUIRE(cond, code, ...)            \
  do {                                             \
    if (!(cond)) return ::moyolo::fail(code, __VA_ARGS__); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Once-per-DEVICE flag: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are properties of a
// device, not of the process, so launchers key their one-time setup by the current device (bit d of the mask).
// Benign if two host threads race: the attribute is idempotent.
struct DeviceOnce {
  std::atomic<unsigned long long> mask{0};
  int n_sm[64] = {};
  static int current() {
    int d = 0;
    cudaGetDevice(&d);
    return d < 0 || d > 63 ? 63 : d;
  }
  bool done(int dev) const { return (mask.load(std::memory_order_acquire) >> dev) & 1ull; }
  void set(int dev) {
    cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev);
    mask.fetch_or(1ull << dev, std::memory_order_release);
  }
};

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization so
// that, inside a stream or a captured graph, the next kernel's CTAs are scheduled while the previous
// kernel drains (the frame is a chain of ~100 short dependent launches). Contract for kernels:
//   * pdl_trigger() once every resource the CTA needs has been acquired (TMEM columns in particular:
//     a dependent that grabbed TMEM first could otherwise starve the kernel it waits for);
//   * pdl_wait() before the first access to any global memory that is not immutable during a frame
//     (weights, biases and LayerNorm parameters are immutable; activations, masks and state are not).
// MOYOLO_PDL=0 in the environment disables the attribute (the device instructions are then no-ops).
bool pdl_enabled();
void note_launch();  // per-thread count of kernels launched through launch_k (moyolo_launch_count)

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  note_launch();
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct LevelTable {
  int n;
  int h[MOYOLO_MAX_LEVELS];
  int w[MOYOLO_MAX_LEVELS];
  int start[MOYOLO_MAX_LEVELS];
};

// Builds the level table and validates sum(H*W) == len_v (transformer.py:262 assert).
int make_levels(const int32_t* shapes_hw_host, int n_levels, int64_t len_v, LevelTable* out);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 128-bit read-only global load.
__device__ __forceinline__ uint4 ldg128(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

__device__ __forceinline__ float2 bf16x2_to_float2(uint32_t u) {
  // bf16 -> fp32 is a 16-bit left shift; low half is element 0.
  float2 r;
  r.x = __uint_as_float(u << 16);
  r.y = __uint_as_float(u & 0xffff0000u);
  return r;
}
__device__ __forceinline__ uint32_t float2_to_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <typename T>
__device__ __forceinline__ float to_float(T v);
template <>
__device__ __forceinline__ float to_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_float(float v);
template <>
__device__ __forceinline__ float from_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- cp.async / ldmatrix / mma.sync helpers (attention.cu, msda.cu) ----
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 -> 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Arguments of the tall Linear(256->256) + LayerNorm + class-score kernel (gemm_tcgen05.cu, selector.cu).
struct RowLnArgs {
  const float* bias;
  const float* gamma;
  const float* beta;
  float eps;
  const uint8_t* zero_acc_rows;  // rows whose INPUT is treated as zero (valid_mask * feats)
  const float* score_w;          // [nc, 256] or NULL
  const float* score_b;
  int nc;
  int64_t M;
  float* out_f32;                // [M, 256] or NULL
  void* out_lp;                  // [M, 256] bf16 or NULL
  float* logits;                 // [M, nc] or NULL
  float* max_logit;              // [M] or NULL
};

// batch index of a row for dense (row_offsets == nullptr) or ragged batches.
__device__ __forceinline__ int batch_of_row(int64_t row, const int32_t* __restrict__ row_offsets,
                                            int batch, int64_t rows_per_batch) {
  if (row_offsets == nullptr) return static_cast<int>(row / rows_per_batch);
  int b = 0;
  while (b + 1 < batch && row >= row_offsets[b + 1]) ++b;
  return b;
}

}  // namespace moyolo
