// Track-query update on device: the ID assigner RuntimeTrackerBase.update
// (ultralytics/nn/modules/head.py:1201-1283) restated as block-wide scans, and the active-track
// selection (MOTR/models/qim.py:184-187 + MOTR/models/structures/instances.py:152-178) as a
// stream compaction. The reference runs these as Python loops with one host sync per element.
//
// The sequential counter `max_obj_id++` in query order == exclusive prefix count of "new" rows.
// The greedy O(n^2) duplicate filter (head.py:1155-1171) == pairwise bit matrix + one-warp sweep.
// IoU arithmetic uses explicit round-to-nearest intrinsics so no FMA contraction can change a
// `> 0.8` decision relative to the reference's separate fp32 tensor ops (head.py:1173-1196).
#include "common.cuh"

namespace moyolo {

constexpr int kTrkThreads = 1024;
constexpr int kTrkMaxN = 4096;
constexpr int kSuppSmemWords = 10240;  // 40 KiB: the bit matrix of up to ~570 active tracks

// Block-wide exclusive scan of one int per thread; returns exclusive prefix, total via *total.
__device__ int block_exclusive_scan(int v, int* total, int* s_warp /* [33] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protect s_warp reuse across calls
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < (blockDim.x >> 5)) ? s_warp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    s_warp[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  *total = s_warp[32];
  return s_warp[warp] + inc - v;
}

__device__ __forceinline__ bool iou_gt(const float* __restrict__ b1, const float* __restrict__ b2, float thr) {
  // boxes are read as (x, y, w, h) exactly as head.py:1173-1196 does
  const float x1 = b1[0], y1 = b1[1], w1 = b1[2], h1 = b1[3];
  const float x2 = b2[0], y2 = b2[1], w2 = b2[2], h2 = b2[3];
  if (fabsf(__fsub_rn(x1, x2)) > __fmul_rn(0.5f, fminf(x1, x2))) return false;
  if (fabsf(__fsub_rn(y1, y2)) > __fmul_rn(0.5f, fminf(y1, y2))) return false;
  const float ix1 = fmaxf(x1, x2), iy1 = fmaxf(y1, y2);
  const float ix2 = fminf(__fadd_rn(x1, w1), __fadd_rn(x2, w2));
  const float iy2 = fminf(__fadd_rn(y1, h1), __fadd_rn(y2, h2));
  const float dx = __fsub_rn(ix2, ix1), dy = __fsub_rn(iy2, iy1);
  const float inter = __fmul_rn(dx > 0.0f ? dx : 0.0f, dy > 0.0f ? dy : 0.0f);
  const float a1 = __fmul_rn(w1, h1), a2 = __fmul_rn(w2, h2);
  const float uni = __fsub_rn(__fadd_rn(a1, a2), inter);
  return __fdiv_rn(inter, uni) > thr;  // NaN (0/0) compares false, as in the reference
}

struct TrackWs {
  int32_t* active_idx;  // [n]
  int32_t* keep;        // [n]
  uint32_t* supp;       // [n * words]
};
__host__ __device__ inline int64_t trk_words(int64_t n) { return (n + 31) / 32; }
__host__ __device__ inline TrackWs carve_ws(void* ws, int64_t n) {
  TrackWs t;
  t.active_idx = static_cast<int32_t*>(ws);
  t.keep = t.active_idx + n;
  t.supp = reinterpret_cast<uint32_t*>(t.keep + n);
  return t;
}

__global__ void __launch_bounds__(kTrkThreads) track_assign_kernel(
    const float* __restrict__ scores, const float* __restrict__ boxes, int64_t* __restrict__ obj_idxes,
    int64_t* __restrict__ disappear_time, int64_t* __restrict__ counters, int n, float score_thresh,
    float filter_thresh, int miss_tolerance, float iou_thresh, void* workspace,
    const int32_t* __restrict__ row_offsets, int64_t ws_stride, const int32_t* __restrict__ ctrl, int do_assign) {
  pdl_trigger();
  pdl_wait();
  __shared__ int s_warp[33];
  __shared__ uint32_t s_zmask[kTrkMaxN / 32];    // rows with a non-empty suppression set
  __shared__ uint32_t s_removed[kTrkMaxN / 32];  // rows dropped by the greedy filter
  __shared__ uint32_t s_supp[kSuppSmemWords];    // suppression bit matrix when it fits (else workspace)
  if (ctrl != nullptr && ctrl[0] != 0) return;  // aborted speculative frame (see frame.cu): ID counters untouched
  if (row_offsets != nullptr) {  // batched: one CTA per sequence, rows [row_offsets[s], row_offsets[s+1])
    const int seq = blockIdx.x;
    const int off = row_offsets[seq];
    n = row_offsets[seq + 1] - off;
    scores += off;
    boxes += static_cast<int64_t>(off) * 4;
    obj_idxes += off;
    disappear_time += off;
    counters += 2 * seq;
    workspace = static_cast<char*>(workspace) + static_cast<int64_t>(seq) * ws_stride;
    if (n <= 0) return;
  }
  const TrackWs ws = carve_ws(workspace, n);
  const int per = (n + kTrkThreads - 1) / kTrkThreads;
  const int begin = min(static_cast<int>(threadIdx.x) * per, n);
  const int end = min(begin + per, n);
  const int64_t max_obj_id = counters[0];
  const int64_t max_obj_id_pre = counters[1];

  // ---- A. ID assignment in query order (head.py:1232-1243) ----
  // (do_assign == 0: the ids were already updated by frame_assign_compact; only the counters' side effects follow)
  int n_new_local = 0;
  if (do_assign)
    for (int i = begin; i < end; ++i)
      n_new_local += (obj_idxes[i] == -1 && scores[i] >= score_thresh) ? 1 : 0;
  int n_new_total = 0;
  int new_rank = do_assign ? block_exclusive_scan(n_new_local, &n_new_total, s_warp) : 0;
  int n_act_local = 0;
  for (int i = begin; i < end; ++i) {
    int64_t id = obj_idxes[i];
    if (do_assign) {
      const float s = scores[i];
      if (id == -1 && s >= score_thresh) {
        id = max_obj_id + new_rank++;
      } else if (id >= 0 && s < filter_thresh) {
        const int64_t dt = disappear_time[i] + 1;
        disappear_time[i] = dt;
        if (dt >= miss_tolerance) id = -1;
      }
      obj_idxes[i] = id;
    }
    n_act_local += id >= 0 ? 1 : 0;
  }
  // ---- B. active subset, in order (head.py:1245-1250) ----
  int n_active;
  int act_rank = block_exclusive_scan(n_act_local, &n_active, s_warp);
  if (n_active == 0) {
    // early return of the reference: counters keep the (unchanged) value; nothing was assigned
    return;
  }
  for (int i = begin; i < end; ++i)
    if (obj_idxes[i] >= 0) ws.active_idx[act_rank++] = i;
  __syncthreads();

  // ---- C. pairwise suppression bits, then the greedy sweep (head.py:1155-1171) ----
  // supp[i] has bit j set iff j > i and IoU(i, j) > thr. Row i only matters to the sweep when it is
  // non-empty (s_zmask), and because its bits all lie above i, "removed by a kept earlier row" is
  // simply the OR of the kept non-empty rows: the sequential part visits the (few) non-empty rows only.
  const int words = static_cast<int>(trk_words(n_active));
  uint32_t* supp = (n_active * words <= kSuppSmemWords) ? s_supp : ws.supp;
  for (int t = threadIdx.x; t < kTrkMaxN / 32; t += kTrkThreads) { s_zmask[t] = 0u; s_removed[t] = 0u; }
  __syncthreads();
  // one WARP per 32-bit word of the matrix, one pair test per lane (ballot -> word): the 32 IoU tests of a word run
  // side by side instead of one after the other in a single thread (16.7 -> ~4 us at 50 active tracks, on the tail's
  // critical path); every pair is tested by exactly the same code, so the bits are unchanged
  for (int t = threadIdx.x >> 5; t < n_active * words; t += kTrkThreads >> 5) {
    const int i = t / words, wj = t % words;
    const int j = wj * 32 + (threadIdx.x & 31);
    bool bit = false;
    if (j > i && j < n_active)
      bit = iou_gt(boxes + static_cast<int64_t>(ws.active_idx[i]) * 4, boxes + static_cast<int64_t>(ws.active_idx[j]) * 4,
                   iou_thresh);
    const uint32_t bits = __ballot_sync(0xffffffffu, bit);
    if ((threadIdx.x & 31) == 0) {
      supp[t] = bits;
      if (bits != 0u) atomicOr(&s_zmask[i >> 5], 1u << (i & 31));
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    constexpr int kSlots = kTrkMaxN / 32 / 32;  // words per lane
    uint32_t removed[kSlots];
#pragma unroll
    for (int s = 0; s < kSlots; ++s) removed[s] = 0u;
    for (int wz = 0; wz < words; ++wz) {
      uint32_t zm = s_zmask[wz];  // warp-uniform
      while (zm != 0u) {
        const int i = wz * 32 + (__ffs(zm) - 1);
        zm &= zm - 1u;
        const int wi = i >> 5, owner = wi & 31, slot = wi >> 5;
        uint32_t wv = 0u;
#pragma unroll
        for (int s = 0; s < kSlots; ++s) wv = (s == slot) ? removed[s] : wv;
        wv = __shfl_sync(0xffffffffu, wv, owner);
        const bool kept = ((wv >> (i & 31)) & 1u) == 0u;
        if (kept) {
#pragma unroll
          for (int s = 0; s < kSlots; ++s) {
            const int w = lane + 32 * s;
            if (w < words) removed[s] |= supp[i * words + w];
          }
        }
      }
    }
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      const int w = lane + 32 * s;
      if (w < words) s_removed[w] = removed[s];
    }
  }
  __syncthreads();
  for (int a = threadIdx.x; a < n_active; a += kTrkThreads) ws.keep[a] = ((s_removed[a >> 5] >> (a & 31)) & 1u) ? 0 : 1;

  // ---- D. renumbering side effect on the counters (head.py:1268-1282) ----
  // kept rows with id > max_obj_id_pre become max_obj_id_pre+1, +2, ... in order; the new
  // max_obj_id is max(renumbered ids)+1. Only the counters survive (the filtered copy is dropped).
  const int aper = (n_active + kTrkThreads - 1) / kTrkThreads;
  const int ab = min(static_cast<int>(threadIdx.x) * aper, n_active);
  const int ae = min(ab + aper, n_active);
  int n_renum_local = 0;
  long long max_old_local = -1;
  for (int a = ab; a < ae; ++a) {
    if ((s_removed[a >> 5] >> (a & 31)) & 1u) continue;
    const int64_t id = obj_idxes[ws.active_idx[a]];
    if (id > max_obj_id_pre) ++n_renum_local;
    else max_old_local = max(max_old_local, static_cast<long long>(id));
  }
  int n_renum;
  block_exclusive_scan(n_renum_local, &n_renum, s_warp);
  // block max of max_old_local (ids fit in int for the reduction: they are < 2^31 in practice)
  int mo = static_cast<int>(max_old_local);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mo = max(mo, __shfl_xor_sync(0xffffffffu, mo, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = mo;
  __syncthreads();
  if (threadIdx.x == 0) {
    int best = -1;
    for (int w = 0; w < (kTrkThreads >> 5); ++w) best = max(best, s_warp[w]);
    long long mx = best;
    if (n_renum > 0) mx = max(mx, static_cast<long long>(max_obj_id_pre) + n_renum);
    counters[0] = mx + 1;
    counters[1] = mx;
  }
  (void)n_new_total;
  (void)max_obj_id;
}

__global__ void __launch_bounds__(kTrkThreads) track_select_kernel(const int64_t* __restrict__ obj_idxes,
                                                                   int n, int32_t* __restrict__ n_active,
                                                                   int32_t* __restrict__ active_index) {
  pdl_trigger();
  pdl_wait();
  __shared__ int s_warp[33];
  const int per = (n + kTrkThreads - 1) / kTrkThreads;
  const int begin = min(static_cast<int>(threadIdx.x) * per, n);
  const int end = min(begin + per, n);
  int local = 0;
  for (int i = begin; i < end; ++i) local += obj_idxes[i] >= 0 ? 1 : 0;
  int total;
  int rank = block_exclusive_scan(local, &total, s_warp);
  for (int i = begin; i < end; ++i)
    if (obj_idxes[i] >= 0) active_index[rank++] = i;
  if (threadIdx.x == 0) *n_active = total;
}

__global__ void track_gather_rows_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                         const int32_t* __restrict__ n_active,
                                         const int32_t* __restrict__ active_index, int row_words) {
  pdl_trigger();
  pdl_wait();
  const int na = *n_active;
  for (int j = blockIdx.x; j < na; j += gridDim.x) {
    const uint32_t* s = src + static_cast<int64_t>(active_index[j]) * row_words;
    uint32_t* d = dst + static_cast<int64_t>(j) * row_words;
    for (int w = threadIdx.x; w < row_words; w += blockDim.x) d[w] = s[w];
  }
}

}  // namespace moyolo

using namespace moyolo;

extern "C" int64_t moyolo_track_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  return 4 * n + 4 * n + 4 * n * trk_words(n) + 256;
}

extern "C" int moyolo_track_assign(const float* scores, const float* boxes, int64_t* obj_idxes,
                                   int64_t* disappear_time, int64_t* counters, int64_t n, float score_thresh,
                                   float filter_thresh, int miss_tolerance, float iou_thresh, void* workspace,
                                   moyolo_stream_t stream) {
  MOYOLO_REQUIRE(scores && boxes && obj_idxes && disappear_time && counters && workspace, MOYOLO_ERR_BAD_ARG,
                 "track_assign: null pointer");
  MOYOLO_REQUIRE(n >= 0 && n <= kTrkMaxN, MOYOLO_ERR_BAD_SHAPE, "track_assign: n must be in [0, %d], got %lld",
                 kTrkMaxN, (long long)n);
  if (n == 0) return MOYOLO_OK;
  launch_k(track_assign_kernel, dim3(1), dim3(kTrkThreads), 0, static_cast<cudaStream_t>(stream), 
      scores, boxes, obj_idxes, disappear_time, counters, static_cast<int>(n), score_thresh, filter_thresh,
      miss_tolerance, iou_thresh, workspace, nullptr, 0, nullptr, 1);
  return check_launch("track_assign_kernel");
}

extern "C" int moyolo_track_assign_batched(const float* scores, const float* boxes, int64_t* obj_idxes,
                                           int64_t* disappear_time, int64_t* counters, const int32_t* row_offsets,
                                           int n_seq, int64_t max_rows_per_seq, float score_thresh,
                                           float filter_thresh, int miss_tolerance, float iou_thresh, void* workspace,
                                           const int32_t* ctrl, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(scores && boxes && obj_idxes && disappear_time && counters && row_offsets && workspace,
                 MOYOLO_ERR_BAD_ARG, "track_assign_batched: null pointer");
  MOYOLO_REQUIRE(n_seq > 0 && max_rows_per_seq > 0 && max_rows_per_seq <= kTrkMaxN, MOYOLO_ERR_BAD_SHAPE,
                 "track_assign_batched: max_rows_per_seq must be in (0, %d]", kTrkMaxN);
  launch_k(track_assign_kernel, dim3(n_seq), dim3(kTrkThreads), 0, static_cast<cudaStream_t>(stream), 
      scores, boxes, obj_idxes, disappear_time, counters, 0, score_thresh, filter_thresh, miss_tolerance, iou_thresh,
      workspace, row_offsets, moyolo_track_workspace_bytes(max_rows_per_seq), ctrl, 1);
  return check_launch("track_assign_kernel(batched)");
}

extern "C" int moyolo_track_suppress_batched(const float* boxes, int64_t* obj_idxes, int64_t* counters,
                                             const int32_t* row_offsets, int n_seq, int64_t max_rows_per_seq,
                                             float iou_thresh, void* workspace, const int32_t* ctrl,
                                             moyolo_stream_t stream) {
  MOYOLO_REQUIRE(boxes && obj_idxes && counters && row_offsets && workspace, MOYOLO_ERR_BAD_ARG,
                 "track_suppress_batched: null pointer");
  MOYOLO_REQUIRE(n_seq > 0 && max_rows_per_seq > 0 && max_rows_per_seq <= kTrkMaxN, MOYOLO_ERR_BAD_SHAPE,
                 "track_suppress_batched: max_rows_per_seq must be in (0, %d]", kTrkMaxN);
  launch_k(track_assign_kernel, dim3(n_seq), dim3(kTrkThreads), 0, static_cast<cudaStream_t>(stream),
      nullptr, boxes, obj_idxes, nullptr, counters, 0, 0.0f, 0.0f, 0, iou_thresh, workspace, row_offsets,
      moyolo_track_workspace_bytes(max_rows_per_seq), ctrl, 0);
  return check_launch("track_assign_kernel(suppress)");
}

extern "C" int moyolo_track_compact(const int64_t* obj_idxes, int64_t n, int32_t* n_active,
                                    int32_t* active_index, const void* const* src_host, void* const* dst_host,
                                    const int64_t* row_bytes_host, int n_fields, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(obj_idxes && n_active && active_index, MOYOLO_ERR_BAD_ARG, "track_compact: null pointer");
  MOYOLO_REQUIRE(n >= 0 && n <= kTrkMaxN, MOYOLO_ERR_BAD_SHAPE, "track_compact: n must be in [0, %d]", kTrkMaxN);
  MOYOLO_REQUIRE(n_fields == 0 || (src_host && dst_host && row_bytes_host), MOYOLO_ERR_BAD_ARG,
                 "track_compact: null field table");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  launch_k(track_select_kernel, dim3(1), dim3(kTrkThreads), 0, st, obj_idxes, static_cast<int>(n), n_active, active_index);
  int rc = check_launch("track_select_kernel");
  if (rc != MOYOLO_OK || n == 0) return rc;
  for (int f = 0; f < n_fields; ++f) {
    MOYOLO_REQUIRE(row_bytes_host[f] > 0 && row_bytes_host[f] % 4 == 0, MOYOLO_ERR_ALIGNMENT,
                   "track_compact: field %d row size must be a positive multiple of 4 bytes", f);
    const int row_words = static_cast<int>(row_bytes_host[f] / 4);
    const int threads = row_words >= 128 ? 128 : 32;
    launch_k(track_gather_rows_kernel, dim3(static_cast<unsigned>(n < 1184 ? n : 1184)), dim3(threads), 0, st, 
        static_cast<const uint32_t*>(src_host[f]), static_cast<uint32_t*>(dst_host[f]), n_active, active_index,
        row_words);
    rc = check_launch("track_gather_rows_kernel");
    if (rc != MOYOLO_OK) return rc;
  }
  return MOYOLO_OK;
}
