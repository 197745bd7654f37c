// Fixed-Size Query Memory update on the device (MOTR/models/fsqm.py:155-180 `FSQM.online_update`): update the
// confidence of tracked slots, inject new detections into free slots, free the slots that stayed below the
// threshold -- the reference runs three per-element Python loops with `.item()` host reads per query
// (fsqm.py:127-132, 70-98, 104-115) on CPU-resident state. Here: ONE single-CTA kernel, state resident in device
// memory, block scans instead of the sequential "first free slot" / "pop next id" loops:
//   the r-th accepted detection (score > in_threshold, in query order) goes to the r-th free slot (in index order)
//   and takes the r-th id from the head of the FIFO id pool (a ring buffer), for r < min(#accepted, #free);
//   freed ids are appended to the tail in slot order.
// Semantics = the repaired specification F1-F3 of oracle/fsqm_port.py::FsqmSpec (== the shipped class wherever
// that class is self-consistent; pinned on tests/golden/fsqm_clean.npz).
#include "common.cuh"

namespace moyolo {

constexpr int kFsqmThreads = 1024;
constexpr int kFsqmMax = 4096;   // slots / detections per call

__device__ int fsqm_scan(int v, int* total, int* s_warp) {   // exclusive block scan (all threads participate)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int w = s_warp[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    s_warp[lane] = winc - w;
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  *total = s_warp[32];
  return s_warp[warp] + inc - v;
}

__global__ void __launch_bounds__(kFsqmThreads) fsqm_update_kernel(
    int N, int d, float in_thr, float out_thr, int frames, float* __restrict__ mem, float* __restrict__ conf,
    int64_t* __restrict__ ids, float* __restrict__ boxes, int32_t* __restrict__ low, int64_t* __restrict__ pool,
    int32_t* __restrict__ pool_hdr /* {head, count} */, int K, const int64_t* __restrict__ t_ids,
    const float* __restrict__ t_scores, const float* __restrict__ t_boxes, int nd, const float* __restrict__ d_emb,
    const float* __restrict__ d_scores, const float* __restrict__ d_boxes) {
  pdl_trigger();
  pdl_wait();
  __shared__ int s_warp[33];
  __shared__ int s_slot_of_rank[kFsqmMax];   // r-th free slot
  __shared__ int s_det_of_rank[kFsqmMax];    // r-th accepted detection
  const int tid = threadIdx.x;
  const int per_n = (N + kFsqmThreads - 1) / kFsqmThreads, per_d = (nd + kFsqmThreads - 1) / kFsqmThreads;

  // ---- 1. update_confidence (fsqm.py:117-132, repair F1: the slot that holds the id; the last query wins) ----
  for (int s = tid; s < N; s += kFsqmThreads) {
    const int64_t id = ids[s];
    if (id < 0) continue;
    int last = -1;
    for (int i = 0; i < K; ++i)
      if (t_ids[i] == id) last = i;
    if (last >= 0) {
      conf[s] = t_scores[last];
#pragma unroll
      for (int k = 0; k < 4; ++k) boxes[s * 4 + k] = t_boxes[last * 4 + k];
    }
  }
  __syncthreads();

  // ---- 2. inject_new_queries (fsqm.py:46-100) ----
  int head = pool_hdr[0], count = pool_hdr[1];
  int n_free_local = 0, n_acc_local = 0;
  const int nb = min(tid * per_n, N), ne = min(nb + per_n, N);
  const int db = min(tid * per_d, nd), de = min(db + per_d, nd);
  for (int s = nb; s < ne; ++s) n_free_local += ids[s] == -1 ? 1 : 0;
  for (int i = db; i < de; ++i) n_acc_local += d_scores[i] > in_thr ? 1 : 0;
  int n_free, n_acc;
  int fr = fsqm_scan(n_free_local, &n_free, s_warp);
  int ar = fsqm_scan(n_acc_local, &n_acc, s_warp);
  for (int s = nb; s < ne; ++s)
    if (ids[s] == -1) s_slot_of_rank[fr++] = s;
  for (int i = db; i < de; ++i)
    if (d_scores[i] > in_thr) s_det_of_rank[ar++] = i;
  __syncthreads();
  const int n_inj = min(min(n_acc, n_free), count);   // memory full / pool empty: the rest is not injected (:78-81)
  for (int r = tid; r < n_inj; r += kFsqmThreads) {
    const int s = s_slot_of_rank[r], i = s_det_of_rank[r];
    ids[s] = pool[(head + r) % (2 * N)];
    conf[s] = d_scores[i];
    low[s] = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) boxes[s * 4 + k] = d_boxes[i * 4 + k];
  }
  __syncthreads();
  for (int e = tid; e < n_inj * d; e += kFsqmThreads) {   // embeddings of the injected queries
    const int r = e / d, c = e % d;
    mem[static_cast<int64_t>(s_slot_of_rank[r]) * d + c] = d_emb[static_cast<int64_t>(s_det_of_rank[r]) * d + c];
  }
  head = (head + n_inj) % (2 * N);
  count -= n_inj;
  __syncthreads();

  // ---- 3. remove_inactive_queries (fsqm.py:102-115, repairs F2 and F3) ----
  int n_kill_local = 0;
  for (int s = nb; s < ne; ++s) {
    bool kill = false;
    if (ids[s] != -1) {
      if (conf[s] < out_thr) {
        const int l = low[s] + 1;
        low[s] = l;
        kill = l >= frames;
      } else {
        low[s] = 0;
      }
    }
    n_kill_local += kill ? 1 : 0;
    s_slot_of_rank[s] = kill ? 1 : 0;   // (the rank table is free again: reuse it as the kill flag)
  }
  int n_kill;
  int kr = fsqm_scan(n_kill_local, &n_kill, s_warp);
  for (int s = nb; s < ne; ++s) {
    if (s_slot_of_rank[s] == 0) continue;
    pool[(head + count + kr++) % (2 * N)] = ids[s];   // recycled ids queue up behind the unused ones, in slot order
    ids[s] = -1;
    conf[s] = 0.0f;
    low[s] = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) boxes[s * 4 + k] = 0.0f;
  }
  __syncthreads();
  for (int e = tid; e < N * d; e += kFsqmThreads)
    if (s_slot_of_rank[e / d] != 0) mem[e] = 0.0f;
  if (tid == 0) {
    pool_hdr[0] = head;
    pool_hdr[1] = count + n_kill;
  }
}

}  // namespace moyolo

using namespace moyolo;

extern "C" int moyolo_fsqm_update(int max_num_queries, int feature_dim, float in_threshold, float out_threshold,
                                  int consecutive_frames, float* query_memory, float* confidence, int64_t* ids,
                                  float* bounding_boxes, int32_t* consecutive_low_frames, int64_t* id_pool,
                                  int32_t* id_pool_header, int n_track, const int64_t* track_ids,
                                  const float* track_scores, const float* track_boxes, int n_detect,
                                  const float* detect_embedding, const float* detect_scores, const float* detect_boxes,
                                  moyolo_stream_t stream) {
  MOYOLO_REQUIRE(query_memory && confidence && ids && bounding_boxes && consecutive_low_frames && id_pool && id_pool_header,
                 MOYOLO_ERR_BAD_ARG, "fsqm_update: null state pointer");
  MOYOLO_REQUIRE(max_num_queries > 0 && max_num_queries <= kFsqmMax && feature_dim > 0, MOYOLO_ERR_BAD_SHAPE,
                 "fsqm_update: max_num_queries must be in [1, %d]", kFsqmMax);
  MOYOLO_REQUIRE(n_track >= 0 && n_detect >= 0 && n_detect <= kFsqmMax, MOYOLO_ERR_BAD_SHAPE,
                 "fsqm_update: n_detect must be in [0, %d]", kFsqmMax);
  MOYOLO_REQUIRE((n_track == 0 || (track_ids && track_scores && track_boxes)) &&
                     (n_detect == 0 || (detect_embedding && detect_scores && detect_boxes)),
                 MOYOLO_ERR_BAD_ARG, "fsqm_update: null query pointer");
  launch_k(fsqm_update_kernel, dim3(1), dim3(kFsqmThreads), 0, static_cast<cudaStream_t>(stream), max_num_queries,
           feature_dim, in_threshold, out_threshold, consecutive_frames, query_memory, confidence, ids, bounding_boxes,
           consecutive_low_frames, id_pool, id_pool_header, n_track, track_ids, track_scores, track_boxes, n_detect,
           detect_embedding, detect_scores, detect_boxes);
  return check_launch("fsqm_update_kernel");
}
