// Device-resident frame state for lock-step sequences: query assembly, active-track compaction and
// state write-back as single launches over all sequences, so a whole frame (assembly -> decoder ->
// ID assignment -> QIM update -> write-back) has static launch shapes for a padded row count and can
// be replayed as one CUDA graph with the per-sequence track counts living in device memory.
//
// Reference semantics:
//   assemble : head.py:1056-1064,1108-1109 (tracks first, then detect queries), head.py:888-900
//              (track content = class embedding of argmax previous logits), transformer.py:183-190
//              (pos2posemb of the detect boxes) and repair R2 (ids = cat(prev, -1), disappear = cat(prev, 0)).
//   compact  : MOTR/models/qim.py:184-187 (ids >= 0) + instances.py:152-178 (row selection).
//   writeback: qim.py:298-300 (query_pos <- QIM output, ref_pts <- inverse_sigmoid(pred_boxes)).
#include "common.cuh"

namespace moyolo {

// control block shared by the frame kernels (device int32[8])
enum { kCtrlAbort = 0, kCtrlFrame = 1, kCtrlCursor = 2, kCtrlTableOverflow = 3, kCtrlAbortRows = 4,
       kCtrlTrackOverflow = 5 };  // sticky: a sequence had more active tracks than the state capacity `cap`

__device__ __forceinline__ float inv_sigmoid_(float x) {
  x = fminf(fmaxf(x, 0.0f), 1.0f);
  return logf(fmaxf(x, 1e-5f) / fmaxf(1.0f - x, 1e-5f));
}

// grid = rows_pad blocks (one row per block), C threads cover the embedding row.
__global__ void frame_assemble_kernel(int n_seq, int n_detect, int C, int cap, const int32_t* __restrict__ n_tracks,
                                      const float* __restrict__ t_ref, const float* __restrict__ t_qpos,
                                      const int32_t* __restrict__ t_label, const int64_t* __restrict__ t_ids,
                                      const int64_t* __restrict__ t_dis, const float* __restrict__ class_embed,
                                      const float* __restrict__ det_embed, const float* __restrict__ det_refer,
                                      float* __restrict__ x, float* __restrict__ refer_logit, float* __restrict__ pos,
                                      int64_t* __restrict__ ids, int64_t* __restrict__ dis,
                                      int32_t* __restrict__ row_offsets, int num_pos_feats, float temperature,
                                      int rows_pad, int32_t* __restrict__ ctrl, float* __restrict__ refer_sig,
                                      void* __restrict__ x_lp, void* __restrict__ xq_lp, int lp_bf16) {
  pdl_trigger();
  // pos2posemb denominators (transformer.py:185-186), one table per CTA, computed before the dependency wait
  __shared__ float s_dimt[256];
  for (int i = threadIdx.x; i < num_pos_feats && i < 256; i += blockDim.x)
    s_dimt[i] = powf(temperature, static_cast<float>(2 * (i / 2)) / static_cast<float>(num_pos_feats));
  __syncthreads();
  pdl_wait();
  const int row = blockIdx.x;
  // optional fused outputs: sigmoid(refer) (transformer.py:690) and the GEMM operand copies x, x + pos
  auto emit_lp = [&](int c, float xv, float pv) {
    const int64_t o = static_cast<int64_t>(row) * C + c;
    if (lp_bf16) {
      if (x_lp) static_cast<__nv_bfloat16*>(x_lp)[o] = __float2bfloat16_rn(xv);
      if (xq_lp) static_cast<__nv_bfloat16*>(xq_lp)[o] = __float2bfloat16_rn(xv + pv);
    } else {
      if (x_lp) static_cast<float*>(x_lp)[o] = xv;
      if (xq_lp) static_cast<float*>(xq_lp)[o] = xv + pv;
    }
  };
  auto emit_ref = [&](int k, float logit) {
    refer_logit[row * 4 + k] = logit;
    if (refer_sig) refer_sig[row * 4 + k] = 1.0f / (1.0f + expf(-logit));
  };
  // Speculative launch guard: the host picks rows_pad from the track counts of an EARLIER frame. If the
  // real row count does not fit (or an earlier frame already aborted), every block consistently
  // builds a detect-only frame (in bounds, results discarded) and the sticky abort flag tells the
  // state-writing kernels of this and later frames to do nothing until the host re-launches.
  int total_rows = 0;
  for (int i = 0; i < n_seq; ++i) total_rows += n_tracks[i] + n_detect;
  const bool aborted = ctrl != nullptr && (ctrl[kCtrlAbort] != 0 || total_rows > rows_pad);
  // locate the sequence of this row: offsets are the running sum of (T_s + n_detect)
  int s = 0, off = 0, T = 0;
  for (; s < n_seq; ++s) {
    T = aborted ? 0 : n_tracks[s];
    if (row < off + T + n_detect) break;
    off += T + n_detect;
  }
  if (row == 0 && threadIdx.x == 0) {
    int acc = 0;
    row_offsets[0] = 0;
    for (int i = 0; i < n_seq; ++i) {
      acc += (aborted ? 0 : n_tracks[i]) + n_detect;
      row_offsets[i + 1] = acc;
    }
    if (aborted && ctrl[kCtrlAbort] == 0) {
      ctrl[kCtrlAbortRows] = total_rows;
      ctrl[kCtrlAbort] = 1;
    }
  }
  float* xr = x + static_cast<int64_t>(row) * C;
  float* pr = pos + static_cast<int64_t>(row) * C;
  if (s == n_seq) {  // padding row: finite zeros, never an object
    for (int c = threadIdx.x; c < C; c += blockDim.x) { xr[c] = 0.0f; pr[c] = 0.0f; emit_lp(c, 0.0f, 0.0f); }
    if (threadIdx.x < 4) emit_ref(threadIdx.x, 0.0f);
    if (threadIdx.x == 0) { ids[row] = -1; dis[row] = 0; }
    return;
  }
  const int j = row - off;
  if (j < T) {  // carried track
    const int64_t src = static_cast<int64_t>(s) * cap + j;
    const float* ce = class_embed + static_cast<int64_t>(t_label[src]) * C;
    const float* qp = t_qpos + src * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float xv = ce[c], pv = qp[c];
      xr[c] = xv;
      pr[c] = pv;
      emit_lp(c, xv, pv);
    }
    if (threadIdx.x < 4) emit_ref(threadIdx.x, t_ref[src * 4 + threadIdx.x]);
    if (threadIdx.x == 0) { ids[row] = t_ids[src]; dis[row] = t_dis[src]; }
  } else {      // detect query
    const int64_t src = static_cast<int64_t>(s) * n_detect + (j - T);
    const float* de = det_embed + src * C;
    const float* dr = det_refer + src * 4;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float xv = de[c];
      xr[c] = xv;
      const int coord = c / num_pos_feats, i = c % num_pos_feats;
      const float p = dr[coord] * 6.283185307179586f;
      const float e = p / s_dimt[i];
      const float pv = (i & 1) ? cosf(e) : sinf(e);
      pr[c] = pv;
      emit_lp(c, xv, pv);
    }
    if (threadIdx.x < 4) emit_ref(threadIdx.x, dr[threadIdx.x]);
    if (threadIdx.x == 0) { ids[row] = -1; dis[row] = 0; }
  }
}

__device__ int block_exclusive_scan_f(int v, int* total, int* s_warp) {
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
    int w = (lane < (blockDim.x >> 5)) ? s_warp[lane] : 0;
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

// grid = (n_seq, kCompactChunks), 1024 threads: select ids >= 0 in order, then copy every field the next
// frame needs. Every CTA of a sequence repeats the (cheap) selection scan into shared memory and then
// copies the compact rows j = blockIdx.y, blockIdx.y + gridDim.y, ...; CTA y == 0 also writes the selection.
constexpr int kCompactChunks = 8;
constexpr int kCompactMaxRows = 4096;

__global__ void __launch_bounds__(1024) frame_compact_kernel(
    int C, int cap, const int32_t* __restrict__ row_offsets, const int64_t* __restrict__ ids,
    const int64_t* __restrict__ dis, const int32_t* __restrict__ labels, const float* __restrict__ refer_logit,
    const float* __restrict__ pos, const float* __restrict__ hs, const float* __restrict__ boxes,
    int32_t* __restrict__ n_active, int32_t* __restrict__ active_index, float* __restrict__ c_ref,
    float* __restrict__ c_pos, float* __restrict__ c_hs, float* __restrict__ c_box, int32_t* __restrict__ t_label,
    int64_t* __restrict__ t_ids, int64_t* __restrict__ t_dis, int32_t* __restrict__ ctrl,
    void* __restrict__ q_qk_lp, void* __restrict__ q_tgt_lp, int lp_bf16, int num_pos_feats, float temperature) {
  pdl_trigger();
  pdl_wait();
  __shared__ int s_warp[33];
  __shared__ int16_t s_sel[kCompactMaxRows];  // row (within the sequence) of the j-th active track
  if (ctrl != nullptr && ctrl[kCtrlAbort] != 0) return;  // aborted frame: leave the track state untouched
  const int s = blockIdx.x;
  const bool first = blockIdx.y == 0;
  const int off = row_offsets[s];
  const int n = min(row_offsets[s + 1] - off, kCompactMaxRows);
  const int per = (n + blockDim.x - 1) / blockDim.x;
  const int begin = min(static_cast<int>(threadIdx.x) * per, n);
  const int end = min(begin + per, n);
  int local = 0;
  for (int i = begin; i < end; ++i) local += ids[off + i] >= 0 ? 1 : 0;
  int total;
  int rank = block_exclusive_scan_f(local, &total, s_warp);
  // more active tracks than the carried state holds: the surplus would silently lose its identity (the reference
  // has no such limit) -> sticky flag, the host raises (TrackEngine.collect / track_table)
  if (total > cap && first && threadIdx.x == 0 && ctrl != nullptr) ctrl[kCtrlTrackOverflow] = 1;
  total = min(total, cap);
  for (int i = begin; i < end; ++i)
    if (ids[off + i] >= 0) {
      if (rank < cap) {
        s_sel[rank] = static_cast<int16_t>(i);
        if (first) active_index[off + rank] = i;
      }
      ++rank;
    }
  if (first && threadIdx.x == 0) n_active[s] = total;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const bool want_q = q_qk_lp != nullptr || q_tgt_lp != nullptr;
  for (int j = blockIdx.y * nwarps + warp; j < total; j += gridDim.y * nwarps) {
    const int64_t src = off + s_sel[j];
    const int64_t dst = off + j;  // compact rows keep the frame layout: sequence s starts at row_offsets[s]
    for (int c = lane; c < C; c += 32) {
      c_pos[dst * C + c] = pos[src * C + c];
      const float h = hs[src * C + c];
      c_hs[dst * C + c] = h;
      // QIM operands (qim.py:255, 271): q = k = tgt + pos2posemb(ref_pts), v = tgt
      if (want_q) {
        const int coord = c / num_pos_feats, i = c % num_pos_feats;
        const float p = refer_logit[src * 4 + coord] * 6.283185307179586f;
        const float e = p / powf(temperature, static_cast<float>(2 * (i / 2)) / static_cast<float>(num_pos_feats));
        const float qp = (i & 1) ? cosf(e) : sinf(e);
        if (lp_bf16) {
          if (q_qk_lp) static_cast<__nv_bfloat16*>(q_qk_lp)[dst * C + c] = __float2bfloat16_rn(qp + h);
          if (q_tgt_lp) static_cast<__nv_bfloat16*>(q_tgt_lp)[dst * C + c] = __float2bfloat16_rn(h);
        } else {
          if (q_qk_lp) static_cast<float*>(q_qk_lp)[dst * C + c] = qp + h;
          if (q_tgt_lp) static_cast<float*>(q_tgt_lp)[dst * C + c] = h;
        }
      }
    }
    if (lane < 4) {
      c_ref[dst * 4 + lane] = refer_logit[src * 4 + lane];
      c_box[dst * 4 + lane] = boxes[src * 4 + lane];
    }
    if (lane == 0) {
      const int64_t st = static_cast<int64_t>(s) * cap + j;
      t_label[st] = labels[src];
      t_ids[st] = ids[src];
      t_dis[st] = dis[src];
    }
  }
}

// grid = (n_seq, kCompactChunks), 256 threads. The ID assignment of RuntimeTrackerBase.update
// (head.py:1232-1243) fused with the active-track compaction: every CTA of a sequence repeats the (cheap,
// deterministic) assignment scan -- `max_obj_id++` in query order == exclusive prefix count of the new rows; the
// packed scan carries the count of active rows in its high half -- and then copies the compact rows
// j = blockIdx.y*32 + warp, ...; CTA y == 0 also writes the updated ids / disappear counters (to separate output
// arrays: the other CTAs still read the inputs) and the selection. The duplicate filter + renumbering of the
// reference only changes the ID counters (its filtered copy is dropped, head.py:1268-1283): it runs afterwards,
// off the critical path, as moyolo_track_suppress_batched.
constexpr int kAssignThreads = 256;  // 8 warps: cheap block scans; each thread owns ceil(n/256) consecutive rows

__global__ void __launch_bounds__(kAssignThreads) frame_assign_compact_kernel(
    int C, int cap, int rows_pad, const int32_t* __restrict__ row_offsets, const float* __restrict__ scores,
    const int64_t* __restrict__ ids_in, const int64_t* __restrict__ dis_in, const int64_t* __restrict__ counters,
    float score_thresh, float filter_thresh, int miss_tolerance, int64_t* __restrict__ ids_out,
    int64_t* __restrict__ dis_out, const int32_t* __restrict__ labels, const float* __restrict__ refer_logit,
    const float* __restrict__ pos, const float* __restrict__ hs, const float* __restrict__ boxes,
    int32_t* __restrict__ n_active, int32_t* __restrict__ active_index, float* __restrict__ c_ref,
    float* __restrict__ c_pos, float* __restrict__ c_hs, float* __restrict__ c_box, int32_t* __restrict__ t_label,
    int64_t* __restrict__ t_ids, int64_t* __restrict__ t_dis, int32_t* __restrict__ ctrl,
    void* __restrict__ q_qk_lp, void* __restrict__ q_tgt_lp, int lp_bf16, int num_pos_feats, float temperature) {
  pdl_trigger();
  __shared__ int s_warp[33];
  __shared__ int16_t s_sel[kCompactMaxRows];  // row (within the sequence) of the j-th active track
  __shared__ float s_dimt[256];               // pos2posemb denominators (transformer.py:185-186)
  for (int i = threadIdx.x; i < num_pos_feats; i += blockDim.x)
    s_dimt[i] = powf(temperature, static_cast<float>(2 * (i / 2)) / static_cast<float>(num_pos_feats));
  pdl_wait();
  if (ctrl != nullptr && ctrl[kCtrlAbort] != 0) return;  // aborted frame: leave the track state untouched
  const int s = blockIdx.x;
  const bool first = blockIdx.y == 0;
  const int off = row_offsets[s];
  const int n = min(row_offsets[s + 1] - off, kCompactMaxRows);
  const int64_t max_obj_id = counters[2 * s];
  const int per = (n + blockDim.x - 1) / blockDim.x;
  const int begin = min(static_cast<int>(threadIdx.x) * per, n);
  const int end = min(begin + per, n);
  // packed per-thread counts: low 16 bits = new objects, high bits = rows that are active after the update
  int local = 0;
  for (int i = begin; i < end; ++i) {
    const int64_t id = ids_in[off + i];
    const float sc = scores[off + i];
    const bool is_new = id == -1 && sc >= score_thresh;
    const bool dies = id >= 0 && sc < filter_thresh && dis_in[off + i] + 1 >= miss_tolerance;
    const bool alive = is_new || (id >= 0 && !dies);
    local += (is_new ? 1 : 0) + (alive ? 65536 : 0);
  }
  int total;
  const int rank = block_exclusive_scan_f(local, &total, s_warp);
  int new_rank = rank & 0xffff, act_rank = rank >> 16;
  if ((total >> 16) > cap && first && threadIdx.x == 0 && ctrl != nullptr) ctrl[kCtrlTrackOverflow] = 1;  // see frame_compact
  const int n_act = min(total >> 16, cap);
  for (int i = begin; i < end; ++i) {
    int64_t id = ids_in[off + i];
    int64_t dt = dis_in[off + i];
    const float sc = scores[off + i];
    if (id == -1 && sc >= score_thresh) {
      id = max_obj_id + new_rank++;
    } else if (id >= 0 && sc < filter_thresh) {
      dt += 1;
      if (dt >= miss_tolerance) id = -1;
    }
    if (first) { ids_out[off + i] = id; dis_out[off + i] = dt; }
    if (id >= 0) {
      if (act_rank < cap) {
        s_sel[act_rank] = static_cast<int16_t>(i);
        if (first) {
          active_index[off + act_rank] = i;
          const int64_t st = static_cast<int64_t>(s) * cap + act_rank;
          t_label[st] = labels[off + i];
          t_ids[st] = id;
          t_dis[st] = dt;
        }
      }
      ++act_rank;
    }
  }
  if (first && threadIdx.x == 0) n_active[s] = n_act;
  if (first && s == static_cast<int>(gridDim.x) - 1)  // padding rows of the frame: never an object
    for (int r = row_offsets[s + 1] + threadIdx.x; r < rows_pad; r += blockDim.x) { ids_out[r] = -1; dis_out[r] = 0; }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const bool want_q = q_qk_lp != nullptr || q_tgt_lp != nullptr;
  for (int j = blockIdx.y * nwarps + warp; j < n_act; j += gridDim.y * nwarps) {
    const int64_t src = off + s_sel[j];
    const int64_t dst = off + j;  // compact rows keep the frame layout: sequence s starts at row_offsets[s]
    const float4 rl = *reinterpret_cast<const float4*>(refer_logit + src * 4);
    for (int c = lane * 4; c < C; c += 128) {
      const float4 pv = *reinterpret_cast<const float4*>(pos + src * C + c);
      const float4 hv = *reinterpret_cast<const float4*>(hs + src * C + c);
      *reinterpret_cast<float4*>(c_pos + dst * C + c) = pv;
      *reinterpret_cast<float4*>(c_hs + dst * C + c) = hv;
      // QIM operands (qim.py:255, 271): q = k = tgt + pos2posemb(ref_pts), v = tgt
      if (want_q) {
        const float h4[4] = {hv.x, hv.y, hv.z, hv.w};
        float qk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int cc = c + u, coord = cc / num_pos_feats, i = cc % num_pos_feats;
          const float r = coord == 0 ? rl.x : (coord == 1 ? rl.y : (coord == 2 ? rl.z : rl.w));
          const float p = r * 6.283185307179586f;
          const float e = p / s_dimt[i];
          qk[u] = ((i & 1) ? cosf(e) : sinf(e)) + h4[u];
        }
        if (lp_bf16) {
          if (q_qk_lp)
            *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(q_qk_lp) + dst * C + c) =
                make_uint2(float2_to_bf16x2(qk[0], qk[1]), float2_to_bf16x2(qk[2], qk[3]));
          if (q_tgt_lp)
            *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(q_tgt_lp) + dst * C + c) =
                make_uint2(float2_to_bf16x2(h4[0], h4[1]), float2_to_bf16x2(h4[2], h4[3]));
        } else {
          if (q_qk_lp) *reinterpret_cast<float4*>(static_cast<float*>(q_qk_lp) + dst * C + c) = make_float4(qk[0], qk[1], qk[2], qk[3]);
          if (q_tgt_lp) *reinterpret_cast<float4*>(static_cast<float*>(q_tgt_lp) + dst * C + c) = hv;
        }
      }
    }
    if (lane == 0) {
      *reinterpret_cast<float4*>(c_ref + dst * 4) = rl;
      // boxes == NULL: the caller lets frame_writeback gather the refined boxes itself, so that this kernel does not
      // have to wait for the last layer's box head
      if (boxes != nullptr) *reinterpret_cast<float4*>(c_box + dst * 4) = *reinterpret_cast<const float4*>(boxes + src * 4);
    }
  }
}

// grid = (n_seq, chunks): t_qpos[s, j] = new_qpos[off_s + j]; t_ref[s, j] = inverse_sigmoid(c_box[off_s + j]).
__global__ void frame_writeback_kernel(int C, int cap, const int32_t* __restrict__ row_offsets,
                                       const int32_t* __restrict__ n_active, const float* __restrict__ new_qpos,
                                       const float* __restrict__ c_box, float* __restrict__ t_qpos,
                                       float* __restrict__ t_ref, int32_t* __restrict__ n_tracks,
                                       const int32_t* __restrict__ ctrl, int32_t* __restrict__ info,
                                       const float* __restrict__ boxes, const int32_t* __restrict__ active_index) {
  pdl_trigger();
  pdl_wait();
  // host-visible frame summary: [n_active (n_seq) | ctrl (8)], written even for an aborted frame
  if (info != nullptr && blockIdx.y == 0 && threadIdx.x < 32) {
    if (threadIdx.x == 0) info[blockIdx.x] = n_active[blockIdx.x];
    if (blockIdx.x == 0 && threadIdx.x < 8 && ctrl != nullptr) info[gridDim.x + threadIdx.x] = ctrl[threadIdx.x];
  }
  if (ctrl != nullptr && ctrl[kCtrlAbort] != 0) return;
  const int s = blockIdx.x;
  const int off = row_offsets[s];
  const int k = n_active[s];
  const int warp = (blockIdx.y * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.y * blockDim.x) >> 5;
  for (int j = warp; j < k; j += nwarps) {
    const int64_t src = off + j, dst = static_cast<int64_t>(s) * cap + j;
    for (int c = lane; c < C; c += 32) t_qpos[dst * C + c] = new_qpos[src * C + c];
    // boxes of the active rows: already compacted (c_box) or gathered here through the selection (boxes, active_index)
    if (lane < 4) {
      const float bx = boxes != nullptr ? boxes[(static_cast<int64_t>(off) + active_index[src]) * 4 + lane] : c_box[src * 4 + lane];
      t_ref[dst * 4 + lane] = inv_sigmoid_(bx);
    }
  }
  if (blockIdx.y == 0 && threadIdx.x == 0) n_tracks[s] = k;
}

// One CTA: (a) packs every row of the frame as [id, cx, cy, w, h, score, label, seq] fp32 into `frame_rows`
// (the per-frame result a host reads back with ONE copy), (b) appends the tracked objects (ids >= 0,
// the compacted order of frame_compact) to the device-resident track table
// [seq, frame, id, cx, cy, w, h, score, cls] at the cursor held in ctrl, (c) advances the frame counter.
__global__ void __launch_bounds__(256) frame_emit_kernel(
    int n_seq, int rows_pad, const int32_t* __restrict__ row_offsets, const int64_t* __restrict__ ids,
    const float* __restrict__ boxes, const float* __restrict__ scores, const int32_t* __restrict__ labels,
    const int32_t* __restrict__ n_active, const int32_t* __restrict__ active_index,
    const int32_t* __restrict__ seq_ids, float* __restrict__ frame_rows, float* __restrict__ table, int table_cap,
    int32_t* __restrict__ ctrl) {
  pdl_trigger();
  pdl_wait();
  if (ctrl[kCtrlAbort] != 0) return;
  const int total = row_offsets[n_seq];
  for (int r = threadIdx.x; r < rows_pad; r += blockDim.x) {
    float* o = frame_rows + static_cast<int64_t>(r) * 8;
    if (r < total) {
      int s = 0;
      while (s + 1 < n_seq && r >= row_offsets[s + 1]) ++s;
      const float4 b = *reinterpret_cast<const float4*>(boxes + static_cast<int64_t>(r) * 4);
      *reinterpret_cast<float4*>(o) = make_float4(static_cast<float>(ids[r]), b.x, b.y, b.z);
      *reinterpret_cast<float4*>(o + 4) = make_float4(b.w, scores[r], static_cast<float>(labels[r]),
                                                      static_cast<float>(seq_ids[s]));
    } else {
      *reinterpret_cast<float4*>(o) = make_float4(-1.0f, 0.0f, 0.0f, 0.0f);
      *reinterpret_cast<float4*>(o + 4) = make_float4(0.0f, 0.0f, 0.0f, -1.0f);
    }
  }
  const int frame = ctrl[kCtrlFrame];
  int base = ctrl[kCtrlCursor];
  bool overflow = false;
  for (int s = 0; s < n_seq; ++s) {
    const int off = row_offsets[s], k = n_active[s];
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      const int64_t src = off + active_index[off + j];
      const int dst = base + j;
      if (dst < table_cap) {
        float* t = table + static_cast<int64_t>(dst) * 9;
        t[0] = static_cast<float>(seq_ids[s]);
        t[1] = static_cast<float>(frame);
        t[2] = static_cast<float>(ids[src]);
        t[3] = boxes[src * 4 + 0];
        t[4] = boxes[src * 4 + 1];
        t[5] = boxes[src * 4 + 2];
        t[6] = boxes[src * 4 + 3];
        t[7] = scores[src];
        t[8] = static_cast<float>(labels[src]);
      } else {
        overflow = true;
      }
    }
    base += k;
  }
  const int any_overflow = __syncthreads_or(overflow ? 1 : 0);  // also orders the reads of ctrl above
  if (threadIdx.x == 0) {
    ctrl[kCtrlCursor] = base < table_cap ? base : table_cap;
    if (any_overflow) ctrl[kCtrlTableOverflow] = 1;
    ctrl[kCtrlFrame] = frame + 1;
  }
}

// ---- final track-row gather (moyolo_b200/sharding.py): fixed-capacity buffers, merged by rank offset ----
// send = [header | rows]: header row = {count, overflow flag, 0...}; one launch.
__global__ void table_pack_kernel(const float* __restrict__ rows, int64_t n_rows, const int32_t* __restrict__ n_rows_dev,
                                  const int32_t* __restrict__ overflow_dev, int64_t cap, float* __restrict__ send) {
  pdl_trigger();
  pdl_wait();
  if (n_rows_dev != nullptr) n_rows = *n_rows_dev;   // the engine's table cursor: no host read before the gather
  const bool over = n_rows > cap || (overflow_dev != nullptr && *overflow_dev != 0);
  const int64_t n = n_rows < cap ? n_rows : cap;
  const int64_t total = (n + 1) * 9;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (i < 9) send[i] = i == 0 ? static_cast<float>(n) : (i == 1 && over ? 1.0f : 0.0f);
    else send[i] = rows[i - 9];
  }
}
// recv = [world][cap + 1][9] -> out rows ordered by rank (exclusive prefix sum of the header counts); info = {total
// rows, number of ranks whose buffer overflowed}. One launch, nothing is read back.
__global__ void table_merge_kernel(const float* __restrict__ recv, int world, int64_t cap, float* __restrict__ out,
                                   int32_t* __restrict__ info) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.y;
  int64_t off = 0;
  int over = 0;
  for (int q = 0; q < world; ++q) {
    const float* h = recv + static_cast<int64_t>(q) * (cap + 1) * 9;
    if (q < r) off += static_cast<int64_t>(h[0] + 0.5f);
    over += h[1] > 0.5f ? 1 : 0;
  }
  const float* src = recv + (static_cast<int64_t>(r) * (cap + 1) + 1) * 9;
  const int64_t n = static_cast<int64_t>(recv[static_cast<int64_t>(r) * (cap + 1) * 9] + 0.5f);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n * 9;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[off * 9 + i] = src[i];
  if (r == world - 1 && blockIdx.x == 0 && threadIdx.x == 0) {
    info[0] = static_cast<int32_t>(off + n);
    info[1] = over;
  }
}

}  // namespace moyolo

using namespace moyolo;

extern "C" int moyolo_frame_assemble(int n_seq, int n_detect, int C, int cap, const int32_t* n_tracks,
                                     const float* t_ref, const float* t_qpos, const int32_t* t_label,
                                     const int64_t* t_ids, const int64_t* t_dis, const float* class_embed,
                                     const float* det_embed, const float* det_refer, float* x, float* refer_logit,
                                     float* pos, int64_t* ids, int64_t* dis, int32_t* row_offsets, int64_t rows_pad,
                                     int num_pos_feats, float temperature, int32_t* ctrl, float* refer_sig,
                                     void* x_lp, void* xq_lp, int lp_dtype, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(n_tracks && t_ref && t_qpos && t_label && t_ids && t_dis && class_embed && det_embed && det_refer &&
                     x && refer_logit && pos && ids && dis && row_offsets,
                 MOYOLO_ERR_BAD_ARG, "frame_assemble: null pointer");
  MOYOLO_REQUIRE(n_seq > 0 && n_detect >= 0 && C > 0 && cap > 0 && rows_pad > 0, MOYOLO_ERR_BAD_SHAPE,
                 "frame_assemble: bad sizes");
  MOYOLO_REQUIRE(C == 4 * num_pos_feats && num_pos_feats <= 256, MOYOLO_ERR_BAD_SHAPE,
                 "frame_assemble: C must equal 4*num_pos_feats (<= 1024)");
  MOYOLO_REQUIRE(lp_dtype == MOYOLO_BF16 || lp_dtype == MOYOLO_F32, MOYOLO_ERR_UNSUPPORTED,
                 "frame_assemble: lp_dtype must be F32 or BF16");
  const int threads = C >= 256 ? 256 : (C >= 128 ? 128 : 64);
  launch_k(frame_assemble_kernel, dim3(static_cast<unsigned>(rows_pad)), dim3(threads), 0, static_cast<cudaStream_t>(stream), 
      n_seq, n_detect, C, cap, n_tracks, t_ref, t_qpos, t_label, t_ids, t_dis, class_embed, det_embed, det_refer, x,
      refer_logit, pos, ids, dis, row_offsets, num_pos_feats, temperature, static_cast<int>(rows_pad), ctrl,
      refer_sig, x_lp, xq_lp, lp_dtype == MOYOLO_BF16 ? 1 : 0);
  return check_launch("frame_assemble_kernel");
}

extern "C" int moyolo_frame_compact(int n_seq, int C, int cap, const int32_t* row_offsets, const int64_t* ids,
                                    const int64_t* dis, const int32_t* labels, const float* refer_logit,
                                    const float* pos, const float* hs, const float* boxes, int32_t* n_active,
                                    int32_t* active_index, float* c_ref, float* c_pos, float* c_hs, float* c_box,
                                    int32_t* t_label, int64_t* t_ids, int64_t* t_dis, int32_t* ctrl,
                                    void* q_qk_lp, void* q_tgt_lp, int lp_dtype, int num_pos_feats,
                                    float temperature, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(row_offsets && ids && dis && labels && refer_logit && pos && hs && boxes && n_active &&
                     active_index && c_ref && c_pos && c_hs && c_box && t_label && t_ids && t_dis,
                 MOYOLO_ERR_BAD_ARG, "frame_compact: null pointer");
  MOYOLO_REQUIRE(n_seq > 0 && C > 0 && cap > 0, MOYOLO_ERR_BAD_SHAPE, "frame_compact: bad sizes");
  MOYOLO_REQUIRE((q_qk_lp == nullptr && q_tgt_lp == nullptr) || C == 4 * num_pos_feats, MOYOLO_ERR_BAD_SHAPE,
                 "frame_compact: QIM operands need C == 4*num_pos_feats");
  launch_k(frame_compact_kernel, dim3(n_seq, kCompactChunks), dim3(1024), 0, static_cast<cudaStream_t>(stream), 
      C, cap, row_offsets, ids, dis, labels, refer_logit, pos, hs, boxes, n_active, active_index, c_ref, c_pos, c_hs,
      c_box, t_label, t_ids, t_dis, ctrl, q_qk_lp, q_tgt_lp, lp_dtype == MOYOLO_BF16 ? 1 : 0, num_pos_feats,
      temperature);
  return check_launch("frame_compact_kernel");
}

extern "C" int moyolo_frame_assign_compact(int n_seq, int C, int cap, int64_t rows_pad, const int32_t* row_offsets,
                                           const float* scores, const int64_t* ids_in, const int64_t* dis_in,
                                           const int64_t* counters, float score_thresh, float filter_thresh,
                                           int miss_tolerance, int64_t* ids_out, int64_t* dis_out,
                                           const int32_t* labels, const float* refer_logit, const float* pos,
                                           const float* hs, const float* boxes, int32_t* n_active,
                                           int32_t* active_index, float* c_ref, float* c_pos, float* c_hs,
                                           float* c_box, int32_t* t_label, int64_t* t_ids, int64_t* t_dis,
                                           int32_t* ctrl, void* q_qk_lp, void* q_tgt_lp, int lp_dtype,
                                           int num_pos_feats, float temperature, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(row_offsets && scores && ids_in && dis_in && counters && ids_out && dis_out && labels &&
                     refer_logit && pos && hs && n_active && active_index && c_ref && c_pos && c_hs &&
                     (boxes == nullptr || c_box) && t_label && t_ids && t_dis,
                 MOYOLO_ERR_BAD_ARG, "frame_assign_compact: null pointer");
  MOYOLO_REQUIRE(ids_in != ids_out && dis_in != dis_out, MOYOLO_ERR_BAD_ARG,
                 "frame_assign_compact: ids/dis outputs must not alias the inputs");
  MOYOLO_REQUIRE(n_seq > 0 && C > 0 && C % 4 == 0 && cap > 0 && rows_pad > 0, MOYOLO_ERR_BAD_SHAPE,
                 "frame_assign_compact: bad sizes");
  MOYOLO_REQUIRE(C == 4 * num_pos_feats && num_pos_feats <= 256, MOYOLO_ERR_BAD_SHAPE,
                 "frame_assign_compact: C must equal 4*num_pos_feats (<= 1024)");
  MOYOLO_REQUIRE(aligned16(refer_logit) && aligned16(pos) && aligned16(hs) && aligned16(boxes) && aligned16(c_ref) &&
                     aligned16(c_pos) && aligned16(c_hs) && aligned16(c_box) && aligned16(q_qk_lp) && aligned16(q_tgt_lp),
                 MOYOLO_ERR_ALIGNMENT, "frame_assign_compact: row buffers must be 16-byte aligned");
  launch_k(frame_assign_compact_kernel, dim3(n_seq, kCompactChunks), dim3(kAssignThreads), 0, static_cast<cudaStream_t>(stream),
      C, cap, static_cast<int>(rows_pad), row_offsets, scores, ids_in, dis_in, counters, score_thresh, filter_thresh,
      miss_tolerance, ids_out, dis_out, labels, refer_logit, pos, hs, boxes, n_active, active_index, c_ref, c_pos,
      c_hs, c_box, t_label, t_ids, t_dis, ctrl, q_qk_lp, q_tgt_lp, lp_dtype == MOYOLO_BF16 ? 1 : 0, num_pos_feats,
      temperature);
  return check_launch("frame_assign_compact_kernel");
}

extern "C" int moyolo_frame_writeback(int n_seq, int C, int cap, const int32_t* row_offsets, const int32_t* n_active,
                                      const float* new_qpos, const float* c_box, float* t_qpos, float* t_ref,
                                      int32_t* n_tracks, const int32_t* ctrl, int32_t* info, const float* boxes,
                                      const int32_t* active_index, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(row_offsets && n_active && new_qpos && t_qpos && t_ref && n_tracks &&
                     (c_box != nullptr || (boxes != nullptr && active_index != nullptr)),
                 MOYOLO_ERR_BAD_ARG, "frame_writeback: null pointer (needs c_box, or boxes + active_index)");
  MOYOLO_REQUIRE(n_seq > 0 && C > 0 && cap > 0, MOYOLO_ERR_BAD_SHAPE, "frame_writeback: bad sizes");
  dim3 grid(n_seq, 8);
  launch_k(frame_writeback_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), C, cap, row_offsets, n_active, new_qpos,
           c_box, t_qpos, t_ref, n_tracks, ctrl, info, c_box != nullptr ? nullptr : boxes, active_index);
  return check_launch("frame_writeback_kernel");
}

extern "C" int moyolo_frame_emit(int n_seq, int64_t rows_pad, const int32_t* row_offsets, const int64_t* ids,
                                 const float* boxes, const float* scores, const int32_t* labels,
                                 const int32_t* n_active, const int32_t* active_index, const int32_t* seq_ids,
                                 float* frame_rows, float* table, int64_t table_cap, int32_t* ctrl,
                                 moyolo_stream_t stream) {
  MOYOLO_REQUIRE(row_offsets && ids && boxes && scores && labels && n_active && active_index && seq_ids &&
                     frame_rows && table && ctrl,
                 MOYOLO_ERR_BAD_ARG, "frame_emit: null pointer");
  MOYOLO_REQUIRE(n_seq > 0 && rows_pad > 0 && table_cap >= 0 && table_cap < (1ll << 31), MOYOLO_ERR_BAD_SHAPE,
                 "frame_emit: bad sizes");
  MOYOLO_REQUIRE(aligned16(boxes) && aligned16(frame_rows), MOYOLO_ERR_ALIGNMENT,
                 "frame_emit: boxes / frame_rows must be 16-byte aligned");
  launch_k(frame_emit_kernel, dim3(1), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      n_seq, static_cast<int>(rows_pad), row_offsets, ids, boxes, scores, labels, n_active, active_index, seq_ids,
      frame_rows, table, static_cast<int>(table_cap), ctrl);
  return check_launch("frame_emit_kernel");
}

extern "C" int moyolo_table_pack(const float* rows, int64_t n_rows, const int32_t* n_rows_dev, const int32_t* overflow_dev,
                                 int64_t capacity, float* send, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(send && (rows || (n_rows == 0 && !n_rows_dev)) && n_rows >= 0 && capacity > 0, MOYOLO_ERR_BAD_ARG,
                 "table_pack: bad arguments");
  const int64_t n = n_rows_dev ? capacity : (n_rows < capacity ? n_rows : capacity);
  const unsigned blocks = static_cast<unsigned>(((n + 1) * 9 + 255) / 256 > 592 ? 592 : ((n + 1) * 9 + 255) / 256);
  launch_k(table_pack_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), rows, n_rows, n_rows_dev,
           overflow_dev, capacity, send);
  return check_launch("table_pack_kernel");
}

extern "C" int moyolo_table_merge(const float* recv, int world, int64_t capacity, float* out, int32_t* info,
                                  moyolo_stream_t stream) {
  MOYOLO_REQUIRE(recv && out && info && world > 0 && capacity > 0, MOYOLO_ERR_BAD_ARG, "table_merge: bad arguments");
  const unsigned bx = static_cast<unsigned>((capacity * 9 + 255) / 256 > 64 ? 64 : (capacity * 9 + 255) / 256);
  launch_k(table_merge_kernel, dim3(bx, static_cast<unsigned>(world)), dim3(256), 0, static_cast<cudaStream_t>(stream), recv,
           world, capacity, out, info);
  return check_launch("table_merge_kernel");
}
