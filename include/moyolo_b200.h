/*
 * moyolo_b200.h — C ABI of libmoyolo_b200.so: the B200 (sm_100a) decoder hot path of
 * DecoderTracker / MO-YOLO.
 *
 * Every entry point takes raw DEVICE pointers (unless the parameter name ends in `_host`),
 * plain sizes and a CUDA stream handle. The library never allocates or frees device memory,
 * never synchronises the device, and enqueues all work on the caller's stream, so every call
 * is CUDA-graph capturable. Functions return MOYOLO_OK (0) or a non-zero status;
 * `moyolo_last_error()` then returns a thread-local, human-readable message.
 *
 * Reference interfaces replaced (paths relative to the reference repository root):
 *   - moyolo_msda_sampled_forward  <-  pybind `ms_deform_attn_forward`
 *         MOTR/models/ops/src/vision.cpp:13-16, MOTR/models/ops/src/ms_deform_attn.h:20-40,
 *         MOTR/models/ops/src/cuda/ms_deform_attn_cuda.cu:20-80 (the legacy FFI), and the
 *         Python core `multi_scale_deformable_attn_pytorch`, ultralytics/nn/modules/utils.py:41-78
 *   - moyolo_msda_sampled_backward <-  pybind `ms_deform_attn_backward`, MOTR/models/ops/src/vision.cpp:15,
 *         MOTR/models/ops/src/cuda/ms_deform_attn_cuda.cu:83-153
 *   - moyolo_msda_fused_forward    <-  ultralytics/nn/modules/transformer.py:268-285
 *         (softmax over L*P, sampling-location arithmetic, grid_sample gather, weighted sum)
 *   - moyolo_linear                <-  nn.Linear call sites transformer.py:264,268,269,286
 *         (value_proj / sampling_offsets / attention_weights / output_proj) and :576-580 (FFN)
 *   - moyolo_self_attention        <-  nn.MultiheadAttention call, transformer.py:637-641
 *   - moyolo_add_layernorm         <-  residual + nn.LayerNorm, transformer.py:640-641,646-647,578-579
 *   - moyolo_box_refine            <-  transformer.py:709 + ultralytics/nn/modules/utils.py:34-38
 *   - moyolo_score_head            <-  transformer.py:717-721, head.py:310
 *   - moyolo_pos2posemb            <-  transformer.py:183-190
 *   - moyolo_track_assign          <-  RuntimeTrackerBase.update, ultralytics/nn/modules/head.py:1201-1283
 *   - moyolo_track_compact         <-  QueryInteractionModule._select_active_tracks, MOTR/models/qim.py:184-187
 *                                      + Instances.__getitem__, MOTR/models/structures/instances.py:152-178
 */
#ifndef MOYOLO_B200_H_
#define MOYOLO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define MOYOLO_VERSION 100 /* 0.1.0 */
#define MOYOLO_MAX_LEVELS 8

/* status codes */
#define MOYOLO_OK 0
#define MOYOLO_ERR_BAD_ARG 1      /* null pointer, negative size, bad enum             -> ValueError   */
#define MOYOLO_ERR_BAD_SHAPE 2    /* sum(H_l*W_l) != Lv, ref last dim not 2|4, ...     -> ValueError   */
#define MOYOLO_ERR_UNSUPPORTED 3  /* dtype / head-dim combination with no kernel       -> RuntimeError */
#define MOYOLO_ERR_CUDA 4         /* launch or driver error                            -> RuntimeError */
#define MOYOLO_ERR_ALIGNMENT 5    /* pointer / stride not 16-byte aligned where needed -> ValueError   */

/* element types */
#define MOYOLO_F32 0
#define MOYOLO_BF16 1
#define MOYOLO_F64 2 /* generic gather only (the legacy FFI is dispatched on float/double) */

/* attention-weight normalisation of the fused gather */
#define MOYOLO_SOFTMAX 0      /* F.softmax over the joint L*P axis, transformer.py:271            */
#define MOYOLO_SOFTMAX_PLUS1 1 /* exp/(1+sum exp), MOTRMSDeformAttn `my_softmax`, transformer.py:239-244 */

/* epilogue flags of moyolo_linear */
#define MOYOLO_EPI_NONE 0
#define MOYOLO_EPI_RELU 1

/* GEMM engine selector of moyolo_linear */
#define MOYOLO_GEMM_AUTO 0    /* bf16 -> tcgen05, f32 -> SIMT */
#define MOYOLO_GEMM_SIMT 1    /* fp32-accumulate CUDA-core kernel (exact-parity path)              */
#define MOYOLO_GEMM_TCGEN05 2 /* TMA + tcgen05.mma + TMEM kernel (bf16 operands only)              */

typedef void* moyolo_stream_t; /* a cudaStream_t */

int moyolo_version(void);
const char* moyolo_last_error(void);
/* 1 if the current device is compute capability 10.x (the only supported target), else 0. */
int moyolo_device_supported(void);
/* Number of kernels this library has launched (or recorded into a graph being captured) from the CALLING host
 * thread since it first called the library. Monotonic, thread-local, never reset: callers take differences
 * (bench.py's `gpu_launches`, the per-graph launch count of moyolo_b200.TrackEngine). */
uint64_t moyolo_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Multi-scale deformable attention gather.
 *
 * value        [B, Lv, n_heads, head_dim]; consecutive spatial positions are `value_pos_stride`
 *              elements apart, consecutive batches `value_batch_stride` elements apart (so one
 *              layer's slice of an all-layers value tensor [B, Lv, n_layers*C] can be addressed
 *              without a copy). Head h of a position starts at element h*head_dim.
 * shapes_hw_host  host int32 [n_levels*2] = (H_0, W_0, H_1, W_1, ...); sum(H_l*W_l) must equal Lv.
 * rows         total query rows R. If `row_offsets` (device int32 [B+1]) is NULL the rows are
 *              dense: batch b owns rows [b*R/B, (b+1)*R/B). Otherwise batch b owns rows
 *              [row_offsets[b], row_offsets[b+1]) (ragged lock-step sequences).
 * out          [R, n_heads*head_dim], row stride `out_row_stride` elements, same dtype as value.
 * -------------------------------------------------------------------------------------------*/

/* Pre-normalised mode (legacy FFI and `multi_scale_deformable_attn_pytorch`):
 *   loc     [R, n_heads, n_levels, n_points, 2] normalised (x, y) in [0,1] (may fall outside),
 *   weights [R, n_heads, n_levels, n_points]; both of dtype `aux_dtype` (F32, or F64 with F64 value). */
int moyolo_msda_sampled_forward(const void* value, int value_dtype, int64_t value_batch_stride,
                                int64_t value_pos_stride, const int32_t* shapes_hw_host,
                                int n_levels, int batch, int64_t len_v, int n_heads, int head_dim,
                                int n_points, const void* loc, const void* weights, int aux_dtype,
                                int64_t rows, const int32_t* row_offsets, void* out,
                                int64_t out_row_stride, moyolo_stream_t stream);

/* moyolo_msda_fused_forward on a HEAD-MAJOR value tensor [B, H, Lv, head_dim] (strides in elements: batch, head,
 * position): the x0 / x1 bilinear corners of a sampling point are one contiguous 2 * head_dim segment. Experimental
 * (bf16, 8 heads x 32, 3 levels x 4 points); same arithmetic and results as the channel-last entry point. */
int moyolo_msda_fused_forward_headmajor(const void* value, int value_dtype, int64_t value_batch_stride,
                                        int64_t value_head_stride, int64_t value_pos_stride,
                                        const int32_t* shapes_hw_host, int n_levels, int batch, int64_t len_v,
                                        int n_heads, int head_dim, int n_points, const float* offsets,
                                        int64_t offsets_row_stride, const float* logits, int64_t logits_row_stride,
                                        const float* refer, int ref_levels, int ref_dim, int softmax_mode, int64_t rows,
                                        const int32_t* row_offsets, void* out, int64_t out_row_stride,
                                        moyolo_stream_t stream);

/* Fused mode with the sampling_offsets | attention_weights projection inside the gather kernel
 * (transformer.py:268-285 in ONE launch): offsets|logits = xq . w_offlog^T + b_offlog is computed per
 * (8 rows, head) CTA on mma.sync and never leaves shared memory.
 *   xq       [R, 256] bf16 (row stride `xq_row_stride` elements): the query operand (x + pos),
 *   w_offlog [n_heads*L*P*3, 256] bf16 = rows of sampling_offsets.weight followed by attention_weights.weight,
 *   b_offlog fp32, same stacking. Serves bf16 value, 8 heads x 32, 3 levels x 4 points (the decoder's
 *   configuration); returns MOYOLO_ERR_UNSUPPORTED otherwise. */
int moyolo_msda_proj_fused_forward(const void* value, int value_dtype, int64_t value_batch_stride,
                                   int64_t value_pos_stride, const int32_t* shapes_hw_host, int n_levels,
                                   int batch, int64_t len_v, int n_heads, int head_dim, int n_points,
                                   const void* xq, int64_t xq_row_stride, const void* w_offlog,
                                   const float* b_offlog, const float* refer, int ref_levels, int ref_dim,
                                   int softmax_mode, int64_t rows, const int32_t* row_offsets, void* out,
                                   int64_t out_row_stride, moyolo_stream_t stream);

/* Backward of the pre-normalised mode: replaces pybind `ms_deform_attn_backward`
 * (MOTR/models/ops/src/vision.cpp:15, ms_deform_attn.h:42-61, cuda/ms_deform_attn_cuda.cu:83-153, kernels
 * cuda/ms_deform_im2col_cuda.cuh:88-159, 301-920) and the autograd of multi_scale_deformable_attn_pytorch.
 *   grad_out     [R, n_heads*head_dim] (row stride `grad_out_row_stride` elements),
 *   grad_value   [B, Lv, n_heads, head_dim] contiguous, ACCUMULATED into (the caller zero-fills it),
 *   grad_loc     [R, n_heads, n_levels, n_points, 2], grad_weights [R, n_heads, n_levels, n_points] (overwritten).
 * Gradients, loc and weights are fp32 for fp32 / bf16 `value` and fp64 for fp64 `value` (`aux_dtype`). */
int moyolo_msda_sampled_backward(const void* value, int value_dtype, int64_t value_batch_stride,
                                 int64_t value_pos_stride, const int32_t* shapes_hw_host, int n_levels,
                                 int batch, int64_t len_v, int n_heads, int head_dim, int n_points,
                                 const void* loc, const void* weights, int aux_dtype, const void* grad_out,
                                 int64_t grad_out_row_stride, int64_t rows, const int32_t* row_offsets,
                                 void* grad_value, void* grad_loc, void* grad_weights, moyolo_stream_t stream);

/* Fused mode (transformer.py:268-285): raw Linear outputs in, no intermediate tensors.
 *   offsets [R, n_heads, n_levels, n_points, 2] fp32, row stride `offsets_row_stride` elements,
 *   logits  [R, n_heads, n_levels*n_points]     fp32, row stride `logits_row_stride` elements
 *           (both may be column slices of one fused GEMM output),
 *   refer   [R, ref_levels, ref_dim] fp32 with ref_levels in {1, n_levels}, ref_dim in {2, 4}:
 *           ref_dim 4: loc = ref_xy + off / n_points * ref_wh * 0.5      (transformer.py:280-282)
 *           ref_dim 2: loc = ref_xy + off / (W_l, H_l)                   (transformer.py:276-279) */
int moyolo_msda_fused_forward(const void* value, int value_dtype, int64_t value_batch_stride,
                              int64_t value_pos_stride, const int32_t* shapes_hw_host, int n_levels,
                              int batch, int64_t len_v, int n_heads, int head_dim, int n_points,
                              const float* offsets, int64_t offsets_row_stride, const float* logits,
                              int64_t logits_row_stride, const float* refer, int ref_levels,
                              int ref_dim, int softmax_mode, int64_t rows,
                              const int32_t* row_offsets, void* out, int64_t out_row_stride,
                              moyolo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * y[M, N] = act(x[M, K] . w[N, K]^T + bias[N]); x row stride ldx, y row stride ldy (elements).
 * in_dtype applies to x and w (F32 or BF16); bias is fp32 (may be NULL); out_dtype F32 or BF16.
 * Optional row mask: if `zero_rows` (device uint8 [M]) is non-NULL, rows with a non-zero entry
 * are written as zeros (value_mask semantics of transformer.py:265-266).
 * The tcgen05 engine needs K % 64 == 0, N % 32 == 0, 16-byte aligned x, w and ldx*2 % 16 == 0.
 * -------------------------------------------------------------------------------------------*/
int moyolo_linear(const void* x, int64_t ldx, const void* w, const float* bias, void* y,
                  int64_t ldy, int64_t M, int N, int K, int in_dtype, int out_dtype, int epilogue,
                  const uint8_t* zero_rows, int engine, moyolo_stream_t stream);

/* The persistent, weight-resident tcgen05 kernel for TALL x (value_proj over all pyramid positions,
 * transformer.py:264) with an explicit CTA budget: y[M,N] = x[M,256] . w[N,256]^T + bias as bf16, rows with
 * zero_rows != 0 written as zeros. max_ctas > 0 bounds the grid (0 = one CTA per SM): a frame runs the layers'
 * value projections NEXT TO its latency-bound main chain, which needs some SMs left free. bf16, K == 256,
 * N % 128 == 0, 16-byte aligned operands and row strides. */
int moyolo_linear_tall(const void* x, int64_t ldx, const void* w, const float* bias, void* y, int64_t ldy,
                       int64_t M, int N, int K, const uint8_t* zero_rows, int max_ctas, moyolo_stream_t stream);

/* Two A operands in ONE launch: y[:, :n_split] = x1 . w[:n_split]^T, y[:, n_split:] = x2 . w[n_split:]^T (+ bias).
 * The packed in-projection of nn.MultiheadAttention with q = k = x + pos, v = x (transformer.py:637-638):
 * x1 = x + pos, x2 = x, n_split = 2C. bf16 operands, tcgen05 engine only; n_split % 64 == 0. */
int moyolo_linear_dual(const void* x1, int64_t ldx1, const void* x2, int64_t ldx2, int n_split,
                       const void* w, const float* bias, void* y, int64_t ldy, int64_t M, int N, int K,
                       int out_dtype, moyolo_stream_t stream);

/* Post-norm residual block in one launch (transformer.py:640-641, 646-647, 578-579; qim.py:277-298):
 *   t = x[M,K] . w[N,K]^T + bias;  out = LayerNorm(t + residual) * gamma + beta   (N == 256, eps as given)
 * written as out_f32 [M,N] fp32, out_lp [M,N] bf16 and out_pos_lp [M,N] bf16 = out + pos (any may be NULL).
 * x, w bf16 (tcgen05 engine); residual/pos/out_* contiguous [M, N]. The LayerNorm row is spread over the
 * 8 CTAs of a thread-block cluster; statistics are exchanged through distributed shared memory. */
int moyolo_linear_add_layernorm(const void* x, int64_t ldx, const void* w, const float* bias,
                                const float* residual, const float* gamma, const float* beta, float eps,
                                int64_t M, int N, int K, float* out_f32, void* out_lp, const float* pos,
                                void* out_pos_lp, moyolo_stream_t stream);

/* moyolo_linear_add_layernorm with the decoder's class-score head fused behind the LayerNorm (transformer.py:717-721,
 * head.py:310): logits[M,nc] = bf16(out) . score_w[nc,256]^T + score_b, scores[M] = sigmoid(max_c logits),
 * labels[M] = argmax_c (first maximum). Each CTA of the cluster sends the partial dot products of its 32 columns to
 * cluster rank 0 (st.async + mbarrier), which finishes the row. 1 <= nc <= 8; logits / scores / labels may be NULL. */
int moyolo_linear_add_layernorm_scores(const void* x, int64_t ldx, const void* w, const float* bias,
                                       const float* residual, const float* gamma, const float* beta, float eps,
                                       int64_t M, int N, int K, float* out_f32, void* out_lp, const float* score_w,
                                       const float* score_b, int nc, float* logits, float* scores, int32_t* labels,
                                       moyolo_stream_t stream);

/* The whole post-norm FFN block in one launch (transformer.py:576-580; QIM qim.py:280-282, 290-298):
 *   out = LayerNorm(residual + relu(x[M,C] . w1[F,C]^T + b1) . w2[C,F]^T + b2) * gamma + beta,  C == 256, F in {256, 1024}
 * outputs as moyolo_linear_add_layernorm. h: bf16 [M, F] scratch (row stride F) that carries the hidden activations
 * between the two GEMM phases of the kernel (one thread-block cluster of 8 CTAs per 128 rows; each CTA computes F/8
 * hidden columns, a cluster barrier publishes them, then each CTA contracts the full hidden row tile for its 32
 * output columns and the cluster normalises the row). Bit-identical to moyolo_linear(relu) + moyolo_linear_add_layernorm. */
int moyolo_ffn_add_layernorm(const void* x, int64_t ldx, const void* w1, const float* b1, const void* w2,
                             const float* b2, void* h, int F, const float* residual, const float* gamma,
                             const float* beta, float eps, int64_t M, int C, float* out_f32, void* out_lp,
                             const float* pos, void* out_pos_lp, moyolo_stream_t stream);

/* Query self-attention over ragged sequences (transformer.py:637-641 with nn.MultiheadAttention
 * semantics: scores = (q/sqrt(head_dim)) . k^T, softmax over all keys of the same sequence, . v).
 * q, k, v: [R, n_heads*head_dim] slices with row strides ldq/ldk/ldv (elements) of dtype `dtype`;
 * attn_mask: optional additive fp32 [Rq, Rk] mask for the dense single-batch case (NULL in eval).
 * out: [R, n_heads*head_dim], row stride ldo, dtype `dtype`. row_offsets as above (host-side
 * `row_offsets_host` [B+1] mirrors it for grid sizing; it may over-estimate the device values, extra
 * CTAs exit). If `seg_len` (device int32 [B]) is non-NULL sequence b attends over rows
 * [row_offsets[b], row_offsets[b] + seg_len[b]) only. head_dim must be 32 or 64. */
int moyolo_self_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                          int64_t ldv, void* out, int64_t ldo, int dtype, int batch,
                          const int32_t* row_offsets, const int32_t* row_offsets_host,
                          const int32_t* seg_len, int n_heads, int head_dim, const float* attn_mask,
                          moyolo_stream_t stream);

/* out = LayerNorm(x + residual) * gamma + beta over the last dim C (eps as given), fp32 math.
 * x: [R, C] fp32 (GEMM output), residual: [R, C] fp32 (may be NULL).
 * Writes up to three results (any may be NULL): out_f32 [R, C] fp32; out_lp [R, C] of `lp_dtype`;
 * out_pos_lp [R, C] of `lp_dtype` = (LayerNorm result + pos) with pos [R, C] fp32
 * (the `with_pos_embed` operand of the next GEMM, transformer.py:637,644). */
int moyolo_add_layernorm(const float* x, const float* residual, const float* gamma,
                         const float* beta, float eps, int64_t rows, int C, float* out_f32,
                         void* out_lp, const float* pos, void* out_pos_lp, int lp_dtype,
                         moyolo_stream_t stream);

/* out_lp = (a [+ b]) converted to lp_dtype; a, b: [R, C] fp32 (b may be NULL). */
int moyolo_add_cast(const float* a, const float* b, void* out_lp, int lp_dtype, int64_t n_elems,
                    moyolo_stream_t stream);

/* new_ref[R,4] = sigmoid(h[R,K] . w3[4,K]^T + b3 + inverse_sigmoid(ref[R,4]))  (transformer.py:709;
 * inverse_sigmoid = log(clamp(x,1e-5)/clamp(1-x,1e-5)) after clamp to [0,1], utils.py:34-38).
 * h is the output of the first two MLP layers, dtype h_dtype. w3/b3/ref/new_ref are fp32. */
int moyolo_box_refine(const void* h, int64_t ldh, int h_dtype, const float* w3, const float* b3,
                      const float* ref, float* new_ref, int64_t rows, int K, moyolo_stream_t stream);

/* logits[R,nc] = x[R,K] . w[nc,K]^T + b; scores[R] = max_c sigmoid(logits) (head.py:310),
 * labels[R] = argmax_c logits (first maximum, torch.max semantics), max_logit[R] = max_c logits
 * (the top-k key of head.py:1048). logits/scores/labels/max_logit may be NULL. */
int moyolo_score_head(const void* x, int64_t ldx, int x_dtype, const float* w, const float* b,
                      float* logits, float* scores, int32_t* labels, int64_t rows, int K, int nc,
                      float* max_logit, moyolo_stream_t stream);

/* refer_sig[R,4] = sigmoid(refer_logit[R,4]) (transformer.py:690) */
int moyolo_sigmoid(const float* x, float* y, int64_t n, moyolo_stream_t stream);
/* y = inverse_sigmoid(x), eps 1e-5 (MOTR/util/misc.py:532-536, qim.py:299) */
int moyolo_inverse_sigmoid(const float* x, float* y, int64_t n, moyolo_stream_t stream);

/* pos2posemb (transformer.py:183-190): pos [R, n_coord] fp32 -> emb [R, n_coord*num_pos_feats] fp32,
 * emb[r, c*F + i] = (i even ? sin : cos)(pos[r,c]*2*pi / temperature^(2*(i/2)/F)). */
int moyolo_pos2posemb(const float* pos, float* emb, int64_t rows, int n_coord, int num_pos_feats,
                      float temperature, moyolo_stream_t stream);

/* pos-MLP first layer (DeformableTransformerDecoder, transformer.py:491 with MLP(4, 2*hd, hd, 2)):
 * y[R,N] = relu(x[R,4] . w[N,4]^T + b) written as `out_dtype`. */
int moyolo_linear_k4_relu(const float* x, const float* w, const float* b, void* y, int out_dtype,
                          int64_t rows, int N, moyolo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Track-query update (one sequence per call; all state on device, no host sync).
 *
 * moyolo_track_assign: RuntimeTrackerBase.update (head.py:1201-1283) restated as parallel scans.
 *   scores [N] fp32, boxes [N,4] fp32 (decoder boxes), obj_idxes [N] int64 in/out,
 *   disappear_time [N] int64 in/out, counters int64 [2] in/out = {max_obj_id, max_obj_id_pre},
 *   workspace: moyolo_track_workspace_bytes(N) bytes.
 *   In query order: id==-1 && score>=score_thresh -> id = max_obj_id++;
 *   id>=0 && score<filter_thresh -> disappear_time++, and at >= miss_tolerance id = -1.
 *   Then the greedy duplicate filter (IoU > iou_thresh, head.py:1155-1196) and the renumbering
 *   (head.py:1268-1282) are evaluated on the active subset only for their side effect on counters.
 * moyolo_track_compact: select rows with obj_idxes >= 0 preserving order; writes n_active (int32
 *   device scalar) and, for each of the `n_fields` row-major arrays (src[i], dst[i], row_bytes[i]
 *   given as HOST arrays of device pointers / sizes), dst[j] = src[index of j-th active row].
 * -------------------------------------------------------------------------------------------*/
int64_t moyolo_track_workspace_bytes(int64_t n);
int moyolo_track_assign(const float* scores, const float* boxes, int64_t* obj_idxes,
                        int64_t* disappear_time, int64_t* counters, int64_t n, float score_thresh,
                        float filter_thresh, int miss_tolerance, float iou_thresh, void* workspace,
                        moyolo_stream_t stream);
/* Batched form: one CTA per sequence over rows [row_offsets[s], row_offsets[s+1]) of the frame arrays;
 * counters int64 [n_seq, 2]; workspace n_seq * moyolo_track_workspace_bytes(max_rows_per_seq) bytes.
 * ctrl (may be NULL): frame control block, see moyolo_frame_assemble; a set abort flag makes the call a no-op. */
int moyolo_track_assign_batched(const float* scores, const float* boxes, int64_t* obj_idxes,
                                int64_t* disappear_time, int64_t* counters, const int32_t* row_offsets,
                                int n_seq, int64_t max_rows_per_seq, float score_thresh, float filter_thresh,
                                int miss_tolerance, float iou_thresh, void* workspace, const int32_t* ctrl,
                                moyolo_stream_t stream);
/* Second half of the batched update only: the duplicate filter + renumbering (head.py:1155-1196, 1268-1282) on ids
 * that were already updated (moyolo_frame_assign_compact). It changes nothing but `counters`, so a frame runs it
 * off its critical path. */
int moyolo_track_suppress_batched(const float* boxes, int64_t* obj_idxes, int64_t* counters,
                                  const int32_t* row_offsets, int n_seq, int64_t max_rows_per_seq,
                                  float iou_thresh, void* workspace, const int32_t* ctrl, moyolo_stream_t stream);
int moyolo_track_compact(const int64_t* obj_idxes, int64_t n, int32_t* n_active,
                         int32_t* active_index, const void* const* src_host, void* const* dst_host,
                         const int64_t* row_bytes_host, int n_fields, moyolo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Device-resident frame state for lock-step sequences (fixed capacity `cap` tracks per sequence; all
 * counts in device memory so a frame replays as one CUDA graph for a padded row count `rows_pad`).
 *
 * moyolo_frame_assemble: builds the frame's query rows, tracks first then detect queries per sequence
 *   (head.py:1056-1064,1108-1109): x [rows_pad,C] (class_embed[t_label] | det_embed), refer_logit
 *   [rows_pad,4], pos [rows_pad,C] (t_qpos | pos2posemb(det_refer)), ids/dis [rows_pad] (prev | -1/0),
 *   row_offsets [n_seq+1]. State arrays are [n_seq, cap, ...]; n_tracks int32 [n_seq]. Optional fused outputs
 *   (NULL = skip): refer_sig [rows_pad,4] = sigmoid(refer_logit) (transformer.py:690) and the first layer's GEMM
 *   operands x_lp = x, xq_lp = x + pos as `lp_dtype`.
 * moyolo_frame_compact: per sequence, rows with ids >= 0 in order: n_active [n_seq], active_index,
 *   QIM inputs gathered to frame-layout compact buffers c_* (sequence s at row_offsets[s]) and
 *   t_label/t_ids/t_dis written straight to the state arrays. Optional (NULL = skip) QIM operands as `lp_dtype`:
 *   q_qk_lp = c_hs + pos2posemb(c_ref) (qim.py:255,271), q_tgt_lp = c_hs.
 * moyolo_frame_assign_compact: the ID assignment of RuntimeTrackerBase.update (head.py:1232-1243: in query order
 *   id==-1 && score>=score_thresh -> id = max_obj_id++ with max_obj_id = counters[2*s]; id>=0 &&
 *   score<filter_thresh -> disappear_time++, id = -1 at >= miss_tolerance) fused with moyolo_frame_compact on the
 *   updated ids, one launch. ids_in/dis_in (from frame_assemble) are read-only; the updated values go to
 *   ids_out/dis_out [rows_pad] (must not alias; padding rows get -1 / 0). counters is NOT modified: follow with
 *   moyolo_track_suppress_batched (any time before the next frame) for the reference's counter side effects.
 * moyolo_frame_writeback: t_qpos <- new_qpos rows, t_ref <- inverse_sigmoid(c_box) (or, with c_box == NULL, of
 *   boxes[row_offsets[s] + active_index[.]]: the frame then passes boxes == NULL to moyolo_frame_assign_compact, which
 *   no longer has to wait for the last layer's box head), n_tracks <- n_active
 *   (MOTR/models/qim.py:298-300); optional info int32 [n_seq+8] = (n_active | ctrl), the frame summary a host reads.
 * moyolo_frame_emit: the frame's results. frame_rows [rows_pad, 8] fp32 = (id, cx, cy, w, h, score, label,
 *   seq) for every row (padding rows id = -1) -- what a host reads back with one copy -- and the tracked
 *   objects (ids >= 0) appended to the device-resident track table [table_cap, 9] fp32 =
 *   (seq, frame, id, cx, cy, w, h, score, cls), the row format of the sharding gather (the MOTChallenge /
 *   TrackResults.save_txt fields, ultralytics/engine/results.py:475-512). seq_ids int32 [n_seq] maps the
 *   lock-step slot to the global sequence index.
 *
 * ctrl: device int32 [8] control block = {abort, frame counter, table cursor, table overflow, rows wanted
 *   by the aborting frame, track overflow, ...}. ctrl[5] (track overflow) is set, sticky, by frame_compact /
 *   frame_assign_compact when a sequence has more active tracks than the carried state holds (`cap`): the surplus
 *   tracks are dropped from the state (the reference's Instances has no such limit), so the host must treat the
 *   flag as an error (moyolo_b200.TrackEngine raises) and re-run with a larger capacity. The host may launch a frame speculatively with a rows_pad derived from an
 *   OLDER frame's track counts: if the real row count does not fit, frame_assemble sets the sticky abort
 *   flag, builds an in-bounds detect-only frame, and every state-writing call (track_assign_batched,
 *   frame_compact, frame_writeback, frame_emit) of this and later frames is a no-op until the host clears
 *   the flag and re-launches. ctrl may be NULL for assemble/compact/writeback (no speculation).
 * -------------------------------------------------------------------------------------------*/
int moyolo_frame_assemble(int n_seq, int n_detect, int C, int cap, const int32_t* n_tracks,
                          const float* t_ref, const float* t_qpos, const int32_t* t_label,
                          const int64_t* t_ids, const int64_t* t_dis, const float* class_embed,
                          const float* det_embed, const float* det_refer, float* x, float* refer_logit,
                          float* pos, int64_t* ids, int64_t* dis, int32_t* row_offsets, int64_t rows_pad,
                          int num_pos_feats, float temperature, int32_t* ctrl, float* refer_sig, void* x_lp,
                          void* xq_lp, int lp_dtype, moyolo_stream_t stream);
int moyolo_frame_compact(int n_seq, int C, int cap, const int32_t* row_offsets, const int64_t* ids,
                         const int64_t* dis, const int32_t* labels, const float* refer_logit,
                         const float* pos, const float* hs, const float* boxes, int32_t* n_active,
                         int32_t* active_index, float* c_ref, float* c_pos, float* c_hs, float* c_box,
                         int32_t* t_label, int64_t* t_ids, int64_t* t_dis, int32_t* ctrl,
                         void* q_qk_lp, void* q_tgt_lp, int lp_dtype, int num_pos_feats, float temperature,
                         moyolo_stream_t stream);
int moyolo_frame_assign_compact(int n_seq, int C, int cap, int64_t rows_pad, const int32_t* row_offsets,
                                const float* scores, const int64_t* ids_in, const int64_t* dis_in,
                                const int64_t* counters, float score_thresh, float filter_thresh,
                                int miss_tolerance, int64_t* ids_out, int64_t* dis_out, const int32_t* labels,
                                const float* refer_logit, const float* pos, const float* hs, const float* boxes,
                                int32_t* n_active, int32_t* active_index, float* c_ref, float* c_pos, float* c_hs,
                                float* c_box, int32_t* t_label, int64_t* t_ids, int64_t* t_dis, int32_t* ctrl,
                                void* q_qk_lp, void* q_tgt_lp, int lp_dtype, int num_pos_feats, float temperature,
                                moyolo_stream_t stream);
int moyolo_frame_writeback(int n_seq, int C, int cap, const int32_t* row_offsets, const int32_t* n_active,
                           const float* new_qpos, const float* c_box, float* t_qpos, float* t_ref,
                           int32_t* n_tracks, const int32_t* ctrl, int32_t* info, const float* boxes,
                           const int32_t* active_index, moyolo_stream_t stream);
int moyolo_frame_emit(int n_seq, int64_t rows_pad, const int32_t* row_offsets, const int64_t* ids,
                      const float* boxes, const float* scores, const int32_t* labels, const int32_t* n_active,
                      const int32_t* active_index, const int32_t* seq_ids, float* frame_rows, float* table,
                      int64_t table_cap, int32_t* ctrl, moyolo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Encoder-side query selection of MYDecoder (ultralytics/nn/modules/head.py:993-1113), SURVEY.md 8(f1).
 *
 * moyolo_enc_output_scores: f = LayerNorm((zero_in_rows[r] ? 0 : x[r]) . w^T + bias) * gamma + beta (head.py:1039,
 *   `enc_output` on valid_mask * feats), logits = f . score_w^T + score_b (:1041), max_logit = max_c logits (:1048).
 *   x [M,256] bf16 (row stride ldx), w [256,256] bf16; out_f32 [M,256] fp32, out_lp [M,256] bf16, logits [M,nc],
 *   max_logit [M] (any output may be NULL); nc <= 8. One tcgen05 kernel, 128 rows per CTA.
 * moyolo_topk: per row of scores [batch, n] (row stride given) the indices of the k largest values, ordered as
 *   torch.topk(sorted=True): descending value, equal values by ascending index. k <= 1024.
 * moyolo_select_gather: embed[b,j] = features[b, idx[b,j]] (fp32 [batch,k,C] and, if embed_lp != NULL, the same as
 *   `lp_dtype`), enc_scores[b,j] = logits[b, idx[b,j]] (head.py:1092, 1104). features [batch, len_v, C] fp32.
 * moyolo_anchor_box: refer[r] = h[r] . w3^T + b3 + anchor(idx[r]) (head.py:1044, 1053): h = the two hidden layers of
 *   enc_bbox_head on the selected rows, anchors per head.py:993-1010 (cell centres, wh = grid_size * 2^level, logit
 *   space, +inf where a coordinate leaves (eps, 1-eps); the reference divides (x, y) by (H, W) -- reproduced).
 * moyolo_anchor_invalid: invalid[p] = 1 where the anchor of pyramid position p is masked (valid_mask == False).
 * moyolo_mask_rows: out[r] = zero_rows[r % period] ? 0 : x[r] (fp32 path of valid_mask * feats).
 * -------------------------------------------------------------------------------------------*/
int moyolo_enc_output_scores(const void* x, int64_t ldx, const void* w, const float* bias, const float* gamma,
                             const float* beta, float eps, const uint8_t* zero_in_rows, const float* score_w,
                             const float* score_b, int nc, int64_t M, float* out_f32, void* out_lp, float* logits,
                             float* max_logit, moyolo_stream_t stream);
int moyolo_topk(const float* scores, int64_t row_stride, int n, int batch, int k, int32_t* idx_out, float* val_out,
                moyolo_stream_t stream);
int moyolo_select_gather(const float* features, const float* logits, const int32_t* idx, int batch, int k,
                         int64_t len_v, int C, int nc, float* embed, void* embed_lp, int lp_dtype, float* enc_scores,
                         moyolo_stream_t stream);
int moyolo_anchor_box(const void* h, int64_t ldh, int h_dtype, const float* w3, const float* b3, const int32_t* idx,
                      const int32_t* shapes_hw_host, int n_levels, int64_t len_v, float grid_size, float eps,
                      float* refer, int64_t rows, int K, moyolo_stream_t stream);
int moyolo_anchor_invalid(const int32_t* shapes_hw_host, int n_levels, int64_t len_v, float grid_size, float eps,
                          uint8_t* invalid, moyolo_stream_t stream);
int moyolo_mask_rows(const float* x, const uint8_t* zero_rows, int64_t period, float* out, int64_t rows, int C,
                     moyolo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Host-side frame submission: the stream operations around one captured frame graph as ONE call
 * (the frame itself takes ~0.26 ms on the device; issuing these one by one from an interpreter costs more).
 *   on copy_stream: [wait ev_slot_free] [wait for main_stream if sync_inputs] n_inputs x memcpyAsync
 *                   (host-pinned or device sources, UVA) -> record ev_copy
 *   on main_stream: wait ev_copy -> launch graph_exec -> n_outputs x memcpyAsync -> record ev_done
 *   With out_stream_valid == 1 the result copies leave the main stream: record ev_graph on main_stream, and on
 *   out_stream: wait ev_graph -> n_outputs x memcpyAsync -> record ev_done -- the next frame's graph then starts
 *   right behind this one (the caller double-buffers the copied-from buffers per frame parity).
 * All handles are plain CUDA runtime handles (cudaStream_t, cudaEvent_t, cudaGraphExec_t) owned by the
 * caller. With n_inputs == 0 only the main-stream part runs.
 * -------------------------------------------------------------------------------------------*/
#define MOYOLO_SUBMIT_MAX_COPIES 4
typedef struct {
  void* copy_stream;
  void* main_stream;
  int main_stream_valid; /* must be 1 (guards against a zeroed descriptor; stream 0 is a valid handle) */
  int sync_inputs;
  void* ev_slot_free;
  void* ev_scratch;
  void* ev_copy;
  void* ev_done;
  void* graph_exec;
  int n_inputs;
  int n_outputs;
  const void* in_src[MOYOLO_SUBMIT_MAX_COPIES];
  void* in_dst[MOYOLO_SUBMIT_MAX_COPIES];
  int64_t in_bytes[MOYOLO_SUBMIT_MAX_COPIES];
  const void* out_src[MOYOLO_SUBMIT_MAX_COPIES];
  void* out_dst[MOYOLO_SUBMIT_MAX_COPIES];
  int64_t out_bytes[MOYOLO_SUBMIT_MAX_COPIES];
  void* out_stream;      /* optional stream for the result copies */
  int out_stream_valid;  /* 1 = use out_stream / ev_graph */
  void* ev_graph;
  /* optional value projection AHEAD of the frame graph (vp_valid == 1): on vp_stream, wait ev_copy and
   * ev_tail_prev (recorded by the previous frame's graph once its last decoder layer is issued; NULL for the
   * first frame), run moyolo_linear_tall(vp_x -> vp_y) on at most vp_max_ctas SMs, record ev_vp; main_stream
   * waits ev_vp before the graph launch. The projection then fills the SMs the previous frame's tail leaves idle. */
  void* vp_stream;
  int vp_valid;
  void* ev_tail_prev;
  void* ev_vp;
  const void* vp_x;
  int64_t vp_ldx;
  const void* vp_w;
  const float* vp_bias;
  void* vp_y;
  int64_t vp_ldy;
  int64_t vp_M;
  int vp_N;
  int vp_max_ctas;
  /* vp_valid == 2: instead of the value projection alone, launch the captured graph `pre_graph_exec` on vp_stream
   * (everything of the frame that depends on its inputs only: input projection, value projection, query selection);
   * same event protocol (wait ev_copy [and ev_tail_prev], record ev_vp, main_stream waits ev_vp). */
  void* pre_graph_exec;
} moyolo_frame_submit_t;
int moyolo_frame_submit(const moyolo_frame_submit_t* d);

/* ---------------------------------------------------------------------------------------------
 * Whole decoder in ONE launch (moyolo_b200/csrc/decoder_cluster.cu). Replaces MOTRTransformerDecoder.forward
 * (ultralytics/nn/modules/transformer.py:676-728): for every layer the self-attention (nn.MultiheadAttention,
 * :637-641), MSDeformAttn (:246-287; value_proj excepted -- `values` holds all layers' projections, one GEMM ahead
 * of the frame), the FFN (:576-580), the three LayerNorms, the box refinement sigmoid(bbox_head[i](x) +
 * inverse_sigmoid(refer)) (:709) and, after the last layer, the class-score head (:717-721).
 * Ragged lock-step batch: rows [row_offsets[s], row_offsets[s+1]) belong to sequence s (device array, int32).
 * A cluster of 8 CTAs owns `rows_per_tile` (= 32) query rows through all layers; the clusters meet at one
 * grid-wide barrier per layer (self-attention keys), so ALL clusters must be co-resident: the call fails with
 * MOYOLO_ERR_UNSUPPORTED when ceil(rows_pad / rows_per_tile) + n_seq - 1 exceeds moyolo_decoder_cluster_limits'
 * max_clusters; sequences longer than kv_cap rows are reported through *status (device int, set to 1).
 * Measured on B200 this one-launch schedule is NOT faster than the launch-chained layers, so
 * moyolo_b200.TrackEngine uses it only on request.
 * bf16 weights [out, in] row-major as in the checkpoint; fp32 biases / LayerNorm parameters; built for d_model 256,
 * 8 heads, d_ffn 1024, 3 levels x 4 points, nc <= 8.
 * -------------------------------------------------------------------------------------------*/
typedef struct {
  const void *wqkv, *wo;        /* self_attn.in_proj_weight [768,256], self_attn.out_proj.weight [256,256]          */
  const void *woff;             /* [sampling_offsets.weight (192) ; attention_weights.weight (96)] x 256            */
  const void *wout;             /* cross_attn.output_proj.weight [256,256]                                          */
  const void *w1, *w2;          /* linear1.weight [1024,256], linear2.weight [256,1024]                             */
  const void *wb1, *wb2;        /* dec_bbox_head[i].layers.{0,1}.weight [256,256]                                   */
  const float *bqkv, *bo, *boff, *bout, *b1, *b2, *bb1, *bb2;
  const float *wb3, *bb3;       /* dec_bbox_head[i].layers.2 weight [4,256] / bias [4], fp32                        */
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *ln3_w, *ln3_b;
} moyolo_decoder_layer_weights_t;

typedef struct {
  int n_layers, d_model, n_heads, d_ffn, n_levels, n_points;
  moyolo_decoder_layer_weights_t layers[8];
  const float* x_in;            /* [rows_pad, 256] content embeddings (residual stream)                             */
  const float* pos;             /* [rows_pad, 256] track_query_embed / query_pos, the same in every layer (:705-707) */
  const float* refer0;          /* [rows_pad, 4] sigmoid(refer_bbox) (:690)                                         */
  float* x_out;                 /* [rows_pad, 256] output embedding of the last layer                               */
  void* x_lp_out;               /* optional bf16 copy of x_out                                                      */
  float* refer_out[8];          /* refer_out[i]: [rows_pad, 4] refined boxes after layer i (entries may be NULL)    */
  void* kv;                     /* scratch, bf16 [2, rows_pad, 512]                                                 */
  const void* values;           /* bf16 [B, Lv, n_layers*256]: layer i reads columns [256 i, 256 i + 256)           */
  int64_t value_batch_stride, value_pos_stride;   /* elements                                                      */
  int32_t value_shapes[2 * MOYOLO_MAX_LEVELS];    /* (H, W) per level, host                                        */
  int softmax_mode;
  const int32_t* row_offsets;   /* device int32 [n_seq + 1]                                                         */
  int n_seq;
  int64_t rows_pad;
  int rows_per_tile;            /* 32                                                                               */
  void* grid_barrier;           /* device uint32, must be 0 when the kernel starts                                  */
  int reset_barrier;            /* 1: the call enqueues a memset of grid_barrier first                              */
  int* status;                  /* optional device int                                                              */
  const float *score_w, *score_b;   /* dec_score_head[last]: fp32 [nc, 256], [nc]; NULL = no score head            */
  int nc;
  float* logits;                /* [rows_pad, nc]                                                                   */
  float* scores;                /* [rows_pad] sigmoid(max logit)                                                    */
  int32_t* labels;              /* [rows_pad] argmax                                                                */
  float eps;                    /* LayerNorm eps                                                                    */
  void* profile;                /* optional device int64 [n_layers, 16]: %globaltimer stamps of the stages of cluster
                                 * 0 / rank 0 (benchmarks/dc_stages.py); NULL in production                          */
} moyolo_decoder_cluster_t;
int moyolo_decoder_cluster_forward(const moyolo_decoder_cluster_t* a, moyolo_stream_t stream);
/* Co-resident clusters of the current device and the longest sequence (rows) the key/value staging holds. */
int moyolo_decoder_cluster_limits(int rows_per_tile, int* max_clusters, int* kv_cap);

/* ---------------------------------------------------------------------------------------------
 * Fixed-Size Query Memory, one `FSQM.online_update` (MOTR/models/fsqm.py:155-180): (1) update_confidence of the
 * tracked slots (:117-132), (2) inject_new_queries into the free slots in index order (:46-100; score >
 * in_threshold; "memory full -> not injected", :78-81), (3) remove_inactive_queries (:102-115). State (device,
 * caller-owned, N = max_num_queries): query_memory fp32 [N, d], confidence fp32 [N], ids int64 [N] (-1 = free),
 * bounding_boxes fp32 [N, 4], consecutive_low_frames int32 [N], id_pool int64 [2N] ring buffer of the FIFO
 * `global_id_pool` with id_pool_header int32 {head, count} (initially pool = 0..N-1, header = {0, N}).
 * Semantics: the repaired specification F1-F3 stated in oracle/fsqm_port.py (identical to the shipped class
 * wherever that class is self-consistent). One single-CTA launch, no host synchronisation.
 * -------------------------------------------------------------------------------------------*/
int moyolo_fsqm_update(int max_num_queries, int feature_dim, float in_threshold, float out_threshold,
                       int consecutive_frames, float* query_memory, float* confidence, int64_t* ids,
                       float* bounding_boxes, int32_t* consecutive_low_frames, int64_t* id_pool,
                       int32_t* id_pool_header, int n_track, const int64_t* track_ids, const float* track_scores,
                       const float* track_boxes, int n_detect, const float* detect_embedding,
                       const float* detect_scores, const float* detect_boxes, moyolo_stream_t stream);

/* fp32 operands on the bf16 tensor cores: three-term bf16 expansion of an fp32 matrix [M, K] laid out as six column
 * blocks [M, 6K] (role 0 = activation order, 1 = weight order) such that ONE bf16 GEMM over K' = 6K (moyolo_linear,
 * fp32 accumulate) equals the fp32 product to ~3e-8 relative to rms. The fp32 precision mode of the drop-in modules
 * uses it for every nn.Linear with K % 64 == 0, N % 32 == 0 (moyolo_b200.ops.linear). */
int moyolo_split_bf16x3(const float* x, int64_t ldx, void* out, int64_t ldo, int64_t M, int K, int role,
                        moyolo_stream_t stream);

/* Final gather of the per-rank track tables (moyolo_b200/sharding.py; the reference evaluates its videos one after
 * the other on one GPU, MOTR/submit_dance.py:499-504): table_pack builds the fixed-capacity send buffer
 * [capacity + 1, 9] = header row {row count, overflow flag} + rows (n_rows_dev, when not NULL, is a device int32
 * holding the row count -- the engine's table cursor -- and overrides n_rows; overflow_dev, when not NULL, is a device
 * int32 that also raises the overflow flag when non-zero); after ONE all_gather table_merge compacts the
 * [world, capacity + 1, 9] receive buffer into `out` [world * capacity, 9] ordered by rank and writes
 * info = {total rows, overflowed ranks}. Two launches around the collective, no host synchronisation. */
int moyolo_table_pack(const float* rows, int64_t n_rows, const int32_t* n_rows_dev, const int32_t* overflow_dev,
                      int64_t capacity, float* send, moyolo_stream_t stream);
int moyolo_table_merge(const float* recv, int world, int64_t capacity, float* out, int32_t* info, moyolo_stream_t stream);

/* Raw CUDA events that may be recorded inside a captured graph and waited on from outside it (ev_tail_prev above):
 * moyolo_event_record uses cudaEventRecordExternal while `stream` is capturing, a plain record otherwise. */
void* moyolo_event_create(void);
int moyolo_event_destroy(void* ev);
int moyolo_event_record(void* ev, moyolo_stream_t stream);
int moyolo_stream_wait_event(moyolo_stream_t stream, void* ev);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* MOYOLO_B200_H_ */
