// Host-side frame runtime: everything the host does per frame, in ONE call.
//
// A frame of the tracker is a captured CUDA graph plus a handful of stream operations around it
// (input copies into the device ring on a copy stream, event hand-offs, result copies to pinned host
// memory). Issued from Python these ~12 driver calls cost more host time than the GPU needs for the
// frame itself; here they are one C call. No reference counterpart: the reference's frame loop
// (MOTRtrack/val.py:288-291 -> TrackingModel.predict, ultralytics/nn/tasks.py:513) is eager PyTorch.
#include "common.cuh"

using namespace moyolo;

#define RT_CHECK(call, what)                                                                 \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) return fail(MOYOLO_ERR_CUDA, "runtime: %s: %s", what, cudaGetErrorString(e_)); \
  } while (0)

extern "C" int moyolo_frame_submit(const moyolo_frame_submit_t* d) {
  MOYOLO_REQUIRE(d != nullptr && d->main_stream_valid == 1, MOYOLO_ERR_BAD_ARG, "frame_submit: null descriptor");
  MOYOLO_REQUIRE(d->n_inputs >= 0 && d->n_inputs <= MOYOLO_SUBMIT_MAX_COPIES && d->n_outputs >= 0 &&
                     d->n_outputs <= MOYOLO_SUBMIT_MAX_COPIES,
                 MOYOLO_ERR_BAD_ARG, "frame_submit: too many copies");
  cudaStream_t cs = static_cast<cudaStream_t>(d->copy_stream);
  cudaStream_t ms = static_cast<cudaStream_t>(d->main_stream);
  if (d->n_inputs > 0) {
    // the ring slot is free once the frame that last read it has finished
    if (d->ev_slot_free != nullptr) RT_CHECK(cudaStreamWaitEvent(cs, static_cast<cudaEvent_t>(d->ev_slot_free), 0), "wait slot");
    if (d->sync_inputs) {  // inputs still being produced on the main stream
      MOYOLO_REQUIRE(d->ev_scratch != nullptr, MOYOLO_ERR_BAD_ARG, "frame_submit: sync_inputs needs ev_scratch");
      RT_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(d->ev_scratch), ms), "record scratch");
      RT_CHECK(cudaStreamWaitEvent(cs, static_cast<cudaEvent_t>(d->ev_scratch), 0), "wait scratch");
    }
    for (int i = 0; i < d->n_inputs; ++i)
      RT_CHECK(cudaMemcpyAsync(d->in_dst[i], d->in_src[i], static_cast<size_t>(d->in_bytes[i]), cudaMemcpyDefault, cs),
               "input copy");
    RT_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(d->ev_copy), cs), "record copy");
    RT_CHECK(cudaStreamWaitEvent(ms, static_cast<cudaEvent_t>(d->ev_copy), 0), "wait copy");
  }
  if (d->vp_valid == 1 || d->vp_valid == 2) {
    MOYOLO_REQUIRE(d->ev_vp != nullptr && (d->vp_valid == 2 ? d->pre_graph_exec != nullptr : (d->vp_x && d->vp_w && d->vp_y)),
                   MOYOLO_ERR_BAD_ARG, "frame_submit: value projection needs ev_vp and x / w / y (or a prelude graph)");
    cudaStream_t vs = static_cast<cudaStream_t>(d->vp_stream);
    if (d->n_inputs > 0) RT_CHECK(cudaStreamWaitEvent(vs, static_cast<cudaEvent_t>(d->ev_copy), 0), "vp wait copy");
    if (d->ev_tail_prev != nullptr)
      RT_CHECK(cudaStreamWaitEvent(vs, static_cast<cudaEvent_t>(d->ev_tail_prev), 0), "vp wait previous tail");
    if (d->vp_valid == 2) {
      RT_CHECK(cudaGraphLaunch(static_cast<cudaGraphExec_t>(d->pre_graph_exec), vs), "prelude graph launch");
    } else {
      const int rc = moyolo_linear_tall(d->vp_x, d->vp_ldx, d->vp_w, d->vp_bias, d->vp_y, d->vp_ldy, d->vp_M, d->vp_N, 256,
                                        nullptr, d->vp_max_ctas, d->vp_stream);
      if (rc != MOYOLO_OK) return rc;
    }
    RT_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(d->ev_vp), vs), "record vp");
    RT_CHECK(cudaStreamWaitEvent(ms, static_cast<cudaEvent_t>(d->ev_vp), 0), "wait vp");
  }
  if (d->graph_exec != nullptr) RT_CHECK(cudaGraphLaunch(static_cast<cudaGraphExec_t>(d->graph_exec), ms), "graph launch");
  cudaStream_t os = ms;
  if (d->out_stream_valid == 1) {  // result copies off the main stream: the next frame's graph follows this one directly
    MOYOLO_REQUIRE(d->ev_graph != nullptr, MOYOLO_ERR_BAD_ARG, "frame_submit: out_stream needs ev_graph");
    os = static_cast<cudaStream_t>(d->out_stream);
    RT_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(d->ev_graph), ms), "record graph");
    RT_CHECK(cudaStreamWaitEvent(os, static_cast<cudaEvent_t>(d->ev_graph), 0), "wait graph");
  }
  for (int i = 0; i < d->n_outputs; ++i)
    RT_CHECK(cudaMemcpyAsync(d->out_dst[i], d->out_src[i], static_cast<size_t>(d->out_bytes[i]), cudaMemcpyDefault, os),
             "result copy");
  if (d->ev_done != nullptr) RT_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(d->ev_done), os), "record done");
  return MOYOLO_OK;
}

// Raw event helpers for the one event that is recorded INSIDE captured frame graphs and waited on from outside
// (the "tail reached" signal that gates the next frame's value projection). Inside a stream capture the record
// becomes an external event-record node (cudaEventRecordExternal is only valid there); outside it is a plain record.
extern "C" void* moyolo_event_create(void) {
  cudaEvent_t ev = nullptr;
  if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return ev;
}

extern "C" int moyolo_event_destroy(void* ev) {
  if (ev != nullptr) RT_CHECK(cudaEventDestroy(static_cast<cudaEvent_t>(ev)), "event destroy");
  return MOYOLO_OK;
}

extern "C" int moyolo_event_record(void* ev, moyolo_stream_t stream) {
  MOYOLO_REQUIRE(ev != nullptr, MOYOLO_ERR_BAD_ARG, "event_record: null event");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  RT_CHECK(cudaStreamIsCapturing(st, &cs), "capture status");
  RT_CHECK(cudaEventRecordWithFlags(static_cast<cudaEvent_t>(ev), st,
                                    cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault),
           "event record");
  return MOYOLO_OK;
}

extern "C" int moyolo_stream_wait_event(moyolo_stream_t stream, void* ev) {
  MOYOLO_REQUIRE(ev != nullptr, MOYOLO_ERR_BAD_ARG, "stream_wait_event: null event");
  RT_CHECK(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), static_cast<cudaEvent_t>(ev), 0), "stream wait event");
  return MOYOLO_OK;
}
