#!/usr/bin/env python
"""Warm in-graph latency of the frame's head / tail kernels on a realistic engine state (MOT17, S=1, ~50 tracks):
each kernel N times back to back in one CUDA graph. Profiling script, not product code."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from moyolo_b200 import ops, synthetic as syn  # noqa: E402
from moyolo_b200.tracker import DecoderWeights, TrackEngine  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, n=30, reps=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / n)
    return round(best * 1e3, 2)


def main():
    spec = syn.DecoderSpec()
    shapes = [list(s) for s in syn.PYRAMIDS["MOT17"]]
    sd = syn.make_decoder_state(spec, 0)
    eng0 = TrackEngine(sd, spec, shapes, dev, "bf16", 300, 1)
    g = syn.SequenceGenerator(syn.SequenceSpec("MOT17", 1, 300, 0, shapes=shapes), spec.d_model, dev)
    f, de, dr = g.next_frame()
    out = eng0.step(f[None], de[None], dr[None])[0]
    sd = syn.calibrate_score_bias(sd, out["logits"], spec, 0.035)
    W = DecoderWeights(sd, spec, dev, "bf16")
    eng = TrackEngine(sd, spec, shapes, dev, "bf16", 300, 1, weights=W)
    gen = syn.SequenceGenerator(syn.SequenceSpec("MOT17", 40, 300, 1, shapes=shapes), spec.d_model, dev)
    for _ in range(40):
        fr = gen.next_frame()
        eng.submit(fr[0][None].to(torch.bfloat16).contiguous(), fr[1][None].contiguous(), fr[2][None].contiguous(),
                   want_rows=False)
    eng.drain()
    torch.cuda.synchronize()
    p = eng._last_plan
    ws, S, C, R, Wt = p.ws, 1, spec.d_model, p.rows_pad, eng.W
    boxes = ws.refer[spec.n_layers]
    st, ft, mt, it = eng.thr
    res = {"rows_pad": R, "tracks": eng.n_tracks_host()}
    res["frame_assemble"] = timed(lambda: ops.frame_assemble(
        S, 300, C, eng.cap, eng.n_tracks, eng.t_ref, eng.t_qpos, eng.t_label, eng.t_ids, eng.t_dis, Wt.class_embed,
        eng.det_embed_in[0], eng.det_refer_in[0], ws.x, ws.refer_logit, ws.pos, ws.ids0, ws.dis0, ws.ro, R, ctrl=eng.ctrl,
        refer_sig=ws.refer[0], x_lp=ws.x_lp, xq_lp=ws.xq_lp))
    res["score_head"] = timed(lambda: ops.score_head(ws.x_lp, Wt.score_w, Wt.score_b, out=(ws.logits, ws.scores, ws.labels)))
    res["frame_assign_compact"] = timed(lambda: ops.frame_assign_compact(
        S, C, eng.cap, R, ws.ro, ws.scores, ws.ids0, ws.dis0, eng.counters, ws.ids, ws.dis, ws.labels, ws.refer_logit,
        ws.pos, ws.x, boxes, ws.n_active, ws.active_index, ws.c_ref, ws.c_pos, ws.c_hs, ws.c_box, eng.t_label, eng.t_ids,
        eng.t_dis, st, ft, mt, ctrl=eng.ctrl, q_qk_lp=ws.q_qk_lp, q_tgt_lp=ws.q_tgt_lp))
    cnt = eng.counters.clone()
    res["track_suppress_batched"] = timed(lambda: ops.track_suppress_batched(boxes, ws.ids, cnt, ws.ro, S, ws.rows_per_seq,
                                                                              ws.assign_ws, it, ctrl=eng.ctrl))
    res["frame_writeback"] = timed(lambda: ops.frame_writeback(S, C, eng.cap, ws.ro, ws.n_active, ws.q_new, ws.c_box,
                                                                eng.t_qpos, eng.t_ref, eng.n_tracks, ctrl=eng.ctrl,
                                                                info=p.info))
    res["box_refine"] = timed(lambda: ops.box_refine(ws.bh2, Wt.bbox[5].last_w, Wt.bbox[5].last_b, ws.refer[5],
                                                      out=ws.refer[6]))
    print(json.dumps(res, indent=1))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "tail_latency.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
