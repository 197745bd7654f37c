#!/usr/bin/env python
"""Where does one frame's time go? (profiling script, not product code)

For the MOT17 workload at S lock-step sequences:
  host_us      host wall time of one TrackEngine.submit() (no device sync): if >= device time the loop is host-bound
  replay_us    device time of the captured frame graph replayed back to back (state restored before every
               block of replays so the row count stays inside the plan)
  replay_nobranch_us   the same with the value-projection / box-head side branches folded into one stream
  eager_us     the un-captured launch chain (host launch-bound)
Writes gpurun_out/frame_breakdown_S{S}.json.
"""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from moyolo_b200 import synthetic as syn  # noqa: E402
from moyolo_b200.tracker import DecoderWeights, TrackEngine  # noqa: E402

dev = torch.device("cuda:0")


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n_frames = 60
    spec = syn.DecoderSpec()
    shapes = [list(s) for s in syn.PYRAMIDS["MOT17"]]
    sd = syn.make_decoder_state(spec, 0)
    eng0 = TrackEngine(sd, spec, shapes, dev, "bf16", 300, 1)
    g = syn.SequenceGenerator(syn.SequenceSpec("MOT17", 1, 300, 0, shapes=shapes), spec.d_model, dev)
    f, de, dr = g.next_frame()
    out = eng0.step(f[None], de[None], dr[None])[0]
    sd = syn.calibrate_score_bias(sd, out["logits"], spec, 0.035)
    W = DecoderWeights(sd, spec, dev, "bf16")
    res = {"S": S}
    gens = [syn.SequenceGenerator(syn.SequenceSpec("MOT17", n_frames, 300, 1 + s, shapes=shapes), spec.d_model, dev)
            for s in range(S)]
    frames = []
    for _ in range(n_frames):
        fr = [gg.next_frame() for gg in gens]
        frames.append(tuple(torch.stack([x[i] for x in fr]).to(torch.bfloat16 if i == 0 else torch.float32).contiguous()
                            for i in range(3)))
    for name, kw in (("branch", {}), ("nobranch", {"branches": False})):
        eng = TrackEngine(sd, spec, shapes, dev, "bf16", 300, S, weights=W, **kw)
        eng.prepare(160)
        # steady state: 40 frames in
        for t in range(40):
            eng.submit(*frames[t], want_rows=False)
        eng.drain()
        torch.cuda.synchronize()
        # host time per submit vs device time over the next 20 frames
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for t in range(40, 60):
            eng.submit(*frames[t], want_rows=False)
        host = (time.perf_counter() - t0) / 20
        e1.record()
        eng.drain()
        torch.cuda.synchronize()
        res[f"{name}_host_us_per_submit"] = round(host * 1e6, 1)
        res[f"{name}_device_us_per_frame_pipelined"] = round(e0.elapsed_time(e1) * 1e3 / 20, 1)
        res[f"{name}_tracks"] = eng.n_tracks_host()
        # pure graph replay, state restored so every replay sees the same row count
        rows = sum(eng._T) + S * 300
        p = eng._plan(eng._round(rows), 0)
        eng.feats_in[0].copy_(frames[59][0]); eng.det_embed_in[0].copy_(frames[59][1]); eng.det_refer_in[0].copy_(frames[59][2])
        snap = eng._state_snapshot()
        ts = []
        for _ in range(10):
            eng._state_restore(snap)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); p.graph.replay(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        res[f"{name}_single_replay_us_median"] = round(ts[len(ts) // 2], 1)
        # 20 replays back to back on one stream (no copies, no event hand-offs, no value projection): what a graph
        # boundary itself costs
        eng._state_restore(snap)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            p.graph.replay()
        b.record()
        torch.cuda.synchronize()
        res[f"{name}_back_to_back_us_per_replay"] = round(a.elapsed_time(b) * 1e3 / 20, 1)
        res[f"{name}_rows_pad"] = p.rows_pad
        res[f"{name}_launches"] = p.n_launch
        del eng
    print(json.dumps(res, indent=1))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"frame_breakdown_S{S}.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
