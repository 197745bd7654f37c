#!/usr/bin/env python
"""In-graph timeline of one frame (profiling script, not product code).

Captures the frame body truncated after k C-ABI launches, for k = 1..n, as CUDA graphs (side branches folded
into the main stream so the order is linear) and replays each one; the difference between consecutive
prefixes is the time the k-th kernel adds to the frame INSIDE the graph (warm caches, programmatic
dependent launch overlap included) -- unlike the ncu launch list, whose per-kernel times are cold and
serialised. Writes gpurun_out/frame_timeline_S{S}.json.
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from moyolo_b200 import _lib, synthetic as syn  # noqa: E402
from moyolo_b200.tracker import DecoderWeights, TrackEngine, _FramePlan  # noqa: E402

dev = torch.device("cuda:0")
_NOT_LAUNCH = {"moyolo_last_error", "moyolo_version", "moyolo_track_workspace_bytes", "moyolo_device_supported",
               "moyolo_event_record", "moyolo_event_create", "moyolo_event_destroy", "moyolo_stream_wait_event"}


class Proxy:
    """Stands in for the CDLL handle: launches beyond `limit` become no-ops; records the call names."""

    def __init__(self, real):
        self.real, self.limit, self.count, self.names = real, 1 << 30, 0, []

    def __getattr__(self, name):
        fn = getattr(self.real, name)
        if name in _NOT_LAUNCH:
            return fn

        def call(*a):
            self.count += 1
            if self.count > self.limit:
                return 0
            self.names.append(name)
            return fn(*a)
        return call


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    branches = len(sys.argv) > 2 and sys.argv[2] == "branch"   # keep the side branches: deltas = critical-path growth
    reps = 15
    spec = syn.DecoderSpec()
    shapes = [list(s) for s in syn.PYRAMIDS["MOT17"]]
    sd = syn.make_decoder_state(spec, 0)
    eng0 = TrackEngine(sd, spec, shapes, dev, "bf16", 300, 1)
    g = syn.SequenceGenerator(syn.SequenceSpec("MOT17", 1, 300, 0, shapes=shapes), spec.d_model, dev)
    f, de, dr = g.next_frame()
    out = eng0.step(f[None], de[None], dr[None])[0]
    sd = syn.calibrate_score_bias(sd, out["logits"], spec, 0.035)
    W = DecoderWeights(sd, spec, dev, "bf16")
    gens = [syn.SequenceGenerator(syn.SequenceSpec("MOT17", 41, 300, 1 + s, shapes=shapes), spec.d_model, dev)
            for s in range(S)]
    eng = TrackEngine(sd, spec, shapes, dev, "bf16", 300, S, weights=W, branches=branches)
    eng.prepare(160)
    last = None
    for _ in range(41):
        fr = [gg.next_frame() for gg in gens]
        last = tuple(torch.stack([x[i] for x in fr]).to(torch.bfloat16 if i == 0 else torch.float32).contiguous()
                     for i in range(3))
        eng.submit(*last, want_rows=False)
    eng.drain()
    torch.cuda.synchronize()
    rows = sum(eng._T) + S * 300
    base = eng._plan(eng._round(rows), 0)
    eng.feats_in[0].copy_(last[0]); eng.det_embed_in[0].copy_(last[1]); eng.det_refer_in[0].copy_(last[2])
    snap = eng._state_snapshot()
    proxy = Proxy(_lib.lib())
    _lib._lib = proxy
    n_total = base.n_launch
    cum, names = [], None
    for k in range(1, n_total + 1):
        p = _FramePlan()
        p.rows_pad, p.slot, p.ws, p.graph, p.n_launch, p.desc = base.rows_pad, 0, base.ws, None, 0, None
        p.info, p.frame_rows = base.info, base.frame_rows
        proxy.limit, proxy.count, proxy.names = k, 0, []
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            eng._body(p)
        if k == n_total:
            names = list(proxy.names)
        ts = []
        for _ in range(reps):
            eng._state_restore(snap)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gr.replay(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        cum.append(ts[len(ts) // 2])
        del gr
    _lib._lib = proxy.real
    rows_out = []
    prev = 0.0
    for k, (n, c) in enumerate(zip(names, cum)):
        rows_out.append({"k": k + 1, "call": n.replace("moyolo_", ""), "cum_us": round(c, 2), "delta_us": round(c - prev, 2)})
        prev = c
    agg = {}
    for r in rows_out:
        agg[r["call"]] = round(agg.get(r["call"], 0.0) + r["delta_us"], 2)
    res = {"S": S, "branches": branches, "rows_pad": base.rows_pad, "tracks": eng.n_tracks_host(), "launches": n_total,
           "note": "prefix-graph replay medians; delta = time the k-th launch adds inside the (branch-free) frame graph; "
                   "the first entry includes the graph-launch overhead",
           "by_call_us": dict(sorted(agg.items(), key=lambda kv: -kv[1])), "timeline": rows_out}
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"frame_timeline_S{S}{'_branch' if branches else ''}.json").write_text(json.dumps(res, indent=1))
    print(json.dumps({k: v for k, v in res.items() if k != "timeline"}, indent=1))
    for r in rows_out:
        print(f"{r['k']:3d} {r['call']:28s} {r['delta_us']:7.2f} {r['cum_us']:8.2f}")


if __name__ == "__main__":
    main()
