#!/usr/bin/env python
"""Soak run (not a test, not a benchmark): two lock-step MOT17-shaped sequences for many frames with a score head
biased so that the track count climbs well past the pre-captured graph sizes (lazy plan capture, speculation aborts,
re-launches), periodic per-sequence resets, results collected every frame; cross-checks the device track table against
the per-frame rows. Prints one summary line."""
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
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    S = 2
    spec = syn.DecoderSpec()
    shapes = [list(s) for s in syn.PYRAMIDS["MOT17"]]
    sd = syn.make_decoder_state(spec, 0)
    eng0 = TrackEngine(sd, spec, shapes, dev, "bf16", 300, 1)
    g = syn.SequenceGenerator(syn.SequenceSpec("MOT17", 1, 300, 0, shapes=shapes), spec.d_model, dev)
    f, de, dr = g.next_frame()
    out = eng0.step(f[None], de[None], dr[None])[0]
    sd = syn.calibrate_score_bias(sd, out["logits"], spec, 0.12)   # ~3x the bench's birth rate
    del eng0
    W = DecoderWeights(sd, spec, dev, "bf16")
    eng = TrackEngine(sd, spec, shapes, dev, "bf16", 300, S, weights=W, margin=16)
    eng.prepare(64)   # deliberately too small: larger plans are captured lazily inside the loop
    gens = [syn.SequenceGenerator(syn.SequenceSpec("MOT17", n_frames, 300, 1 + s, shapes=shapes), spec.d_model, dev)
            for s in range(S)]
    t0 = time.time()
    active_rows, max_tracks, resets = 0, 0, 0
    for t in range(n_frames):
        fr = [gg.next_frame() for gg in gens]
        batch = tuple(torch.stack([x[i] for x in fr]).to(torch.bfloat16 if i == 0 else torch.float32).contiguous()
                      for i in range(3))
        if t and t % 150 == 0:
            eng.reset(seq=t // 150 % S)
            resets += 1
        eng.submit(*batch, want_rows=True)
        if t > 0:
            for o in eng.collect(t - 1):
                active_rows += int((o["ids"] >= 0).sum())
        max_tracks = max(max_tracks, max(eng._T))
    for o in eng.collect(n_frames - 1):
        active_rows += int((o["ids"] >= 0).sum())
    table = eng.track_table()
    ok = int(table.shape[0]) == active_rows
    ids_ok = bool((table[:, 2] >= 0).all())
    print(json.dumps({"frames": n_frames, "S": S, "seconds": round(time.time() - t0, 1), "plans": len(eng._plans),
                      "aborts": eng.aborts, "resets": resets, "max_tracks_per_seq": max_tracks,
                      "table_rows": int(table.shape[0]), "rows_from_frames": active_rows, "consistent": ok and ids_ok}))
    assert ok and ids_ok


if __name__ == "__main__":
    main()
