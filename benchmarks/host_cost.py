#!/usr/bin/env python
"""Host-side cost of one TrackEngine.submit() and of the native moyolo_frame_submit call inside it (MOT17, S=1)."""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from moyolo_b200 import _lib, synthetic as syn  # noqa: E402
from moyolo_b200.tracker import DecoderWeights, TrackEngine  # noqa: E402

dev = torch.device("cuda:0")
spec = syn.DecoderSpec()
shapes = [list(s) for s in syn.PYRAMIDS["MOT17"]]
sd = syn.make_decoder_state(spec, 0)
W = DecoderWeights(sd, spec, dev, "bf16")
eng = TrackEngine(sd, spec, shapes, dev, "bf16", 300, 1, weights=W)
eng.prepare(160)
gen = syn.SequenceGenerator(syn.SequenceSpec("MOT17", 100, 300, 1, shapes=shapes), spec.d_model, dev)
frames = []
for _ in range(100):
    fr = gen.next_frame()
    frames.append((fr[0][None].to(torch.bfloat16).contiguous(), fr[1][None].contiguous(), fr[2][None].contiguous()))
real = _lib.lib().moyolo_frame_submit
acc = {"native": 0.0, "n": 0}


class Wrap:
    def __getattr__(self, name):
        fn = getattr(_lib._lib_real, name)
        if name != "moyolo_frame_submit":
            return fn

        def timed(*a):
            t0 = time.perf_counter()
            r = fn(*a)
            acc["native"] += time.perf_counter() - t0
            acc["n"] += 1
            return r
        return timed


_lib._lib_real = _lib._lib
_lib._lib = Wrap()
for t in range(30):
    eng.submit(*frames[t], want_rows=False)
eng.drain()
torch.cuda.synchronize()
acc.update(native=0.0, n=0)
t0 = time.perf_counter()
for t in range(30, 100):
    eng.submit(*frames[t], want_rows=False)
host = (time.perf_counter() - t0) / 70
eng.drain()
# pure host cost without a device to wait for: submit while the device is idle (drain after every frame)
idle = 0.0
for t in range(30):
    eng.drain()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    eng.submit(*frames[t], want_rows=False)
    idle += time.perf_counter() - t1
print(json.dumps({"submit_us_pipelined": round(host * 1e6, 1), "native_call_us": round(acc["native"] / acc["n"] * 1e6, 1),
                  "submit_us_device_idle": round(idle / 30 * 1e6, 1)}))
