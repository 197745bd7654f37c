#!/usr/bin/env python
"""Why does the device-resident loop differ from the host-input loop? (profiling script)
Runs the 300-frame MOT17 loop of bench.py in several submission modes and prints ms/frame."""
import json
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from moyolo_b200 import synthetic as syn  # noqa: E402
from moyolo_b200.tracker import DecoderWeights, TrackEngine  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda:0")
K = 300


class A:
    workload, n_detect, precision, seqs_per_gpu = "MOT17", 300, "bf16", 1


def main():
    spec = syn.DecoderSpec()
    shapes = [list(s) for s in syn.PYRAMIDS["MOT17"]]
    sd = bench.build_state(A, spec, syn, shapes, dev)
    W = DecoderWeights(sd, spec, dev, "bf16")
    frames = bench.make_frames(A, syn, spec, shapes, dev, 1, K, torch.bfloat16)
    frames = [tuple(x[None].contiguous() for x in f) for f in frames]
    host = [tuple(x.cpu().pin_memory() for x in f) for f in frames]
    res = {}
    for name, kw, src, block in (("dev_runahead", {}, frames, False), ("dev_margin0", {"margin": 0}, frames, False),
                                 ("dev_blocking", {}, frames, True), ("host_runahead", {}, host, False),
                                 ("host_blocking", {}, host, True), ("dev_bucket32", {"bucket": 32, "margin": 16}, frames, False)):
        eng = TrackEngine(sd, spec, shapes, dev, "bf16", 300, 1, weights=W, **kw)
        eng.prepare(160)
        for rep in range(2):
            eng.reset()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for t in range(K):
                eng.submit(*src[t], want_rows=block)
                if block and t > 0:
                    eng.collect(t - 1)
            eng.drain()
            b.record()
            torch.cuda.synchronize()
        res[name] = {"ms_per_frame": round(a.elapsed_time(b) / K, 4), "aborts": eng.aborts,
                     "plans": sorted({k[0] for k in eng._plans})}
        del eng
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
