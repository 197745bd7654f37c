#!/usr/bin/env python
"""Does a head-major value layout [B, H, Lv, 32] (x0 / x1 corners contiguous) make the deformable gather faster than
the channel-last layout it reads today (one layer's 256-column slice of the [B, Lv, 6*256] all-layers tensor)?
Same kernel (two items per warp), same inputs, bit-identical outputs required; warm (back-to-back in a graph) and cold
(256 MB L2 flush before every launch). VERDICT r1 task 3(a). Profiling script, not product code."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from moyolo_b200 import ops, synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def warm(fn, n=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / n)
    return best * 1e3


def cold(fn, reps=9):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


H, D, L, P, n_layers = 8, 32, 3, 4, 6
C = H * D
for name in ("MOT17", "DanceTrack"):
    shapes = [list(s) for s in syn.PYRAMIDS[name]]
    Lv = syn.level_sizes(shapes)
    for B, Q in ((16, 300), (8, 382), (4, 382)):
        g = torch.Generator().manual_seed(B + Q)
        R = B * Q
        allv = torch.randn(B, Lv, n_layers * C, generator=g).to(dev).to(torch.bfloat16)   # as the value projection writes it
        v_cl = allv[:, :, 2 * C:3 * C]                                                     # layer 2's slice, in place
        v_hm = v_cl.reshape(B, Lv, H, D).permute(0, 2, 1, 3).contiguous()                  # [B, H, Lv, D]
        offsets = torch.randn(R, H * L * P * 2, generator=g).to(dev)
        logits = torch.randn(R, H * L * P, generator=g).to(dev)
        refer = torch.cat([torch.rand(R, 1, 2, generator=g), torch.rand(R, 1, 2, generator=g) * 0.3 + 0.02], -1).to(dev)
        o1 = torch.empty(R, C, dtype=torch.bfloat16, device=dev)
        o2 = torch.empty_like(o1)
        f_cl = lambda: ops.msda_fused(v_cl, shapes, offsets, logits, refer, H, P, B, out=o1)          # noqa: E731
        f_hm = lambda: ops.msda_fused_headmajor(v_hm, shapes, offsets, logits, refer, P, out=o2)      # noqa: E731
        f_cl(); f_hm(); torch.cuda.synchronize()
        same = bool(torch.equal(o1, o2))
        comp = B * Lv * C * 2 + R * H * L * P * 12 + R * 16 + R * C * 2
        rec = {"pyramid": name, "B": B, "rows": R, "bytes": comp, "bit_equal": same}
        for tag, f in (("channel_last", f_cl), ("head_major", f_hm)):
            w, c = warm(f), cold(f)
            rec[tag] = {"warm_us": round(w, 2), "cold_us": round(c, 2), "warm_frac": round(comp / w / 1e3 / PEAK, 3),
                        "cold_frac": round(comp / c / 1e3 / PEAK, 3)}
        print(json.dumps(rec), flush=True)
