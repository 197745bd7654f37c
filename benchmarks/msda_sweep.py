#!/usr/bin/env python
"""MSDeformAttn gather microbenchmark sweep (BASELINE.json configs[4]).

Queries 300-1200, levels 3-4, points 4-8, batch 1-16, fp32 vs bf16, on the 640x640 pyramid (+ a 10x10
level when L=4). For every point: the fused gather kernel (softmax + locations + bilinear gather) timed
  warm  — 20 back-to-back launches in one CUDA graph, value L2-resident (how it runs inside a frame)
  cold  — a 256 MB L2 flush before every launch, CUDA events around the launch alone
and, in fp32 on pre-normalised inputs, side by side with (i) the reference's own CUDA kernel compiled
for sm_100a (oracle/_ref) and (ii) the reference's PyTorch grid_sample path on the same GPU.
Bandwidths use the compulsory bytes of SURVEY.md §8(d); `touched` = bytes of all sampled corner rows.
Writes gpurun_out/msda_sweep.json (+ a markdown table on stdout). Profiling script, not product code.
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from moyolo_b200 import ops, synthetic as syn  # noqa: E402
from oracle import ref_cuda, torch_port as tp  # noqa: E402

dev = torch.device("cuda:0")
PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def time_warm(fn, n=20, reps=5):
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
    return best * 1e3  # us


def time_cold(fn, reps=7):
    ts = []
    for _ in range(reps):
        flush_buf.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    H, D = 8, 32
    C = H * D
    rows = []
    for L, P in ((3, 4), (4, 8)):
        shapes = [list(s) for s in syn.PYRAMIDS["C1"]] + ([[10, 10]] if L == 4 else [])
        Lv = syn.level_sizes(shapes)
        for B in (1, 4, 16):
            for Q in (300, 600, 1200):
                g = torch.Generator().manual_seed(Q + B)
                R = B * Q
                value32 = torch.randn(B, Lv, C, generator=g).to(dev)
                offsets = torch.randn(R, H * L * P * 2, generator=g).to(dev)
                logits = torch.randn(R, H * L * P, generator=g).to(dev)
                cxcy = torch.rand(R, 1, 2, generator=g)
                wh = torch.rand(R, 1, 2, generator=g) * 0.48 + 0.02
                refer = torch.cat([cxcy, wh], -1).to(dev)
                for dt, name in ((torch.bfloat16, "bf16"), (torch.float32, "fp32")):
                    value = value32.to(dt)
                    s = value.element_size()
                    out = torch.empty(R, C, dtype=dt, device=dev)
                    fn = lambda: ops.msda_fused(value, shapes, offsets, logits, refer, H, P, B, out=out)  # noqa: E731
                    comp = B * Lv * C * s + R * H * L * P * 12 + R * 16 + R * C * s
                    touched = R * H * L * P * 4 * D * s
                    warm, cold = time_warm(fn), time_cold(fn)
                    rec = {"L": L, "P": P, "B": B, "Q": Q, "dtype": name, "compulsory_bytes": comp,
                           "touched_bytes": touched, "warm_us": round(warm, 2), "cold_us": round(cold, 2),
                           "warm_GBs": round(comp / warm / 1e3, 1), "cold_GBs": round(comp / cold / 1e3, 1),
                           "cold_frac_of_measured_peak": round(comp / cold / 1e3 / PEAK, 3),
                           "warm_touched_GBs": round(touched / warm / 1e3, 1)}
                    if name == "fp32":  # same-input comparison with the reference implementations
                        loc = (refer[:, :, None, :2].view(R, 1, 1, 1, 2) + offsets.view(R, H, L, P, 2) / P *
                               refer[:, 0, 2:].view(R, 1, 1, 1, 2) * 0.5).view(B, Q, H, L, P, 2).contiguous()
                        w = torch.softmax(logits.view(R, H, L * P), -1).view(B, Q, H, L, P).contiguous()
                        v4 = value.view(B, Lv, H, D)
                        ours = lambda: ops.msda_sampled(v4, shapes, loc, w)  # noqa: E731
                        rec["sampled_warm_us"] = round(time_warm(ours), 2)
                        rec["sampled_cold_us"] = round(time_cold(ours), 2)
                        if ref_cuda.available():
                            rk = lambda: ref_cuda.msda_im2col(v4, shapes, loc, w)  # noqa: E731
                            err = float((rk() - ours()).abs().max())
                            rec["refcuda_max_abs_diff"] = err
                            rec["refcuda_warm_us"] = round(time_warm(rk), 2)
                            rec["refcuda_cold_us"] = round(time_cold(rk), 2)
                        # backward: ours vs the reference's col2im kernel, both incl. the zero-fill
                        go = torch.randn(B, Q, C, generator=g).to(dev)
                        bw = lambda: ops.msda_sampled_backward(v4, shapes, loc, w, go)  # noqa: E731
                        rec["bwd_warm_us"] = round(time_warm(bw, n=5), 2)
                        if ref_cuda.available():
                            rbw = lambda: ref_cuda.msda_col2im(v4, shapes, loc, w, go)  # noqa: E731
                            rec["refcuda_bwd_warm_us"] = round(time_warm(rbw, n=5), 2)
                            rec["refcuda_bwd_max_rel_diff"] = max(
                                float((x - y).abs().max() / (y.pow(2).mean().sqrt() + 1e-30)) for x, y in zip(bw(), rbw()))
                        if B * Q <= 4800:
                            pt = lambda: tp.msda_core_gridsample(v4, shapes, loc, w)  # noqa: E731
                            rec["torch_gridsample_warm_us"] = round(time_warm(pt, n=3, reps=3), 2)
                    rows.append(rec)
                    print(json.dumps(rec), flush=True)
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    (out_dir / "msda_sweep.json").write_text(json.dumps({"peak_hbm_gbs_measured": PEAK, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
