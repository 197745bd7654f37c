#!/usr/bin/env python
"""Fused gather alone at a C5 sweep point (profiling script): B sequences x Q queries on the C1 pyramid."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from moyolo_b200 import ops, synthetic as syn  # noqa: E402
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 300
wl = sys.argv[3] if len(sys.argv) > 3 else "C1"
H, D, L, P = 8, 32, 3, 4
C = H * D
shapes = [list(s) for s in syn.PYRAMIDS[wl]]
Lv = syn.level_sizes(shapes)
g = torch.Generator().manual_seed(1)
R = B * Q
value = torch.randn(B, Lv, C, generator=g).to(dev).bfloat16()
offsets = torch.randn(R, H * L * P * 2, generator=g).to(dev)
logits = torch.randn(R, H * L * P, generator=g).to(dev)
refer = torch.cat([torch.rand(R, 1, 2, generator=g), torch.rand(R, 1, 2, generator=g) * 0.48 + 0.02], -1).to(dev)
out = torch.empty(R, C, dtype=torch.bfloat16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    ops.msda_fused(value, shapes, offsets, logits, refer, H, P, B, out=out)
ts = []
for _ in range(10):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.msda_fused(value, shapes, offsets, logits, refer, H, P, B, out=out); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
comp = B * Lv * C * 2 + R * H * L * P * 12 + R * 16 + R * C * 2
print(f"B={B} Q={Q} cold median {ts[5]:.2f} us  compulsory {comp/1e6:.1f} MB -> {comp/ts[5]/1e3:.0f} GB/s")
