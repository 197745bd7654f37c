#!/usr/bin/env python
"""Self-attention kernel alone at the frame's shape (profiling script): R rows, 8 heads x 32."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from moyolo_b200 import ops  # noqa: E402
dev = torch.device("cuda:0")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 382
qkv = torch.randn(R, 768, device=dev).bfloat16()
ro = torch.tensor([0, R], dtype=torch.int32, device=dev)
att = torch.empty(R, 256, dtype=torch.bfloat16, device=dev)
for _ in range(5):
    ops.self_attention(qkv[:, :256], qkv[:, 256:512], qkv[:, 512:], ro, [0, R], 8, out=att)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    ops.self_attention(qkv[:, :256], qkv[:, 256:512], qkv[:, 512:], ro, [0, R], 8, out=att)
b.record(); torch.cuda.synchronize()
print("eager us/launch", a.elapsed_time(b) * 20)
