"""Accuracy of the fp32 Linear on the tensor cores (bf16 three-term split, ops.linear) against an fp64 product,
next to the CUDA-core fp32 kernel, over the shapes the decoder / selector use."""
import sys
import torch
sys.path.insert(0, ".")
from moyolo_b200 import ops, _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
for M, K, N, relu, zr in ((300, 256, 256, False, False), (300, 1024, 256, False, False), (8400, 256, 1536, False, False),
                          (300, 256, 288, False, False), (300, 256, 1024, True, False), (8400, 256, 256, False, False),
                          (8400, 256, 256, True, True), (13566, 512, 256, False, False), (300, 512, 256, False, False),
                          (8400, 256, 1536, False, True)):
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    z = (torch.rand(M, device=dev) < 0.2).to(torch.uint8) if zr else None
    ref = x.double() @ w.double().T + b.double()
    if zr:
        ref = ref * (1 - z.double())[:, None]
    if relu:
        ref = ref.relu()
    for name, eng in (("tensor", _lib.GEMM_AUTO), ("simt", _lib.GEMM_SIMT)):
        y = ops.linear(x, w, b, relu=relu, zero_rows=z, engine=eng)
        print(M, K, N, "relu" if relu else "", "zero_rows" if zr else "", name,
              float((y.double() - ref).abs().max() / ref.pow(2).mean().sqrt()))


def _time(fn, n=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


print("timing (us per call, 20 back-to-back calls, weights expanded once):")
for M, K, N in ((13566, 256, 1536), (340, 256, 256), (340, 1024, 256), (340, 256, 1024)):
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    t = {name: _time(lambda: ops.linear(x, w, b, engine=eng)) for name, eng in (("tensor", _lib.GEMM_AUTO), ("simt", _lib.GEMM_SIMT))}
    print(M, K, N, {k: round(v, 1) for k, v in t.items()})
