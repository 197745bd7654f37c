"""GPU: the tcgen05/TMA/TMEM GEMM engine against an fp64 reference on the same bf16-rounded operands.
Exact products of bf16 values accumulated in fp32: tolerance 2e-5 * rms on fp32 output (accumulation
order only), plus bf16 output rounding (2^-8 relative) when the output is bf16."""
import pytest
import torch

from conftest import rel_rms

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


SHAPES = [  # (M, N, K) — the decoder's GEMMs at the named configs
    (300, 288, 256),      # sampling_offsets|attention_weights, C1
    (357, 512, 256),      # MHA q,k in-proj, MOT17 with tracks
    (300, 256, 1024),     # FFN second layer (K loop wraps the 4-stage ring 4x)
    (300, 1024, 256),     # FFN first layer
    (8400, 1536, 256),    # value_proj of all 6 layers, C1
    (13566, 1536, 256),   # value_proj of all 6 layers, MOT17
    (1, 256, 256), (127, 32, 64), (129, 64, 128), (2049, 128, 64), (4096, 256, 256),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tcgen05_linear(dev, M, N, K):
    from moyolo_b200 import _lib, ops
    g = torch.Generator().manual_seed(M * 7 + N)
    x = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, generator=g)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    xd, wd, bd = x.to(dev), w.to(dev), b.to(dev)
    y = ops.linear(xd, wd, bd, out_dtype=torch.float32, engine=_lib.GEMM_TCGEN05)
    torch.cuda.synchronize()
    assert rel_rms(y.cpu().numpy(), ref.numpy()) < 2e-5
    ys = ops.linear(xd, wd, bd, out_dtype=torch.float32, engine=_lib.GEMM_SIMT)
    assert rel_rms(y.cpu().numpy(), ys.cpu().numpy()) < 2e-5
    yb = ops.linear(xd, wd, bd, out_dtype=torch.bfloat16, relu=True, engine=_lib.GEMM_TCGEN05)
    refb = ref.relu()
    assert float((yb.float().cpu().double() - refb).abs().max()) <= float(refb.abs().max()) * 2 ** -8 + 1e-6
    zr = torch.rand(M, generator=g) < 0.25
    yz = ops.linear(xd, wd, bd, out_dtype=torch.float32, zero_rows=zr.to(dev, torch.uint8), engine=_lib.GEMM_TCGEN05)
    assert torch.equal(yz.cpu()[zr], torch.zeros(int(zr.sum()), N)) and torch.equal(yz.cpu()[~zr], y.cpu()[~zr])


def test_tcgen05_strided_views(dev):
    """x as a column slice of a wider buffer (ldx > K) and y written into a column slice (ldy > N)."""
    from moyolo_b200 import _lib, ops
    g = torch.Generator().manual_seed(5)
    M, K, N = 333, 256, 512
    big = torch.randn(M, 3 * K, generator=g).bfloat16().to(dev)
    w = (torch.randn(N, K, generator=g) / 16).bfloat16().to(dev)
    b = torch.randn(N, generator=g).to(dev)
    out = torch.zeros(M, 768, dtype=torch.bfloat16, device=dev)
    ops.linear(big[:, K:2 * K], w, b, out=out[:, 256:], engine=_lib.GEMM_TCGEN05)
    ref = torch.nn.functional.linear(big[:, K:2 * K].double(), w.double(), b.double()).cpu()
    got = out[:, 256:].float().cpu().double()
    assert float((got - ref).abs().max()) <= float(ref.abs().max()) * 2 ** -8 + 1e-6
    assert torch.equal(out[:, :256].cpu(), torch.zeros(M, 256, dtype=torch.bfloat16))


def test_tcgen05_rejects_bad_shapes(dev):
    from moyolo_b200 import _lib, ops
    x = torch.zeros(8, 100, dtype=torch.bfloat16, device=dev)
    w = torch.zeros(32, 100, dtype=torch.bfloat16, device=dev)
    with pytest.raises(ValueError):
        ops.linear(x, w, None, engine=_lib.GEMM_TCGEN05)  # K % 64 != 0
    y = ops.linear(x, w, None)  # AUTO falls back to the CUDA-core engine
    assert y.shape == (8, 32)
