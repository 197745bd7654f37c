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


@pytest.mark.parametrize("M,N", [(4096, 256), (8400, 1536), (13566, 1536), (54264, 1536), (4097, 128)])
def test_stream_gemm_value_proj(dev, M, N):
    """Persistent weight-resident kernel (tall x, K = 256, bf16 out through TMA stores), incl. the value_mask
    row zeroing (transformer.py:265-266) and writing into a column slice of a wider tensor."""
    from moyolo_b200 import _lib, ops
    K = 256
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, generator=g)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    xd, wd, bd = x.to(dev), w.to(dev), b.to(dev)
    tol = float(ref.abs().max()) * 2 ** -8 + 1e-6
    y = ops.linear(xd, wd, bd, out_dtype=torch.bfloat16, engine=_lib.GEMM_TCGEN05)
    torch.cuda.synchronize()
    assert float((y.float().cpu().double() - ref).abs().max()) <= tol
    zr = torch.rand(M, generator=g) < 0.25
    wide = torch.full((M, N + 128), 7.0, dtype=torch.bfloat16, device=dev)
    ops.linear(xd, wd, bd, zero_rows=zr.to(dev, torch.uint8), out=wide[:, 64:64 + N], engine=_lib.GEMM_TCGEN05)
    got = wide[:, 64:64 + N].float().cpu()
    assert torch.equal(got[zr], torch.zeros(int(zr.sum()), N)) and torch.equal(got[~zr], y.float().cpu()[~zr])
    assert bool((wide[:, :64] == 7).all()) and bool((wide[:, 64 + N:] == 7).all())


@pytest.mark.parametrize("M", [1, 300, 382, 1408])
def test_linear_dual(dev, M):
    """q,k from x+pos and v from x in one launch == two separate GEMMs (bitwise: same tiles, same order)."""
    from moyolo_b200 import _lib, ops
    C = 256
    g = torch.Generator().manual_seed(M)
    x1 = torch.randn(M, C, generator=g).bfloat16().to(dev)
    x2 = torch.randn(M, C, generator=g).bfloat16().to(dev)
    w = (torch.randn(3 * C, C, generator=g) / 16).bfloat16().to(dev)
    b = torch.randn(3 * C, generator=g).to(dev)
    out = torch.empty(M, 3 * C, dtype=torch.bfloat16, device=dev)
    ops.linear_dual(x1, x2, 2 * C, w, b, out=out)
    ref = torch.cat([torch.nn.functional.linear(x1.double(), w[:2 * C].double(), b[:2 * C].double()),
                     torch.nn.functional.linear(x2.double(), w[2 * C:].double(), b[2 * C:].double())], 1).cpu()
    assert float((out.float().cpu().double() - ref).abs().max()) <= float(ref.abs().max()) * 2 ** -8 + 1e-6
    a = ops.linear(x1, w[:2 * C], b[:2 * C], out_dtype=torch.bfloat16, engine=_lib.GEMM_TCGEN05)
    c = ops.linear(x2, w[2 * C:], b[2 * C:], out_dtype=torch.bfloat16, engine=_lib.GEMM_TCGEN05)
    assert torch.equal(out[:, :2 * C], a) and torch.equal(out[:, 2 * C:], c)


@pytest.mark.parametrize("M,K", [(1, 256), (300, 256), (382, 1024), (1408, 256), (1000, 1024)])
def test_linear_add_layernorm(dev, M, K):
    """GEMM + residual + LayerNorm (+pos operand) in one launch (cluster/DSMEM row statistics) against an fp64
    reference on the same bf16 operands: fp32 output within 2e-5*rms (accumulation order), bf16 outputs within
    bf16 rounding; and against the unfused kernel pair within 2e-5*rms."""
    from moyolo_b200 import _lib, ops
    C = 256
    g = torch.Generator().manual_seed(M + K)
    x = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(C, K, generator=g) / K ** 0.5).bfloat16()
    b, res, pos = torch.randn(C, generator=g), torch.randn(M, C, generator=g) * 2 + 0.5, torch.randn(M, C, generator=g)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    t = torch.nn.functional.linear(x.double(), w.double(), b.double()) + res.double()
    ref = torch.nn.functional.layer_norm(t, (C,), gamma.double(), beta.double(), 1e-5)
    d = lambda z: z.to(dev)  # noqa: E731
    o32 = torch.empty(M, C, device=dev)
    olp = torch.empty(M, C, dtype=torch.bfloat16, device=dev)
    opos = torch.empty(M, C, dtype=torch.bfloat16, device=dev)
    ops.linear_add_layernorm(d(x), d(w), d(b), d(res), d(gamma), d(beta), 1e-5, out_f32=o32, out_lp=olp, pos=d(pos),
                             out_pos=opos)
    torch.cuda.synchronize()
    assert rel_rms(o32.cpu().numpy(), ref.numpy()) < 2e-5
    assert float((olp.float().cpu().double() - ref).abs().max()) <= float(ref.abs().max()) * 2 ** -8 + 1e-5
    rp = ref + pos.double()
    assert float((opos.float().cpu().double() - rp).abs().max()) <= float(rp.abs().max()) * 2 ** -8 + 1e-5
    tt = ops.linear(d(x), d(w), d(b), out_dtype=torch.float32, engine=_lib.GEMM_TCGEN05)
    u32, _, _ = ops.add_layernorm(tt, d(res), d(gamma), d(beta), 1e-5)
    assert rel_rms(o32.cpu().numpy(), u32.cpu().numpy()) < 2e-5
    # optional outputs / no residual
    o2 = torch.empty(M, C, device=dev)
    ops.linear_add_layernorm(d(x), d(w), None, None, d(gamma), d(beta), 1e-5, out_f32=o2)
    ref2 = torch.nn.functional.layer_norm(torch.nn.functional.linear(x.double(), w.double()), (C,), gamma.double(),
                                          beta.double(), 1e-5)
    assert rel_rms(o2.cpu().numpy(), ref2.numpy()) < 2e-5


@pytest.mark.parametrize("M,F,with_pos", [(382, 1024, True), (77, 1024, False), (300, 256, False), (129, 256, True),
                                          (3056, 1024, True)])
def test_ffn_add_layernorm_fused(dev, M, F, with_pos):
    """The one-launch FFN block (GEMM -> ReLU -> GEMM -> +residual -> LayerNorm in one cluster kernel) against the
    two launches it replaces (bit-identical: same operand rounding and K order) and against an fp64 reference of
    transformer.py:576-580 on the same bf16-rounded operands."""
    from moyolo_b200 import ops
    C = 256
    assert ops.ffn_fused_supported(torch.bfloat16, C, F) in (True, False)  # opt-in (MOYOLO_FFN_FUSED=1): measured slower
    g = torch.Generator().manual_seed(M + F)
    x = torch.randn(M, C, generator=g).bfloat16()
    w1 = (torch.randn(F, C, generator=g) / C ** 0.5).bfloat16()
    b1 = torch.randn(F, generator=g) * 0.1
    w2 = (torch.randn(C, F, generator=g) / F ** 0.5).bfloat16()
    b2 = torch.randn(C, generator=g) * 0.1
    res = torch.randn(M, C, generator=g)
    pos = torch.randn(M, C, generator=g)
    gam, bet = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    xd, w1d, b1d, w2d, b2d, resd, posd, gd, bd = (t.to(dev) for t in (x, w1, b1, w2, b2, res, pos, gam, bet))

    def outs():
        return (torch.zeros(M, C, device=dev), torch.zeros(M, C, dtype=torch.bfloat16, device=dev),
                torch.zeros(M, C, dtype=torch.bfloat16, device=dev) if with_pos else None)

    # two-launch path
    h2 = torch.zeros(M, F, dtype=torch.bfloat16, device=dev)
    a32, alp, apos = outs()
    ops.linear(xd, w1d, b1d, relu=True, out=h2)
    ops.linear_add_layernorm(h2, w2d, b2d, resd, gd, bd, 1e-5, out_f32=a32, out_lp=alp, pos=posd if with_pos else None,
                             out_pos=apos)
    # fused
    h1 = torch.zeros(M, F, dtype=torch.bfloat16, device=dev)
    f32, flp, fpos = outs()
    ops.ffn_add_layernorm(xd, w1d, b1d, w2d, b2d, h1, resd, gd, bd, 1e-5, out_f32=f32, out_lp=flp,
                          pos=posd if with_pos else None, out_pos=fpos)
    torch.cuda.synchronize()
    assert torch.equal(h1, h2), "hidden activations differ"
    assert torch.equal(f32, a32) and torch.equal(flp, alp)
    if with_pos:
        assert torch.equal(fpos, apos)
    # fp64 reference on the same operands (hidden rounded to bf16 as both device paths do)
    hid = torch.relu(x.double() @ w1.double().T + b1.double()).bfloat16().double()
    t = hid @ w2.double().T + b2.double() + res.double()
    ref = torch.nn.functional.layer_norm(t, (C,), gam.double(), bet.double(), 1e-5)
    assert rel_rms(f32.cpu().double().numpy(), ref.numpy()) < 5e-3  # bf16 rounding of the hidden layer dominates


@pytest.mark.parametrize("M,K,nc", [(382, 1024, 1), (77, 256, 5), (300, 1024, 8)])
def test_linear_add_layernorm_scores(dev, M, K, nc):
    """GEMM + residual + LayerNorm with the class-score head fused behind it (partials sent to cluster rank 0 with
    st.async) against the two launches it replaces: rows bit-identical, logits equal up to fp32 summation order,
    same labels, and an fp64 reference of the score head on the bf16-rounded rows."""
    from moyolo_b200 import ops
    C = 256
    g = torch.Generator().manual_seed(M + K + nc)
    x = torch.randn(M, K, generator=g).bfloat16().to(dev)
    w = (torch.randn(C, K, generator=g) / K ** 0.5).bfloat16().to(dev)
    b = (torch.randn(C, generator=g) * 0.1).to(dev)
    res = torch.randn(M, C, generator=g).to(dev)
    gam, bet = (torch.rand(C, generator=g) + 0.5).to(dev), (torch.randn(C, generator=g) * 0.1).to(dev)
    sw, sb = (torch.randn(nc, C, generator=g) * 0.3).to(dev), torch.randn(nc, generator=g).to(dev)
    a32, alp = torch.zeros(M, C, device=dev), torch.zeros(M, C, dtype=torch.bfloat16, device=dev)
    ops.linear_add_layernorm(x, w, b, res, gam, bet, 1e-5, out_f32=a32, out_lp=alp)
    lg_a, sc_a, lb_a = ops.score_head(alp, sw, sb)
    f32, flp = torch.zeros(M, C, device=dev), torch.zeros(M, C, dtype=torch.bfloat16, device=dev)
    lg = torch.zeros(M, nc, device=dev)
    sc = torch.zeros(M, device=dev)
    lb = torch.full((M,), -1, dtype=torch.int32, device=dev)
    ops.linear_add_layernorm_scores(x, w, b, res, gam, bet, 1e-5, sw, sb, out_f32=f32, out_lp=flp, logits=lg, scores=sc,
                                    labels=lb)
    torch.cuda.synchronize()
    assert torch.equal(f32, a32) and torch.equal(flp, alp)
    ref = flp.double().cpu() @ sw.double().cpu().T + sb.double().cpu()
    assert rel_rms(lg.cpu().double().numpy(), ref.numpy()) < 2e-5
    assert rel_rms(lg.cpu().numpy(), lg_a.cpu().numpy()) < 2e-5
    assert rel_rms(sc.cpu().numpy(), sc_a.cpu().numpy()) < 2e-5
    top2 = ref.topk(min(2, nc), dim=1).values
    clear = torch.ones(M, dtype=torch.bool) if nc == 1 else (top2[:, 0] - top2[:, 1]) > 1e-4
    assert torch.equal(lb.cpu()[clear].long(), ref.argmax(1)[clear]) and torch.equal(lb.cpu()[clear], lb_a.cpu()[clear])
