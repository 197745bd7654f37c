"""GPU parity tests: every CUDA entry point (called through the C ABI via moyolo_b200.ops / the drop-in
modules) against the golden vectors of the unmodified reference and against the CPU oracle.

Tolerances: fp32 max|a-b| <= 1e-4 * rms(ref) per tensor; bf16 (value + GEMM operands
bf16; locations, softmax, accumulation, LayerNorm fp32) max|a-b| <= 2e-2 * rms(ref) per MSDeformAttn
call and <= 5e-3 absolute on boxes after 6 layers. Integer/ID work: bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_rms

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2         # gather alone: bf16 value, everything else fp32 (SURVEY.md §8(c) measured 1.5e-2)
BF16_MODULE_TOL = 3e-2  # whole MSDeformAttn / decoder layer: bf16 value, GEMM operands AND gather output


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from moyolo_b200 import _lib
    assert _lib.lib().moyolo_device_supported() == 1, "libmoyolo_b200 targets sm_100a (B200) only"
    return torch.device("cuda:0")


def _mods():
    import moyolo_b200 as m
    from moyolo_b200 import ops, synthetic as syn
    from oracle import make_golden as mg
    from oracle import torch_port as tp
    return m, ops, syn, mg, tp


# ------------------------------------------------------------------ a1: the gather core
def test_kat0_legacy_ffi(dev):
    """The reference's own known-answer case (MOTR/models/ops/test.py:21-60) through the stand-in
    `MultiScaleDeformableAttention.ms_deform_attn_forward`, double and float."""
    from moyolo_b200 import msda_ext
    meta, g = load_golden("kat0")
    shapes = torch.as_tensor(meta["shapes"], dtype=torch.long, device=dev)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    for tag, dt in (("double", torch.float64), ("float", torch.float32)):
        v, loc, aw = (torch.from_numpy(g[f"{k}_{tag}"]).to(dev, dt) for k in ("value", "loc", "aw"))
        out = msda_ext.ms_deform_attn_forward(v, shapes, lsi, loc, aw, 2).cpu().numpy()
        ref = g[f"out_{tag}"]
        if tag == "double":
            assert np.allclose(out, ref)
        else:
            assert np.allclose(out, ref, rtol=1e-2, atol=1e-3)
            assert rel_rms(out, ref) < FP32_TOL
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        msda_ext.ms_deform_attn_forward(v.cpu(), shapes, lsi, loc, aw, 2)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        msda_ext.ms_deform_attn_backward(v.cpu(), shapes, lsi, loc, aw, torch.from_numpy(out).to(dev), 2)


def test_core_golden(dev):
    m, ops, syn, mg, tp = _mods()
    for case in mg.CORE_CASES:
        meta, g = load_golden(case["name"])
        value, loc, w = syn.make_core_inputs(case["seed"], case["B"], case["Q"], case["H"], case["D"], case["shapes"],
                                             case["P"])
        out = ops.msda_sampled(value.to(dev), meta["shapes"], loc.to(dev), w.to(dev)).cpu().numpy()
        assert rel_rms(out, g["out_f32"]) < FP32_TOL, case["name"]
        assert rel_rms(out, g["out_f64"]) < FP32_TOL, case["name"]
        out_fn = m.multi_scale_deformable_attn(value.to(dev), meta["shapes"], loc.to(dev), w.to(dev))
        assert torch.equal(out_fn.cpu(), torch.from_numpy(out))
        # bf16 value, fp32 locations/weights/accumulation
        ob = ops.msda_sampled(value.to(dev, torch.bfloat16), meta["shapes"], loc.to(dev), w.to(dev)).float().cpu()
        assert rel_rms(ob.numpy(), g["out_f64"]) < BF16_TOL, case["name"]


@pytest.mark.parametrize("B,Q,H,D,P,name", [(1, 300, 8, 32, 4, "C1"), (2, 357, 8, 32, 4, "MOT17"), (1, 64, 8, 32, 8, "tiny"),
                                            (1, 100, 4, 64, 4, "KITTI"), (3, 17, 8, 32, 4, "tiny")])
def test_core_vs_c_oracle(dev, B, Q, H, D, P, name):
    """Named-config shapes against the C restatement (oracle/msda_core.c), fp32 and bf16 value."""
    m, ops, syn, mg, tp = _mods()
    from oracle import c_core
    shapes = [list(s) for s in syn.PYRAMIDS[name]]
    value, loc, w = syn.make_core_inputs(100 + Q, B, Q, H, D, shapes, P)
    ref = c_core.msda_core(value.double().numpy(), shapes, loc.double().numpy(), w.double().numpy())
    out = ops.msda_sampled(value.to(dev), shapes, loc.to(dev), w.to(dev)).cpu().numpy()
    assert rel_rms(out, ref) < FP32_TOL
    ob = ops.msda_sampled(value.to(dev, torch.bfloat16), shapes, loc.to(dev), w.to(dev)).float().cpu().numpy()
    assert rel_rms(ob, ref) < BF16_TOL


def test_core_properties_full_size(dev):
    """Size-independent properties at the DanceTrack shape with B=4, Q=500: linearity in value,
    partition of unity (value == 1 inside the image -> output == sum of in-range weights), and
    independence of the batch rows (ragged row_offsets == dense)."""
    m, ops, syn, mg, tp = _mods()
    shapes = [list(s) for s in syn.PYRAMIDS["DanceTrack"]]
    B, Q, H, D, P = 4, 500, 8, 32, 4
    value, loc, w = syn.make_core_inputs(7, B, Q, H, D, shapes, P, outside_frac=0.0)
    loc = loc.clamp(0.05, 0.95)
    v, l, ww = value.to(dev), loc.to(dev), w.to(dev)
    o1 = ops.msda_sampled(v, shapes, l, ww)
    o2 = ops.msda_sampled(2.5 * v, shapes, l, ww)
    assert rel_rms((o2 / 2.5).cpu().numpy(), o1.cpu().numpy()) < 1e-5
    ones = torch.ones_like(v)
    o3 = ops.msda_sampled(ones, shapes, l, ww)
    assert torch.allclose(o3, torch.ones_like(o3), atol=1e-5)
    ro = torch.tensor([0, Q, 2 * Q, 3 * Q, 4 * Q], dtype=torch.int32, device=dev)
    o4 = ops.msda_sampled(v, shapes, l, ww, row_offsets=ro)
    assert torch.equal(o4, o1)
    # empty query set
    e = ops.msda_sampled(v, shapes, l[:, :0], ww[:, :0])
    assert e.shape == (B, 0, H * D)


# ------------------------------------------------------------------ small ops
def test_linear_simt(dev):
    m, ops, syn, mg, tp = _mods()
    from moyolo_b200 import _lib
    g = torch.Generator().manual_seed(1)
    for M, N, K in ((300, 288, 256), (77, 1024, 256), (65, 4, 19), (1, 256, 1024), (1000, 96, 64)):
        x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
        ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
        y = ops.linear(x.to(dev), w.to(dev), b.to(dev), engine=_lib.GEMM_SIMT).cpu()
        assert rel_rms(y.numpy(), ref.numpy()) < 1e-5
        yr = ops.linear(x.to(dev), w.to(dev), b.to(dev), relu=True, engine=_lib.GEMM_SIMT).cpu()
        assert rel_rms(yr.numpy(), ref.relu().numpy()) < 1e-5
        zr = (torch.rand(M, generator=g) < 0.3)
        yz = ops.linear(x.to(dev), w.to(dev), b.to(dev), zero_rows=zr.to(dev, torch.uint8), engine=_lib.GEMM_SIMT).cpu()
        assert torch.equal(yz[zr], torch.zeros_like(yz[zr])) and torch.equal(yz[~zr], y[~zr])
        xb, wb = x.bfloat16(), w.bfloat16()
        refb = torch.nn.functional.linear(xb.double(), wb.double(), b.double())
        yb = ops.linear(xb.to(dev), wb.to(dev), b.to(dev), out_dtype=torch.float32, engine=_lib.GEMM_SIMT).cpu()
        assert rel_rms(yb.numpy(), refb.numpy()) < 1e-5


def test_posemb_sigmoid_ops(dev):
    m, ops, syn, mg, tp = _mods()
    _, g = load_golden("posemb")
    emb = ops.pos2posemb(torch.from_numpy(g["pos"]).to(dev)).cpu().numpy()
    assert rel_rms(emb, g["emb"]) < FP32_TOL
    inv = ops.inverse_sigmoid(torch.from_numpy(g["x"]).to(dev)).cpu().numpy()
    assert np.allclose(inv, g["inv"], rtol=1e-5, atol=1e-5)
    x = torch.randn(1000) * 4
    assert torch.allclose(ops.sigmoid(x.to(dev)).cpu(), x.sigmoid(), atol=1e-6)


def test_self_attention(dev):
    m, ops, syn, mg, tp = _mods()
    g = torch.Generator().manual_seed(3)
    C, H = 256, 8
    lens = [300, 1, 77, 357]
    offs = [0]
    for n in lens:
        offs.append(offs[-1] + n)
    R = offs[-1]
    q, k, v = (torch.randn(R, C, generator=g) for _ in range(3))
    ro = torch.tensor(offs, dtype=torch.int32, device=dev)
    out = ops.self_attention(q.to(dev), k.to(dev), v.to(dev), ro, offs, H).cpu()
    for b, n in enumerate(lens):
        s = slice(offs[b], offs[b + 1])
        qq, kk, vv = (t[s].double().view(n, H, C // H).transpose(0, 1) for t in (q, k, v))
        ref = (torch.softmax(qq @ kk.transpose(1, 2) / (C // H) ** 0.5, -1) @ vv).transpose(0, 1).reshape(n, C)
        assert rel_rms(out[s].numpy(), ref.numpy()) < 1e-5
    ob = ops.self_attention(q.to(dev, torch.bfloat16), k.to(dev, torch.bfloat16), v.to(dev, torch.bfloat16), ro, offs,
                            H).float().cpu()
    assert rel_rms(ob.numpy(), out.numpy()) < 0.1  # max-norm over rms with bf16 inputs AND bf16 output rounding


def test_add_layernorm_heads(dev):
    m, ops, syn, mg, tp = _mods()
    g = torch.Generator().manual_seed(4)
    R, C = 333, 256
    x, r, pos = (torch.randn(R, C, generator=g) for _ in range(3))
    ga, be = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ref = torch.nn.functional.layer_norm((x + r).double(), (C,), ga.double(), be.double(), 1e-5)
    f32, lp, plp = ops.add_layernorm(x.to(dev), r.to(dev), ga.to(dev), be.to(dev), 1e-5, True, True, torch.bfloat16,
                                     pos.to(dev))
    assert rel_rms(f32.cpu().numpy(), ref.numpy()) < 1e-5
    # bf16 copies: half an ulp of the largest element (|x| < 16 -> 2^-5) over the rms
    assert float((lp.float().cpu() - ref).abs().max()) < 2 ** -5
    assert float((plp.float().cpu() - (ref + pos)).abs().max()) < 2 ** -5
    # box refine + score head vs torch
    h = torch.randn(R, C, generator=g)
    w3, b3 = torch.randn(4, C, generator=g) * 0.05, torch.randn(4, generator=g) * 0.1
    refb = torch.rand(R, 4, generator=g)
    want = torch.sigmoid(torch.nn.functional.linear(h, w3, b3) + tp.inverse_sigmoid(refb))
    got = ops.box_refine(h.to(dev), w3.to(dev), b3.to(dev), refb.to(dev)).cpu()
    assert torch.allclose(got, want, atol=2e-6)
    ws, bs = torch.randn(5, C, generator=g) * 0.1, torch.randn(5, generator=g)
    logits, scores, labels = ops.score_head(h.to(dev), ws.to(dev), bs.to(dev))
    wl = torch.nn.functional.linear(h, ws, bs)
    assert torch.allclose(logits.cpu(), wl, atol=1e-4)
    assert torch.allclose(scores.cpu(), wl.sigmoid().max(-1).values, atol=1e-5)
    assert torch.equal(labels.cpu().long(), wl.argmax(-1))


# ------------------------------------------------------------------ a2-a5: modules vs reference goldens
def _load_layer(m, syn, sd, prefix, layer):
    layer.load_state_dict(syn.sub_state(sd, prefix))
    return layer.eval()


@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_MODULE_TOL)])
def test_msdeform_attn_module(dev, precision, tol):
    m, ops, syn, mg, tp = _mods()
    spec = syn.DecoderSpec()
    for case in mg.MSDA_CASES:
        meta, g = load_golden(case["name"])
        sd = syn.make_decoder_state(spec, meta["weight_seed"])
        mod = m.MSDeformAttn(spec.d_model, spec.n_levels, spec.n_heads, spec.n_points)
        mod.load_state_dict(syn.sub_state(sd, "layers.0.cross_attn."))
        mod = mod.to(dev).eval()
        mod.precision = precision
        q, refer, feats, _ = syn.make_module_inputs(case["seed"], case["B"], case["Q"], spec.d_model, case["shapes"],
                                                    case["ref_dim"], case["ref_levels"])
        mask = None
        if case["mask"]:
            mask = (torch.rand(case["B"], feats.shape[1], generator=torch.Generator().manual_seed(case["seed"])) < 0.2).to(dev)
        out = mod(q.to(dev), refer.to(dev), feats.to(dev), meta["shapes"], mask)
        assert out.dtype == torch.float32 and out.shape == g["out"].shape
        assert rel_rms(out.cpu().numpy(), g["out"]) < tol, case["name"]


@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_MODULE_TOL)])
def test_decoder_layers(dev, precision, tol):
    m, ops, syn, mg, tp = _mods()
    spec = syn.DecoderSpec()
    for case in mg.LAYER_CASES:
        meta, g = load_golden(case["name"])
        sd = syn.make_decoder_state(spec, meta["weight_seed"])
        cls = getattr(m, case["cls"])
        layer = cls(spec.d_model, spec.n_heads, spec.d_ffn, 0.0, torch.nn.ReLU(), spec.n_levels, spec.n_points)
        layer = _load_layer(m, syn, sd, "layers.1.", layer).to(dev)
        layer.precision = precision
        q, refer, feats, qpos = syn.make_module_inputs(case["seed"], case["B"], case["Q"], spec.d_model, case["shapes"], 4, 1)
        out = layer(q.to(dev), refer[:, :, 0].to(dev), feats.to(dev), meta["shapes"], None, None, qpos.to(dev))
        assert rel_rms(out.cpu().numpy(), g["out"]) < tol, case["name"]


def _build_decoder(m, syn, spec, sd, mode, dev, precision):
    layer_cls = m.MOTRDecoderLayer if mode == "motr" else m.DeformableTransformerDecoderLayer
    layer = layer_cls(spec.d_model, spec.n_heads, spec.d_ffn, 0.0, torch.nn.ReLU(), spec.n_levels, spec.n_points)
    dec = (m.MOTRTransformerDecoder if mode == "motr" else m.DeformableTransformerDecoder)(spec.d_model, layer,
                                                                                          spec.n_layers)
    dec.load_state_dict({k: v for k, v in sd.items() if k.startswith("layers.")})
    bbox = torch.nn.ModuleList([m.MLP(spec.d_model, spec.d_model, 4, 3) for _ in range(spec.n_layers)])
    bbox.load_state_dict(syn.sub_state(sd, "dec_bbox_head."))
    score = torch.nn.ModuleList([torch.nn.Linear(spec.d_model, spec.nc) for _ in range(spec.n_layers)])
    score.load_state_dict(syn.sub_state(sd, "dec_score_head."))
    pos = m.MLP(4, spec.pos_hidden, spec.d_model, 2)
    pos.load_state_dict(syn.sub_state(sd, "query_pos_head."))
    dec.precision = precision
    return dec.to(dev).eval(), bbox.to(dev).eval(), score.to(dev).eval(), pos.to(dev).eval()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_decoders_vs_reference_golden(dev, precision):
    """6-layer decoders (both classes) incl. the 640x640 / 300-query configuration C1."""
    m, ops, syn, mg, tp = _mods()
    for case in mg.DECODER_CASES:
        meta, g = load_golden(case["name"])
        spec = syn.DecoderSpec(nc=case["nc"])
        sd = syn.make_decoder_state(spec, meta["weight_seed"])
        dec, bbox, score, pos = _build_decoder(m, syn, spec, sd, case["mode"], dev, precision)
        embed, refer, feats, qpos = syn.make_decoder_inputs(case["seed"], case["B"], case["Q"], spec.d_model, case["shapes"])
        args = (embed.to(dev), refer.to(dev), feats.to(dev), meta["shapes"], bbox, score, pos)
        if case["mode"] == "motr":
            b, s, hs = dec(*args, track_query_embed=qpos.to(dev))
        else:
            b, s = dec(*args)
            hs = None
        assert b.shape == g["boxes"].shape and s.shape == g["scores"].shape
        if precision == "fp32":
            assert rel_rms(b.cpu().numpy(), g["boxes"]) < FP32_TOL, case["name"]
            assert rel_rms(s.cpu().numpy(), g["scores"]) < FP32_TOL, case["name"]
            if hs is not None:
                # 256-wide output embedding after 6 layers on white-noise feature maps: max-norm over 76 800 values;
                # the fp32 reference's own deviation from an fp64 evaluation is of this size, so two
                # different fp32 evaluation orders (CUDA-core or split-bf16 tensor-core linears) can differ by it
                assert rel_rms(hs.cpu().numpy(), g["hs"]) < 2 * FP32_TOL, case["name"]
        else:
            assert float(np.abs(b.cpu().numpy() - g["boxes"]).max()) < 5e-3, case["name"]  # normalised coords
            assert rel_rms(s.cpu().numpy(), g["scores"]) < 5e-2, case["name"]


# ------------------------------------------------------------------ a8: tracker kernels, bit exact
@pytest.mark.parametrize("name", ["tracker_a", "tracker_b", "tracker_empty"])
def test_track_assign_bit_exact(dev, name):
    m, ops, syn, mg, tp = _mods()
    meta, g = load_golden(name)
    counters = torch.zeros(2, dtype=torch.int64, device=dev)
    for t in range(meta["n_frames"]):
        assert counters.tolist() == [int(x) for x in g[f"counters_in_{t}"]]
        ids = torch.from_numpy(g[f"ids_in_{t}"]).to(dev)
        dis = torch.from_numpy(g[f"dis_in_{t}"]).to(dev)
        n = ids.shape[0]
        ws = torch.empty(ops.track_workspace_bytes(n), dtype=torch.uint8, device=dev)
        ops.track_assign(torch.from_numpy(g[f"scores_{t}"]).to(dev), torch.from_numpy(g[f"boxes_{t}"]).to(dev), ids,
                         dis, counters, ws)
        assert np.array_equal(ids.cpu().numpy(), g[f"ids_out_{t}"]), f"frame {t}"
        assert np.array_equal(dis.cpu().numpy(), g[f"dis_out_{t}"]), f"frame {t}"
        assert counters.tolist() == [int(x) for x in g[f"counters_out_{t}"]], f"frame {t}"
        # compaction == boolean indexing
        act = g[f"ids_out_{t}"] >= 0
        boxes = torch.from_numpy(g[f"boxes_{t}"]).to(dev)
        out_b, out_i = torch.zeros_like(boxes), torch.zeros_like(ids)
        n_act = torch.zeros(1, dtype=torch.int32, device=dev)
        idx = torch.zeros(n, dtype=torch.int32, device=dev)
        ops.track_compact(ids, [boxes, ids], [out_b, out_i], n_act, idx)
        k = int(n_act.item())
        assert k == int(act.sum())
        assert np.array_equal(out_b[:k].cpu().numpy(), g[f"boxes_{t}"][act])
        assert np.array_equal(out_i[:k].cpu().numpy(), g[f"ids_out_{t}"][act])


# ------------------------------------------------------------------ a7-a9: carried-track frame loop (O3)
def _margin_ok(rec, margin):
    s = rec["scores"]
    return not (np.any(np.abs(s - 0.4) < margin) or np.any(np.abs(s - 0.5) < margin))


def _calibrated_state(syn, tp, spec, sd, frame0, shapes):
    feats, de, dr = frame0
    with torch.no_grad():
        _, s, _ = tp.decoder_forward(sd, de[None], dr[None], feats[None], shapes, spec.n_heads, spec.n_levels,
                                     spec.n_points, spec.n_layers, "motr", tp.pos2posemb(dr)[None])
    return syn.calibrate_score_bias(sd, s[0, 0], spec, 0.08)


@pytest.mark.parametrize("precision,margin,box_tol,graphs", [("fp32", 2e-5, 1e-4, True), ("bf16", 2e-2, 5e-3, True),
                                                             ("fp32", 2e-5, 1e-4, False)])
def test_track_sequence_vs_oracle(dev, precision, margin, box_tol, graphs):
    """Track-ID assignment on synthetic sequences: two lock-step sequences on the GPU vs two
    independent CPU oracle (O3) runs.
    (1) free-running: IDs must be identical and boxes within tolerance on every frame until an oracle
        score comes within `margin` of the 0.4 / 0.5 thresholds (threshold-margin rule, SURVEY.md
        §8(c)); after that the two trajectories may legitimately diverge and the free-running
        comparison of that sequence stops. fp32 must survive >= 6 of 8 frames per the seeds used.
    (2) teacher-forced, every frame and both precisions: the oracle's ID assigner applied to the GPU's
        own scores/boxes/previous IDs must reproduce the GPU IDs, disappear counters and ID counters
        bit-exactly (integer logic has no tolerance)."""
    m, ops, syn, mg, tp = _mods()
    from moyolo_b200.tracker import TrackEngine
    from oracle.tracker_port import TrackerPort, track_sequence_port
    spec = syn.DecoderSpec()
    sd = syn.make_decoder_state(spec, 7)
    shapes = [list(s) for s in syn.PYRAMIDS["tiny"]]
    n_frames, nd, S = 8, 64, 2
    gens = [syn.SequenceGenerator(syn.SequenceSpec(name="tiny", n_frames=n_frames, n_detect=nd, seed=s, shapes=shapes),
                                  spec.d_model) for s in range(S)]
    frames = [[tuple(t.clone() for t in g.next_frame()) for _ in range(n_frames)] for g in gens]
    sd = _calibrated_state(syn, tp, spec, sd, frames[0][0], shapes)
    refs = [track_sequence_port(sd, frames[s], shapes, spec.n_heads, spec.n_levels, spec.n_points, spec.n_layers,
                                spec.nc) for s in range(S)]
    eng = TrackEngine(sd, spec, shapes, dev, precision, nd, S, use_graphs=graphs)
    alive = [True] * S
    compared = 0
    forced = [TrackerPort() for _ in range(S)]
    prev = [(np.zeros(0, np.int64), np.zeros(0, np.int64)) for _ in range(S)]
    for t in range(n_frames):
        feats = torch.stack([frames[s][t][0] for s in range(S)]).to(dev)
        de = torch.stack([frames[s][t][1] for s in range(S)]).to(dev)
        dr = torch.stack([frames[s][t][2] for s in range(S)]).to(dev)
        outs = eng.step(feats, de, dr)
        for s in range(S):
            ids_gpu = outs[s]["ids"].cpu().numpy()
            # (2) teacher-forced integer logic
            ids = np.concatenate([prev[s][0], np.full(nd, -1, np.int64)])
            dis = np.concatenate([prev[s][1], np.zeros(nd, np.int64)])
            forced[s].update(outs[s]["scores"].cpu().numpy(), outs[s]["boxes"].cpu().numpy(), ids, dis)
            assert np.array_equal(ids_gpu, ids), (s, t, "teacher-forced ids")
            assert eng.counters[s].tolist() == [forced[s].max_obj_id, forced[s].max_obj_id_pre], (s, t)
            act = ids >= 0
            prev[s] = (ids[act], dis[act])
            assert np.array_equal(eng.track_ids(s).cpu().numpy(), ids[act]) and \
                np.array_equal(eng.track_disappear(s).cpu().numpy(), dis[act]), (s, t, "carried state")
            # (1) free-running vs O3
            if not alive[s]:
                continue
            r = refs[s][t]
            if not _margin_ok(r, margin):
                alive[s] = False
                continue
            assert np.array_equal(ids_gpu, r["ids"]), (s, t, "free-running ids")
            scale = 1.0 if precision == "bf16" else float(np.sqrt((r["boxes"] ** 2).mean()))
            assert float(np.abs(outs[s]["boxes"].cpu().numpy() - r["boxes"]).max()) < box_tol * scale, (s, t)
            assert eng.counters[s].tolist() == list(r["counters"]), (s, t)
            compared += 1
    print(f"[{precision}] free-running frames compared: {compared} of {S * n_frames}")
    if precision == "fp32":
        assert compared >= 10, f"margin rule excluded too many frames ({compared} compared)"
    assert max(rr["n_tracks_in"] for rr in refs[0]) > 0, "sequence never carried a track"


def test_core_vs_reference_cuda_kernel(dev):
    """Our gather against the REFERENCE's own CUDA kernel (MOTR/models/ops/src/cuda/
    ms_deform_im2col_cuda.cuh) compiled for sm_100a into oracle/_ref by oracle/Makefile."""
    from oracle import ref_cuda
    if not ref_cuda.available():
        pytest.skip("oracle/_ref/libmsda_ref_cuda.so not built (needs the reference tree at build time)")
    m, ops, syn, mg, tp = _mods()
    for name, B, Q in (("tiny", 2, 50), ("C1", 1, 300), ("MOT17", 2, 357)):
        shapes = [list(s) for s in syn.PYRAMIDS[name]]
        value, loc, w = syn.make_core_inputs(3 + Q, B, Q, 8, 32, shapes, 4)
        v, l, ww = value.to(dev), loc.to(dev), w.to(dev)
        ref = ref_cuda.msda_im2col(v, shapes, l, ww)
        out = ops.msda_sampled(v, shapes, l, ww)
        assert rel_rms(out.cpu().numpy(), ref.cpu().numpy()) < 1e-5, name


def test_pipelined_submit_equals_synchronous_steps(dev):
    """The host-ahead pipeline (submit/collect, speculative padded sizes, device-side abort + re-launch)
    must produce exactly the rows of the synchronous step() loop: bit-identical IDs AND boxes/scores
    (same kernels, same inputs; padding rows never influence real rows). margin=0 with an 8-row bucket
    forces mis-speculation whenever the track count crosses a bucket boundary, so the abort path runs."""
    m, ops, syn, mg, tp = _mods()
    from moyolo_b200.tracker import DecoderWeights, TrackEngine
    spec = syn.DecoderSpec()
    sd = syn.make_decoder_state(spec, 7)
    shapes = [list(s) for s in syn.PYRAMIDS["tiny"]]
    n_frames, nd, S = 10, 64, 2
    gens = [syn.SequenceGenerator(syn.SequenceSpec(name="tiny", n_frames=n_frames, n_detect=nd, seed=s, shapes=shapes),
                                  spec.d_model) for s in range(S)]
    frames = [[tuple(t.clone() for t in g.next_frame()) for _ in range(n_frames)] for g in gens]
    sd = _calibrated_state(syn, tp, spec, sd, frames[0][0], shapes)
    W = DecoderWeights(sd, spec, dev, "bf16")
    batches = [tuple(torch.stack([frames[s][t][k] for s in range(S)]).to(dev) for k in range(3)) for t in range(n_frames)]
    torch.cuda.synchronize()

    ref = TrackEngine(sd, spec, shapes, dev, "bf16", nd, S, weights=W)
    ref_rows = []
    for t in range(n_frames):
        outs = ref.step(*batches[t])
        ref_rows.append([{k: v.clone().cpu() for k, v in o.items()} for o in outs])
    ref_table = ref.track_table().clone().cpu()
    assert ref_table.shape[0] > 0 and max(ref.n_tracks_host()) > 0

    for margin, bucket in ((0, 8), (32, 64)):
        eng = TrackEngine(sd, spec, shapes, dev, "bf16", nd, S, weights=W, margin=margin, bucket=bucket)
        eng.set_seq_ids([5, 9])
        got = {}
        for t in range(n_frames):
            eng.submit(*batches[t], want_rows=True)
            if t > 0:
                got[t - 1] = [{k: v.clone() for k, v in o.items()} for o in eng.collect(t - 1)]
        got[n_frames - 1] = [{k: v.clone() for k, v in o.items()} for o in eng.collect(n_frames - 1)]
        table = eng.track_table().clone().cpu()
        if margin == 0:
            assert eng.aborts > 0, "the abort / re-launch path was not exercised"
        for t in range(n_frames):
            for s in range(S):
                for k in ("ids", "boxes", "scores"):
                    assert torch.equal(got[t][s][k], ref_rows[t][s][k].to(got[t][s][k].dtype)), (margin, t, s, k)
                assert torch.equal(got[t][s]["labels"], ref_rows[t][s]["labels"]), (margin, t, s)
        # the device-resident table: same rows, slot ids remapped to the global sequence ids
        want = ref_table.clone()
        want[:, 0] = torch.where(ref_table[:, 0] == 0, torch.tensor(5.0), torch.tensor(9.0))
        assert torch.equal(table, want), margin
        assert ref.n_tracks_host() == eng.n_tracks_host()

    # host running up to 6 frames ahead (host_lag, what sharding.run_sharded / bench.py use): nothing is collected
    # in the loop, so the queue really gets that deep; with margin 0 the re-run path has to reload the inputs of
    # aborted frames whose input slots were overwritten by later frames
    for margin, bucket in ((0, 8), (32, 64)):
        eng = TrackEngine(sd, spec, shapes, dev, "bf16", nd, S, weights=W, margin=margin, bucket=bucket, host_lag=6)
        eng.set_seq_ids([5, 9])
        for t in range(n_frames):
            eng.submit(*batches[t], want_rows=False)
        table = eng.track_table().clone().cpu()
        if margin == 0:
            assert eng.aborts > 0, "the deep-queue re-run path was not exercised"
        assert torch.equal(table, want), ("host_lag 6", margin)
        assert ref.n_tracks_host() == eng.n_tracks_host()


def test_frame_assign_compact_equals_two_step(dev):
    """moyolo_frame_assign_compact + moyolo_track_suppress_batched (the frame path: ID assignment fused with the
    compaction, counters updated off the critical path) against moyolo_track_assign_batched +
    moyolo_frame_compact, which test_track_assign_bit_exact pins to the reference: every output bit-identical,
    on ragged lock-step sequences including one with no active row and one that overflows `cap`."""
    m, ops, syn, mg, tp = _mods()
    g = torch.Generator().manual_seed(11)
    C, cap, S = 256, 48, 4
    ns = [37, 300, 1, 420]
    R = 832
    ro_host = [0]
    for n in ns:
        ro_host.append(ro_host[-1] + n)
    ro = torch.tensor(ro_host, dtype=torch.int32, device=dev)
    scores = torch.rand(R, generator=g)
    scores[ro_host[2]:ro_host[3]] = 0.1                       # sequence 2: nothing becomes active
    ids0 = torch.full((R,), -1, dtype=torch.int64)
    dis0 = torch.zeros(R, dtype=torch.int64)
    for s, n in enumerate(ns):                                # a carried prefix with ids and disappear counters
        k = n // 3 if s != 2 else 0
        ids0[ro_host[s]:ro_host[s] + k] = torch.randperm(100, generator=g)[:k] if k <= 100 else torch.arange(k)
        dis0[ro_host[s]:ro_host[s] + k] = torch.randint(0, 6, (k,), generator=g)
    boxes = torch.rand(R, 4, generator=g) * 0.5
    boxes[5] = boxes[4]                                       # exact duplicates exercise the IoU filter
    boxes[ro_host[1] + 7] = boxes[ro_host[1] + 3]
    labels = torch.randint(0, 5, (R,), generator=g, dtype=torch.int32)
    refer_logit, pos, hs = torch.randn(R, 4, generator=g), torch.randn(R, C, generator=g), torch.randn(R, C, generator=g)
    counters0 = torch.tensor([[100, 99], [7, 6], [0, 0], [1000, 999]], dtype=torch.int64)
    to = lambda *ts: [t.to(dev) for t in ts]  # noqa: E731
    scores, boxes, labels, refer_logit, pos, hs = to(scores, boxes, labels, refer_logit, pos, hs)
    z = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype, device=dev)  # noqa: E731

    def outputs():
        return dict(n_active=z(S, dtype=torch.int32), active_index=z(R, dtype=torch.int32), c_ref=z(R, 4), c_pos=z(R, C),
                    c_hs=z(R, C), c_box=z(R, 4), t_label=z(S, cap, dtype=torch.int32), t_ids=z(S, cap, dtype=torch.int64),
                    t_dis=z(S, cap, dtype=torch.int64), q_qk=z(R, C, dtype=torch.bfloat16), q_tgt=z(R, C, dtype=torch.bfloat16))

    aws = torch.zeros(S * ops.track_workspace_bytes(512), dtype=torch.uint8, device=dev)
    # two-step path (in place)
    a = outputs()
    ids_a, dis_a, cnt_a = ids0.to(dev), dis0.to(dev), counters0.to(dev)
    ops.track_assign_batched(scores, boxes, ids_a, dis_a, cnt_a, ro, S, 512, aws)
    ops.frame_compact(S, C, cap, ro, ids_a, dis_a, labels, refer_logit, pos, hs, boxes, a["n_active"], a["active_index"],
                      a["c_ref"], a["c_pos"], a["c_hs"], a["c_box"], a["t_label"], a["t_ids"], a["t_dis"],
                      q_qk_lp=a["q_qk"], q_tgt_lp=a["q_tgt"])
    # fused path
    b = outputs()
    ids_in, dis_in, cnt_b = ids0.to(dev), dis0.to(dev), counters0.to(dev)
    ids_b, dis_b = z(R, dtype=torch.int64), z(R, dtype=torch.int64)
    ops.frame_assign_compact(S, C, cap, R, ro, scores, ids_in, dis_in, cnt_b, ids_b, dis_b, labels, refer_logit, pos, hs,
                             boxes, b["n_active"], b["active_index"], b["c_ref"], b["c_pos"], b["c_hs"], b["c_box"],
                             b["t_label"], b["t_ids"], b["t_dis"], q_qk_lp=b["q_qk"], q_tgt_lp=b["q_tgt"])
    assert torch.equal(cnt_b, counters0.to(dev)), "frame_assign_compact must not touch the counters"
    ops.track_suppress_batched(boxes, ids_b, cnt_b, ro, S, 512, aws)
    torch.cuda.synchronize()
    n_rows = ro_host[-1]
    assert torch.equal(ids_a[:n_rows], ids_b[:n_rows]) and torch.equal(dis_a[:n_rows], dis_b[:n_rows])
    assert bool((ids_b[n_rows:] == -1).all()) and bool((dis_b[n_rows:] == 0).all())
    assert torch.equal(cnt_a, cnt_b), (cnt_a.tolist(), cnt_b.tolist())
    na = a["n_active"].tolist()
    assert na == b["n_active"].tolist() and na[2] == 0 and na[3] == cap, na
    for s in range(S):
        lo, k = ro_host[s], na[s]
        for key in ("active_index", "c_ref", "c_pos", "c_hs", "c_box", "q_qk", "q_tgt"):
            assert torch.equal(a[key][lo:lo + k], b[key][lo:lo + k]), (s, key)
        for key in ("t_label", "t_ids", "t_dis"):
            assert torch.equal(a[key][s, :k], b[key][s, :k]), (s, key)
    # and the integer logic against the numpy restatement of RuntimeTrackerBase.update
    from oracle.tracker_port import TrackerPort
    for s in range(S):
        port = TrackerPort()
        port.max_obj_id, port.max_obj_id_pre = int(counters0[s, 0]), int(counters0[s, 1])
        lo, hi = ro_host[s], ro_host[s + 1]
        ids_p, dis_p = ids0[lo:hi].numpy().copy(), dis0[lo:hi].numpy().copy()
        port.update(scores[lo:hi].cpu().numpy(), boxes[lo:hi].cpu().numpy(), ids_p, dis_p)
        assert np.array_equal(ids_p, ids_b[lo:hi].cpu().numpy()) and np.array_equal(dis_p, dis_b[lo:hi].cpu().numpy()), s
        assert [port.max_obj_id, port.max_obj_id_pre] == cnt_b[s].tolist(), s


# ------------------------------------------------------------------ f4: backward of the gather
def _legacy_shapes(shapes, dev):
    t = torch.as_tensor([list(s) for s in shapes], dtype=torch.long, device=dev)
    return t, torch.cat((t.new_zeros((1,)), t.prod(1).cumsum(0)[:-1]))


def test_backward_vs_reference_autograd_golden(dev):
    """ms_deform_attn_backward through the legacy FFI stand-in against torch.autograd of the reference's
    multi_scale_deformable_attn_pytorch (tests/golden/core_grad_*.npz): fp32 within 1e-4*rms, fp64 within 1e-10;
    D=32 / D=64 run the warp kernel, D=6 and fp64 the generic one."""
    m, ops, syn, mg, tp = _mods()
    from moyolo_b200 import msda_ext
    for c in mg.GRAD_CASES:
        meta, g = load_golden(c["name"])
        value, loc, w, go = syn.make_core_grad_inputs(c["seed"], c["B"], c["Q"], c["H"], c["D"], c["shapes"], c["P"])
        assert abs(syn.checksum(value, loc, w, go) - meta["checksum"]) < 1e-6 * max(1.0, abs(meta["checksum"]))
        shapes, lsi = _legacy_shapes(c["shapes"], dev)
        for tag, dt, tol in (("f32", torch.float32, FP32_TOL), ("f64", torch.float64, 1e-10)):
            gv, gl, gw = msda_ext.ms_deform_attn_backward(value.to(dev, dt), shapes, lsi, loc.to(dev, dt), w.to(dev, dt),
                                                          go.to(dev, dt), 64)
            assert gv.dtype == dt and gv.shape == value.shape and gl.shape == loc.shape and gw.shape == w.shape
            for name, t in (("grad_value", gv), ("grad_loc", gl), ("grad_w", gw)):
                assert rel_rms(t.cpu().numpy(), g[f"{name}_{tag}"]) < tol, (c["name"], tag, name)


@pytest.mark.parametrize("B,Q,name", [(1, 300, "C1"), (2, 357, "MOT17")])
def test_backward_vs_c_oracle_full_size(dev, B, Q, name):
    """Warp kernel (fp32 and bf16 value) against the plain-C restatement of the col2im arithmetic at the
    frame's real sizes; plus a size-independent property: the three gradients are linear in grad_out."""
    m, ops, syn, mg, tp = _mods()
    from oracle import c_core
    shapes = [list(s) for s in syn.PYRAMIDS[name]]
    value, loc, w, go = syn.make_core_grad_inputs(100 + Q, B, Q, 8, 32, shapes, 4)
    ref = c_core.msda_core_backward(value.numpy(), shapes, loc.numpy(), w.numpy(), go.numpy())
    v, l, a, g = value.to(dev), loc.to(dev), w.to(dev), go.to(dev)
    got = ops.msda_sampled_backward(v, shapes, l, a, g)
    for nm, t, r in zip(("grad_value", "grad_loc", "grad_w"), got, ref):
        assert rel_rms(t.cpu().numpy(), r) < FP32_TOL, (name, nm)
    got_bf = ops.msda_sampled_backward(v.to(torch.bfloat16), shapes, l, a, g)
    ref_bf = c_core.msda_core_backward(value.to(torch.bfloat16).float().numpy(), shapes, loc.numpy(), w.numpy(), go.numpy())
    for nm, t, r in zip(("grad_value", "grad_loc", "grad_w"), got_bf, ref_bf):
        assert t.dtype == torch.float32 and rel_rms(t.cpu().numpy(), r) < FP32_TOL, (name, "bf16 value", nm)
    got2 = ops.msda_sampled_backward(v, shapes, l, a, -2.0 * g)
    for t, t2 in zip(got, got2):
        assert rel_rms(t2.cpu().numpy(), -2.0 * t.cpu().numpy()) < 1e-5


def test_backward_gradcheck_like_reference_test(dev):
    """The reference's own gradient test (MOTR/models/ops/test.py:63-79): torch.autograd.gradcheck in double
    on MSDeformAttnFunction with its sizes (N=1, M=2, L=2, P=2, Lq=2), channels 30 and 32 here."""
    from moyolo_b200.msda_ext import MSDeformAttnFunction
    N, M, Lq, L, P = 1, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long, device=dev)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    torch.manual_seed(3)
    for channels in (30, 32):
        value = (torch.rand(N, S, M, channels, device=dev) * 0.01).double().requires_grad_(True)
        loc = torch.rand(N, Lq, M, L, P, 2, device=dev).double().requires_grad_(True)
        aw = torch.rand(N, Lq, M, L, P, device=dev) + 1e-5
        aw = (aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)).double().requires_grad_(True)
        assert torch.autograd.gradcheck(MSDeformAttnFunction.apply, (value, shapes, lsi, loc, aw, 2))


def test_backward_vs_reference_cuda_kernel(dev):
    """Our backward against the REFERENCE's own col2im kernel (ms_deform_im2col_cuda.cuh:956-1130) compiled for
    sm_100a into oracle/_ref (both accumulate grad_value with atomics in arbitrary order -> tolerance, not bits)."""
    from oracle import ref_cuda
    if not ref_cuda.available() or not hasattr(__import__("ctypes").CDLL(str(ref_cuda._SO)), "ref_msda_col2im_f32"):
        pytest.skip("oracle/_ref/libmsda_ref_cuda.so (with the backward wrapper) not built")
    m, ops, syn, mg, tp = _mods()
    for name, B, Q in (("tiny", 2, 50), ("MOT17", 1, 357)):
        shapes = [list(s) for s in syn.PYRAMIDS[name]]
        value, loc, w, go = syn.make_core_grad_inputs(7 + Q, B, Q, 8, 32, shapes, 4)
        v, l, a, g = value.to(dev), loc.to(dev), w.to(dev), go.to(dev)
        ref = ref_cuda.msda_col2im(v, shapes, l, a, g)
        got = ops.msda_sampled_backward(v, shapes, l, a, g)
        for nm, t, r in zip(("grad_value", "grad_loc", "grad_w"), got, ref):
            assert rel_rms(t.cpu().numpy(), r.cpu().numpy()) < FP32_TOL, (name, nm)


@pytest.mark.parametrize("B,Q,name,ref_dim", [(1, 382, "MOT17", 4), (2, 77, "tiny", 4), (1, 9, "tiny", 2)])
def test_msda_proj_fused_equals_gemm_plus_gather(dev, B, Q, name, ref_dim):
    """The gather with the sampling_offsets|attention_weights projection fused in (one launch) against the
    two-launch path it replaces (tcgen05 GEMM -> fused gather), and both against the fp32 oracle of
    MSDeformAttn's middle section (transformer.py:268-285) on the same bf16 operands. Ragged rows included."""
    m, ops, syn, mg, tp = _mods()
    H, D, L, P, C = 8, 32, 3, 4, 256
    shapes = [list(s) for s in syn.PYRAMIDS[name]]
    Lv = syn.level_sizes(shapes)
    g = torch.Generator().manual_seed(5 + Q)
    R = B * Q
    value = torch.randn(B, Lv, C, generator=g).to(dev, torch.bfloat16)
    xq = torch.randn(R, C, generator=g).to(dev, torch.bfloat16)
    w = (torch.randn(H * L * P * 3, C, generator=g) * 0.05).to(dev, torch.bfloat16)
    bias = (torch.randn(H * L * P * 3, generator=g) * 0.5).to(dev)
    cxcy = torch.rand(R, 1, 2, generator=g)
    refer = (torch.cat([cxcy, torch.rand(R, 1, 2, generator=g) * 0.4 + 0.02], -1) if ref_dim == 4 else cxcy).to(dev)
    ro = None
    if B == 2:  # ragged: sequence 0 owns 50 rows, sequence 1 the rest
        ro = torch.tensor([0, 50, R], dtype=torch.int32, device=dev)
    fused = ops.msda_proj_fused(value, shapes, xq, w, bias, refer, H, P, B, row_offsets=ro)
    ol = ops.linear(xq, w, bias, out_dtype=torch.float32)
    n_off = H * L * P * 2
    two = ops.msda_fused(value, shapes, ol[:, :n_off], ol[:, n_off:], refer, H, P, B, row_offsets=ro)
    # fp32 oracle on the same (bf16-rounded) operands
    olr = xq.float().cpu() @ w.float().cpu().T + bias.cpu()
    off = olr[:, :n_off].view(R, H, L, P, 2)
    aw = torch.softmax(olr[:, n_off:].view(R, H, L * P), -1).view(R, H, L, P)
    rf = refer.cpu()
    if ref_dim == 4:
        loc = rf[:, :, None, None, :2] + off / P * rf[:, :, None, None, 2:] * 0.5
    else:
        norm = torch.tensor([[s[1], s[0]] for s in shapes], dtype=torch.float32).view(1, 1, L, 1, 2)
        loc = rf[:, :, None, None, :] + off / norm
    bounds = [0, 50, R] if B == 2 else [0, R]
    ref = torch.cat([tp.msda_core_gather(value[b:b + 1].float().cpu().view(1, Lv, H, D), shapes,
                                         loc[bounds[b]:bounds[b + 1]][None], aw[bounds[b]:bounds[b + 1]][None])[0]
                     for b in range(B)])
    assert rel_rms(fused.float().cpu().numpy(), ref.numpy()) < BF16_TOL
    assert rel_rms(two.float().cpu().numpy(), ref.numpy()) < BF16_TOL
    assert rel_rms(fused.float().cpu().numpy(), two.float().cpu().numpy()) < BF16_TOL


def test_value_projection_ahead_equals_in_graph(dev):
    """The value projection launched AHEAD of the frame graph (own stream, CTA budget, gated on the previous
    frame's tail event, double-buffered value tensor) must give exactly the rows of the engine that keeps it
    inside the graph: same kernel, same tiles -> bit-identical IDs, boxes and scores over a pipelined sequence
    on the C1 pyramid (Lv = 8400, large enough for the persistent kernel), incl. the synchronous step() path."""
    m, ops, syn, mg, tp = _mods()
    from moyolo_b200.tracker import DecoderWeights, TrackEngine
    spec = syn.DecoderSpec()
    shapes = [list(s) for s in syn.PYRAMIDS["C1"]]
    sd = syn.make_decoder_state(spec, 3)
    n_frames, nd, S = 7, 100, 1
    gen = syn.SequenceGenerator(syn.SequenceSpec(name="C1", n_frames=n_frames, n_detect=nd, seed=5, shapes=shapes),
                                spec.d_model, dev)
    frames = [tuple(t.clone() for t in gen.next_frame()) for _ in range(n_frames)]
    eng0 = TrackEngine(sd, spec, shapes, dev, "bf16", nd, S, value_ahead=False)
    out0 = eng0.step(frames[0][0][None], frames[0][1][None], frames[0][2][None])[0]
    sd = syn.calibrate_score_bias(sd, out0["logits"], spec, 0.1)
    W = DecoderWeights(sd, spec, dev, "bf16")
    batches = [tuple(x[None].to(torch.bfloat16 if i == 0 else torch.float32).contiguous() for i, x in enumerate(f))
               for f in frames]
    results = {}
    for ahead in (False, True):
        eng = TrackEngine(sd, spec, shapes, dev, "bf16", nd, S, weights=W, value_ahead=ahead)
        assert eng._vp_ahead == ahead
        got = {}
        for t in range(n_frames):
            eng.submit(*batches[t], want_rows=True)
            if t > 0:
                got[t - 1] = {k: v.clone() for k, v in eng.collect(t - 1)[0].items()}
        got[n_frames - 1] = {k: v.clone() for k, v in eng.collect(n_frames - 1)[0].items()}
        results[ahead] = (got, eng.track_table().clone().cpu(), eng.n_tracks_host())
        # synchronous path on a fresh sequence
        eng.reset()
        outs = [eng.step(*batches[t])[0] for t in range(3)]
        results[(ahead, "step")] = [{k: v.clone().cpu() for k, v in o.items()} for o in outs]
    a, b = results[False], results[True]
    assert a[2] == b[2] and max(a[2]) > 0, "no track was carried"
    assert torch.equal(a[1], b[1])
    for t in range(n_frames):
        for k in ("ids", "boxes", "scores", "labels"):
            assert torch.equal(a[0][t][k], b[0][t][k]), (t, k)
    for oa, ob in zip(results[(False, "step")], results[(True, "step")]):
        for k in oa:
            assert torch.equal(oa[k], ob[k]), ("step", k)


@pytest.mark.parametrize("ref_dim,with_mask", [(4, False), (2, True)])
def test_msdeform_attn_module_training_path(dev, ref_dim, with_mask):
    """MSDeformAttn with `differentiable = True`: output and the gradients w.r.t. every parameter and input against
    torch.autograd through the fp64 CPU oracle of the module (oracle/torch_port.msdeform_attn_forward, pinned to the
    reference by msda_*.npz); without the switch an input that requires grad is refused loudly."""
    m, ops, syn, mg, tp = _mods()
    spec = syn.DecoderSpec()
    shapes = [list(s) for s in syn.PYRAMIDS["tiny"]]
    B, Q = 2, 19
    attn = m.MSDeformAttn(spec.d_model, spec.n_levels, spec.n_heads, spec.n_points)
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():   # non-degenerate offsets / logits (the default init zeroes those weights)
        attn.sampling_offsets.weight.copy_(torch.randn(attn.sampling_offsets.weight.shape, generator=g) * 0.02)
        attn.attention_weights.weight.copy_(torch.randn(attn.attention_weights.weight.shape, generator=g) * 0.05)
    query, refer, feats, _ = syn.make_module_inputs(3, B, Q, spec.d_model, shapes, ref_dim=ref_dim,
                                                    ref_levels=1 if ref_dim == 4 else spec.n_levels)
    mask = (torch.rand(B, feats.shape[1], generator=g) < 0.1) if with_mask else None
    go = torch.randn(B, Q, spec.d_model, generator=g)
    # oracle: fp64 CPU autograd
    p64 = {k: v.detach().double().clone().requires_grad_(True) for k, v in attn.state_dict().items()}
    q64, v64, r64 = (t.double().clone().requires_grad_(True) for t in (query, feats, refer))
    out64 = tp.msdeform_attn_forward(p64, q64, r64, v64, shapes, spec.n_heads, spec.n_levels, spec.n_points, mask,
                                     core=tp.msda_core_gather)
    out64.backward(go.double())
    # device: training path
    attn = attn.to(dev)
    qd, vd, rd = (t.to(dev).clone().requires_grad_(True) for t in (query, feats, refer))
    with pytest.raises(RuntimeError, match="differentiable"):
        attn(qd, rd, vd, shapes, None if mask is None else mask.to(dev))
    attn.differentiable = True
    out = attn(qd, rd, vd, shapes, None if mask is None else mask.to(dev))
    out.backward(go.to(dev))
    assert rel_rms(out.detach().cpu().numpy(), out64.detach().numpy()) < FP32_TOL
    for name, got, ref in (("query", qd.grad, q64.grad), ("value", vd.grad, v64.grad), ("refer", rd.grad, r64.grad)):
        assert rel_rms(got.cpu().numpy(), ref.numpy()) < 5e-4, name
    for k, prm in attn.named_parameters():
        assert rel_rms(prm.grad.cpu().numpy(), p64[k].grad.numpy()) < 5e-4, k
    # and under no_grad the module still runs the fused inference kernels
    with torch.no_grad():
        out_inf = attn(qd.detach(), rd.detach(), vd.detach(), shapes, None if mask is None else mask.to(dev))
    assert rel_rms(out_inf.cpu().numpy(), out64.detach().numpy()) < FP32_TOL


# ------------------------------------------------------------------ a9: QIM against the reference goldens, directly
@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_MODULE_TOL)])
@pytest.mark.parametrize("name", ["qim_a", "qim_one"])
def test_qim_update_vs_reference_golden(dev, name, precision, tol):
    """QueryInteractionModule._update_track_embedding (MOTR/models/qim.py:251-301): the frame's QIM kernels
    (tracker.qim_update -> qim_update_ws) against outputs of the reference module itself (oracle/make_golden.py
    gen_qim), T = 23 tracks and the single-track edge case."""
    m, ops, syn, mg, tp = _mods()
    from moyolo_b200.tracker import DecoderWeights, qim_update
    meta, g = load_golden(name)
    spec = syn.DecoderSpec()
    sd = syn.make_decoder_state(spec, meta["weight_seed"])
    W = DecoderWeights(sd, spec, dev, precision)
    t = {k: torch.from_numpy(v).to(dev) for k, v in g.items()}
    qp, rp = qim_update(W, t["ref_pts"], t["query_pos"], t["out_embed"], t["pred_boxes"])
    assert qp.shape == g["new_query_pos"].shape
    assert rel_rms(qp.cpu().numpy(), g["new_query_pos"]) < tol
    assert np.allclose(rp.cpu().numpy(), g["new_ref_pts"], rtol=2e-6, atol=2e-6)   # logf rounding only


# ------------------------------------------------------------------ a3: MOTRMSDeformAttn(my_softmax=True)
@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", 4e-2)])   # max-norm over up to 76 800 outputs
def test_motr_msdeform_attn_my_softmax(dev, precision, tol):
    """transformer.py:239-244, 369-371: attention weights exp(x) / (1 + sum exp(x)) instead of softmax. The
    reference method lacks `self` and raises when enabled, so the check is against the oracle's restatement of
    :241-244 (documented there as unpinned). Also checks that the flag changes the result and that the plain
    MSDeformAttn class ignores it (it never reads my_softmax, transformer.py:271)."""
    m, ops, syn, mg, tp = _mods()
    spec = syn.DecoderSpec()
    sd = syn.make_decoder_state(spec, 5)
    p = syn.sub_state(sd, "layers.0.cross_attn.")
    for B, Q, pyr, ref_dim, ref_levels in ((2, 77, "tiny", 4, 1), (1, 300, "C1", 4, 1), (1, 9, "tiny", 2, 3)):
        shapes = [list(s) for s in syn.PYRAMIDS[pyr]]
        q, refer, feats, _ = syn.make_module_inputs(31 + Q, B, Q, spec.d_model, shapes, ref_dim, ref_levels)
        with torch.no_grad():
            ref = tp.msdeform_attn_forward(p, q, refer, feats, shapes, spec.n_heads, spec.n_levels, spec.n_points,
                                           my_softmax=True).numpy()
            ref_plain = tp.msdeform_attn_forward(p, q, refer, feats, shapes, spec.n_heads, spec.n_levels,
                                                 spec.n_points).numpy()
        mod = m.MOTRMSDeformAttn(spec.d_model, spec.n_levels, spec.n_heads, spec.n_points, my_softmax=True)
        mod.load_state_dict(p)
        mod = mod.to(dev).eval()
        mod.precision = precision
        out = mod(q.to(dev), refer.to(dev), feats.to(dev), shapes).cpu().numpy()
        assert rel_rms(out, ref) < tol, (B, Q, pyr)
        assert rel_rms(ref, ref_plain) > 0.05, "the flag must matter for this check to mean anything"
        plain = m.MSDeformAttn(spec.d_model, spec.n_levels, spec.n_heads, spec.n_points, my_softmax=True)
        plain.load_state_dict(p)
        plain = plain.to(dev).eval()
        plain.precision = precision
        out2 = plain(q.to(dev), refer.to(dev), feats.to(dev), shapes).cpu().numpy()
        assert rel_rms(out2, ref_plain) < tol, (B, Q, pyr, "MSDeformAttn ignores my_softmax")


# ------------------------------------------------------------------ a4 / a5 with attn_mask and padding_mask
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_masked_layers_and_decoders_vs_reference_golden(dev, precision):
    """Self-attention mask (bool [Q, Q] as ultralytics/models/utils/ops.py:363-375 builds it, and the additive float
    form nn.MultiheadAttention also takes) and value padding mask through the drop-in layer and decoder classes
    (transformer.py:637-645, 265-266) against goldens minted from the unmodified reference."""
    m, ops, syn, mg, tp = _mods()
    for case in mg.MASKED_CASES:
        meta, g = load_golden(case["name"])
        spec = syn.DecoderSpec(nc=case.get("nc", 1))
        sd = syn.make_decoder_state(spec, meta["weight_seed"])
        attn, pad = mg.masks_for(case)
        attn, pad = attn.to(dev), pad.to(dev)
        if case["kind"] == "layer":
            layer = getattr(m, case["cls"])(spec.d_model, spec.n_heads, spec.d_ffn, 0.0, torch.nn.ReLU(), spec.n_levels,
                                            spec.n_points)
            layer = _load_layer(m, syn, sd, "layers.1.", layer).to(dev)
            layer.precision = precision
            q, refer, feats, qpos = syn.make_module_inputs(case["seed"], case["B"], case["Q"], spec.d_model, case["shapes"], 4, 1)
            out = layer(q.to(dev), refer[:, :, 0].to(dev), feats.to(dev), meta["shapes"], pad, attn, qpos.to(dev))
            assert rel_rms(out.cpu().numpy(), g["out"]) < (FP32_TOL if precision == "fp32" else BF16_MODULE_TOL), case["name"]
        else:
            dec, bbox, score, pos = _build_decoder(m, syn, spec, sd, case["mode"], dev, precision)
            embed, refer, feats, qpos = syn.make_decoder_inputs(case["seed"], case["B"], case["Q"], spec.d_model, case["shapes"])
            args = (embed.to(dev), refer.to(dev), feats.to(dev), meta["shapes"], bbox, score, pos)
            if case["mode"] == "motr":
                b, s, _ = dec(*args, attn_mask=attn, padding_mask=pad, track_query_embed=qpos.to(dev))
            else:
                b, s = dec(*args, attn_mask=attn, padding_mask=pad)
            if precision == "fp32":
                assert rel_rms(b.cpu().numpy(), g["boxes"]) < FP32_TOL and rel_rms(s.cpu().numpy(), g["scores"]) < FP32_TOL, case["name"]
            else:
                assert float(np.abs(b.cpu().numpy() - g["boxes"]).max()) < 5e-3, case["name"]
                assert rel_rms(s.cpu().numpy(), g["scores"]) < 5e-2, case["name"]
    # masks the kernels do not implement are refused, not ignored
    layer = m.MOTRDecoderLayer(256, 8, 1024, 0.0, torch.nn.ReLU(), 3, 4).to(dev).eval()
    q, refer, feats, qpos = syn.make_module_inputs(1, 1, 8, 256, syn.PYRAMIDS["tiny"], 4, 1)
    with pytest.raises(NotImplementedError):
        layer(q.to(dev), refer[:, :, 0].to(dev), feats.to(dev), [list(x) for x in syn.PYRAMIDS["tiny"]], None,
              torch.zeros(8, 8, 8, dtype=torch.bool, device=dev), qpos.to(dev))


def test_headmajor_gather_bit_equal(dev):
    """moyolo_msda_fused_forward_headmajor (value stored [B, H, Lv, 32]; profiles/r02_experiments/gather_head_major.md)
    runs the same arithmetic as the channel-last entry point: bit-identical output, incl. ragged batches."""
    m, ops, syn, mg, tp = _mods()
    shapes = [list(s) for s in syn.PYRAMIDS["MOT17"]]
    Lv = syn.level_sizes(shapes)
    H, D, L, P, B = 8, 32, 3, 4, 3
    g = torch.Generator().manual_seed(5)
    allv = torch.randn(B, Lv, 2 * H * D, generator=g).to(dev).to(torch.bfloat16)
    v_cl = allv[:, :, H * D:]
    v_hm = v_cl.reshape(B, Lv, H, D).permute(0, 2, 1, 3).contiguous()
    row_offsets = torch.tensor([0, 400, 401, 1100], dtype=torch.int32, device=dev)
    R = 1100
    offsets = torch.randn(R, H * L * P * 2, generator=g).to(dev) * 3
    logits = torch.randn(R, H * L * P, generator=g).to(dev)
    refer = torch.cat([torch.rand(R, 1, 2, generator=g), torch.rand(R, 1, 2, generator=g) * 0.4], -1).to(dev)
    a = ops.msda_fused(v_cl, shapes, offsets, logits, refer, H, P, B, row_offsets=row_offsets)
    b = ops.msda_fused_headmajor(v_hm, shapes, offsets, logits, refer, P, row_offsets=row_offsets)
    assert torch.equal(a, b) and float(a.float().abs().max()) > 0
