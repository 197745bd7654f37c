"""CPU: pin the oracle (oracle/torch_port.py, oracle/msda_core.c, oracle/tracker_port.py) against the
golden vectors minted from the unmodified reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_rms
from moyolo_b200 import synthetic as syn
from oracle import c_core, make_golden as mg
from oracle import torch_port as tp
from oracle.tracker_port import TrackerPort

FP32_TOL = 1e-4  # max|a-b| <= 1e-4 * rms(ref), SURVEY.md §8(c)


def test_kat0_reference_known_answer():
    """MOTR/models/ops/test.py:21-60: fp64 allclose (default tol), fp32 rtol=1e-2 atol=1e-3."""
    meta, g = load_golden("kat0")
    shapes = meta["shapes"]
    for tag, dt in (("double", np.float64), ("float", np.float32)):
        v, loc, aw = (g[f"{k}_{tag}"].astype(dt) for k in ("value", "loc", "aw"))
        ref = g[f"out_{tag}"]
        out_c = c_core.msda_core(v, shapes, loc, aw)
        tv, tl, ta = map(torch.from_numpy, (v, loc, aw))
        out_gs = tp.msda_core_gridsample(tv, shapes, tl, ta).numpy()
        out_ga = tp.msda_core_gather(tv, shapes, tl, ta).numpy()
        for o in (out_c, out_gs, out_ga):
            if tag == "double":
                assert np.allclose(o, ref)
            else:
                assert np.allclose(o, ref, rtol=1e-2, atol=1e-3)
            assert rel_rms(o, ref) < FP32_TOL


@pytest.mark.parametrize("case", mg.CORE_CASES, ids=lambda c: c["name"])
def test_core_restatements(case):
    meta, g = load_golden(case["name"])
    value, loc, w = syn.make_core_inputs(case["seed"], case["B"], case["Q"], case["H"], case["D"], case["shapes"],
                                         case["P"])
    assert abs(syn.checksum(value, loc, w) - meta["checksum"]) < 1e-6 * max(1.0, abs(meta["checksum"]))
    shapes = meta["shapes"]
    assert rel_rms(tp.msda_core_gridsample(value, shapes, loc, w).numpy(), g["out_f32"]) < 1e-6
    assert rel_rms(tp.msda_core_gather(value, shapes, loc, w).numpy(), g["out_f32"]) < FP32_TOL
    assert rel_rms(c_core.msda_core(value.numpy(), shapes, loc.numpy(), w.numpy()), g["out_f32"]) < FP32_TOL
    o64 = c_core.msda_core(value.double().numpy(), shapes, loc.double().numpy(), w.double().numpy())
    assert rel_rms(o64, g["out_f64"]) < 1e-12
    o64 = tp.msda_core_gather(value.double(), shapes, loc.double(), w.double()).numpy()
    assert rel_rms(o64, g["out_f64"]) < 1e-12


@pytest.mark.parametrize("case", mg.GRAD_CASES, ids=lambda c: c["name"])
def test_core_backward_restatement(case):
    """oracle/msda_core.c backward (col2im arithmetic, ms_deform_im2col_cuda.cuh:88-159) against
    torch.autograd of the reference's multi_scale_deformable_attn_pytorch."""
    meta, g = load_golden(case["name"])
    value, loc, w, go = syn.make_core_grad_inputs(case["seed"], case["B"], case["Q"], case["H"], case["D"],
                                                  case["shapes"], case["P"])
    assert abs(syn.checksum(value, loc, w, go) - meta["checksum"]) < 1e-6 * max(1.0, abs(meta["checksum"]))
    for tag, dt, tol in (("f32", np.float32, FP32_TOL), ("f64", np.float64, 1e-12)):
        got = c_core.msda_core_backward(value.numpy().astype(dt), meta["shapes"], loc.numpy().astype(dt),
                                        w.numpy().astype(dt), go.numpy().astype(dt))
        for name, a in zip(("grad_value", "grad_loc", "grad_w"), got):
            assert rel_rms(a, g[f"{name}_{tag}"]) < tol, (tag, name)


def test_posemb_and_inverse_sigmoid():
    _, g = load_golden("posemb")
    assert rel_rms(tp.pos2posemb(torch.from_numpy(g["pos"])).numpy(), g["emb"]) < 1e-6
    assert np.array_equal(tp.inverse_sigmoid(torch.from_numpy(g["x"])).numpy(), g["inv"])


@pytest.mark.parametrize("case", mg.MSDA_CASES, ids=lambda c: c["name"])
def test_msdeform_attn_port(case):
    meta, g = load_golden(case["name"])
    spec = syn.DecoderSpec()
    sd = syn.make_decoder_state(spec, meta["weight_seed"])
    q, refer, feats, _ = syn.make_module_inputs(case["seed"], case["B"], case["Q"], spec.d_model, case["shapes"],
                                                case["ref_dim"], case["ref_levels"])
    assert abs(syn.checksum(q, refer, feats) - meta["checksum"]) < 1e-6 * max(1.0, abs(meta["checksum"]))
    mask = None
    if case["mask"]:
        mask = torch.rand(case["B"], feats.shape[1], generator=torch.Generator().manual_seed(case["seed"])) < 0.2
    p = syn.sub_state(sd, "layers.0.cross_attn.")
    for core in (tp.msda_core_gridsample, tp.msda_core_gather):
        o = tp.msdeform_attn_forward(p, q, refer, feats, meta["shapes"], spec.n_heads, spec.n_levels, spec.n_points,
                                     mask, core)
        assert rel_rms(o.numpy(), g["out"]) < FP32_TOL


@pytest.mark.parametrize("case", mg.LAYER_CASES, ids=lambda c: c["name"])
def test_decoder_layer_port(case):
    meta, g = load_golden(case["name"])
    spec = syn.DecoderSpec()
    sd = syn.make_decoder_state(spec, meta["weight_seed"])
    q, refer, feats, qpos = syn.make_module_inputs(case["seed"], case["B"], case["Q"], spec.d_model, case["shapes"], 4, 1)
    o = tp.decoder_layer_forward(syn.sub_state(sd, "layers.1."), q, refer[:, :, 0], feats, meta["shapes"],
                                 spec.n_heads, spec.n_levels, spec.n_points, None, None, qpos)
    assert rel_rms(o.numpy(), g["out"]) < FP32_TOL


@pytest.mark.parametrize("case", [c for c in mg.DECODER_CASES if "c1" not in c["name"]], ids=lambda c: c["name"])
def test_decoder_port(case):
    meta, g = load_golden(case["name"])
    spec = syn.DecoderSpec(nc=case["nc"])
    sd = syn.make_decoder_state(spec, meta["weight_seed"])
    embed, refer, feats, qpos = syn.make_decoder_inputs(case["seed"], case["B"], case["Q"], spec.d_model, case["shapes"])
    assert abs(syn.checksum(embed, refer, feats, qpos) - meta["checksum"]) < 1e-6 * max(1.0, abs(meta["checksum"]))
    with torch.no_grad():
        b, s, hs = tp.decoder_forward(sd, embed, refer, feats, meta["shapes"], spec.n_heads, spec.n_levels,
                                      spec.n_points, spec.n_layers, case["mode"], qpos)
    assert b.shape == g["boxes"].shape and s.shape == g["scores"].shape
    assert rel_rms(b.numpy(), g["boxes"]) < FP32_TOL
    assert rel_rms(s.numpy(), g["scores"]) < FP32_TOL
    if case["mode"] == "motr":
        assert rel_rms(hs.numpy(), g["hs"]) < FP32_TOL


@pytest.mark.parametrize("name", ["tracker_a", "tracker_b", "tracker_empty"])
def test_tracker_port_matches_reference_class(name):
    """Bit-exact IDs / disappear counters / max_obj_id against the reference RuntimeTrackerBase."""
    meta, g = load_golden(name)
    trk = TrackerPort()
    for t in range(meta["n_frames"]):
        assert (trk.max_obj_id, trk.max_obj_id_pre) == tuple(int(x) for x in g[f"counters_in_{t}"])
        ids, dis = g[f"ids_in_{t}"].copy(), g[f"dis_in_{t}"].copy()
        trk.update(g[f"scores_{t}"], g[f"boxes_{t}"], ids, dis)
        assert np.array_equal(ids, g[f"ids_out_{t}"]), f"frame {t}"
        assert np.array_equal(dis, g[f"dis_out_{t}"]), f"frame {t}"
        assert (trk.max_obj_id, trk.max_obj_id_pre) == tuple(int(x) for x in g[f"counters_out_{t}"]), f"frame {t}"


@pytest.mark.parametrize("name", ["qim_a", "qim_one"])
def test_qim_port(name):
    meta, g = load_golden(name)
    sd = syn.make_decoder_state(syn.DecoderSpec(), meta["weight_seed"])
    t = {k: torch.from_numpy(v) for k, v in g.items()}
    with torch.no_grad():
        qp, rp = tp.qim_update(sd, t["ref_pts"], t["query_pos"], t["out_embed"], t["pred_boxes"])
    assert rel_rms(qp.numpy(), g["new_query_pos"]) < FP32_TOL
    assert np.array_equal(rp.numpy(), g["new_ref_pts"])


@pytest.mark.parametrize("name", ["select_tiny", "select_c1", "select_kitti_nc5"])
def test_query_selection_restatement(name):
    """oracle encoder_input / generate_anchors / query_selection (head.py:993-1113) against the reference MYDecoder."""
    meta, g = load_golden(name)
    spec = syn.DecoderSpec(nc=meta["nc"])
    sd = syn.make_selector_state(spec, (256, 512, 512), meta["weight_seed"])
    maps = syn.make_pyramid_maps(meta["seed"], meta["B"], meta["shapes"])
    assert abs(syn.checksum(*maps) - meta["checksum"]) < 1e-6 * max(1.0, abs(meta["checksum"]))
    with torch.no_grad():
        feats, shapes = tp.encoder_input(sd, maps)
        sel = tp.query_selection(sd, feats, shapes, meta["nq"])
    assert shapes == meta["shapes"]
    assert rel_rms(feats[:, ::meta["feats_row_step"]].numpy(), g["feats"]) < 1e-6
    # the selected set and its ORDER (torch.topk, descending) are part of the contract: IDs follow query order
    assert rel_rms(sel["enc_scores"].numpy(), g["enc_scores"]) < 1e-6
    assert rel_rms(sel["embed"].numpy(), g["embed"]) < 1e-6
    ref, got = g["refer"], sel["refer"].numpy()
    fin = np.isfinite(ref)
    assert np.array_equal(fin, np.isfinite(got))  # anchors outside (eps, 1-eps) are +inf (head.py:1006-1009)
    assert rel_rms(got[fin], ref[fin]) < 1e-6
    assert rel_rms(tp.pos2posemb(sel["refer"]).nan_to_num(0.0, 0.0, 0.0).numpy(),
                   np.nan_to_num(g["query_pos"], nan=0.0, posinf=0.0, neginf=0.0)) < 1e-5


@pytest.mark.parametrize("case", mg.MASKED_CASES, ids=lambda c: c["name"])
def test_masked_layers_and_decoders_port(case):
    """attn_mask (bool [Q, Q] of the denoising groups, or additive float) and padding_mask at layer and decoder level
    (transformer.py:637-645; models/utils/ops.py:363-375) against goldens minted from the unmodified reference."""
    meta, g = load_golden(case["name"])
    spec = syn.DecoderSpec(nc=case.get("nc", 1))
    sd = syn.make_decoder_state(spec, meta["weight_seed"])
    attn, pad = mg.masks_for(case)
    with torch.no_grad():
        if case["kind"] == "layer":
            q, refer, feats, qpos = syn.make_module_inputs(case["seed"], case["B"], case["Q"], spec.d_model, case["shapes"], 4, 1)
            out = tp.decoder_layer_forward(syn.sub_state(sd, "layers.1."), q, refer[:, :, 0], feats, meta["shapes"],
                                           spec.n_heads, spec.n_levels, spec.n_points, pad, attn, qpos)
            assert rel_rms(out.numpy(), g["out"]) < FP32_TOL
        else:
            embed, refer, feats, qpos = syn.make_decoder_inputs(case["seed"], case["B"], case["Q"], spec.d_model, case["shapes"])
            b, s, hs = tp.decoder_forward(sd, embed, refer, feats, meta["shapes"], spec.n_heads, spec.n_levels,
                                          spec.n_points, spec.n_layers, case["mode"], qpos, attn_mask=attn, padding_mask=pad)
            assert rel_rms(b.numpy(), g["boxes"]) < FP32_TOL and rel_rms(s.numpy(), g["scores"]) < FP32_TOL
            if case["mode"] == "motr":
                assert rel_rms(hs.numpy(), g["hs"]) < FP32_TOL
