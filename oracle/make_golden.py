"""ORACLE — TEST INFRASTRUCTURE ONLY. Mints tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
Inputs and weights are regenerated from seeds by moyolo_b200.synthetic (plain data generators, no
compute); each file stores the reference outputs, the seeds/config and a checksum of the inputs so
a consumer can detect RNG drift. Reference entry points executed:
  kat0      MOTR/models/ops/functions/ms_deform_attn_func.py:44-64 on the case of MOTR/models/ops/test.py:21-60
  core_*    ultralytics/nn/modules/utils.py:41-78   multi_scale_deformable_attn_pytorch
  core_grad_*  torch.autograd through the same function (the gradients ms_deform_attn_backward returns)
  msda_*    ultralytics/nn/modules/transformer.py:193-287  MSDeformAttn
  layer_*   transformer.py:394-450 DeformableTransformerDecoderLayer, :515-652 MOTRDecoderLayer
  decoder_* transformer.py:453-510 DeformableTransformerDecoder, :663-728 MOTRTransformerDecoder
  posemb    transformer.py:183-190 pos2posemb; utils.py:34-38 inverse_sigmoid
  tracker_* ultralytics/nn/modules/head.py:1143-1283 RuntimeTrackerBase on MOTR Instances
  qim_*     MOTR/models/qim.py:251-301 QueryInteractionModule._update_track_embedding
  formats   MOTR/submit_dance.py:410-419 Detector.write_results; ultralytics/engine/results.py:475-512 save_txt
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from moyolo_b200 import synthetic as syn  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLD = ROOT / "tests" / "golden"


def save(name, meta, **arrays):
    GOLD.mkdir(parents=True, exist_ok=True)
    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(GOLD / f"{name}.npz", meta=json.dumps(meta), **arrays)
    print(f"  wrote {name}.npz  ({(GOLD / (name + '.npz')).stat().st_size / 1024:.1f} KiB)")


CORE_CASES = [
    dict(name="core_a", seed=11, B=2, Q=50, H=8, D=32, shapes=syn.PYRAMIDS["tiny"], P=4),
    dict(name="core_b", seed=12, B=1, Q=37, H=4, D=64, shapes=[(5, 7), (3, 2)], P=8),
    dict(name="core_c", seed=13, B=3, Q=9, H=2, D=2, shapes=[(6, 4), (3, 2), (5, 7), (1, 1)], P=2),
]
MSDA_CASES = [
    dict(name="msda_ref4", seed=21, B=2, Q=33, shapes=syn.PYRAMIDS["tiny"], ref_dim=4, ref_levels=1, mask=False),
    dict(name="msda_ref2", seed=22, B=1, Q=20, shapes=syn.PYRAMIDS["tiny"], ref_dim=2, ref_levels=3, mask=False),
    dict(name="msda_mask", seed=23, B=2, Q=17, shapes=syn.PYRAMIDS["tiny"], ref_dim=4, ref_levels=1, mask=True),
]
LAYER_CASES = [
    dict(name="layer_deformable", seed=31, cls="DeformableTransformerDecoderLayer", B=2, Q=40, shapes=syn.PYRAMIDS["tiny"]),
    dict(name="layer_motr", seed=32, cls="MOTRDecoderLayer", B=1, Q=77, shapes=syn.PYRAMIDS["tiny"]),
]
DECODER_CASES = [
    dict(name="decoder_motr_tiny", seed=41, mode="motr", B=2, Q=40, shapes=syn.PYRAMIDS["tiny"], nc=1),
    dict(name="decoder_deformable_tiny", seed=42, mode="deformable", B=1, Q=33, shapes=syn.PYRAMIDS["tiny"], nc=1),
    dict(name="decoder_motr_kitti_nc5", seed=43, mode="motr", B=1, Q=64, shapes=[(12, 39), (6, 20), (3, 10)], nc=5),
    dict(name="decoder_motr_c1", seed=44, mode="motr", B=1, Q=300, shapes=syn.PYRAMIDS["C1"], nc=1),
    dict(name="decoder_deformable_c1", seed=45, mode="deformable", B=1, Q=300, shapes=syn.PYRAMIDS["C1"], nc=1),
]
WEIGHT_SEED = 7


def gen_kat0():
    """MOTR/models/ops/test.py:21-60 with the CUDA op replaced by nothing: we keep the PyTorch side."""
    import importlib
    f = importlib.import_module("MOTR.models.ops.functions.ms_deform_attn_func")
    N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    torch.manual_seed(3)
    out = {}
    for tag in ("double", "float"):
        value = torch.rand(N, S, M, D) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2)
        aw = torch.rand(N, Lq, M, L, P) + 1e-5
        aw /= aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
        if tag == "double":
            o = f.ms_deform_attn_core_pytorch(value.double(), shapes, loc.double(), aw.double())
        else:
            o = f.ms_deform_attn_core_pytorch(value, shapes, loc, aw)
        out.update({f"value_{tag}": value, f"loc_{tag}": loc, f"aw_{tag}": aw, f"out_{tag}": o})
    save("kat0", dict(shapes=[[6, 4], [3, 2]], N=N, M=M, D=D, Lq=Lq, L=L, P=P, seed=3,
                      source="MOTR/models/ops/test.py:21-60"), **out)


def gen_core(U):
    for c in CORE_CASES:
        value, loc, w = syn.make_core_inputs(c["seed"], c["B"], c["Q"], c["H"], c["D"], c["shapes"], c["P"])
        o32 = U.multi_scale_deformable_attn_pytorch(value, c["shapes"], loc, w)
        o64 = U.multi_scale_deformable_attn_pytorch(value.double(), c["shapes"], loc.double(), w.double())
        save(c["name"], dict(c, shapes=[list(s) for s in c["shapes"]], checksum=syn.checksum(value, loc, w)),
             out_f32=o32, out_f64=o64)


GRAD_CASES = [
    dict(name="core_grad_a", seed=51, B=2, Q=24, H=8, D=32, shapes=syn.PYRAMIDS["tiny"], P=4),
    dict(name="core_grad_b", seed=52, B=1, Q=19, H=4, D=64, shapes=[(5, 7), (3, 2)], P=8),
    dict(name="core_grad_c", seed=53, B=3, Q=9, H=2, D=6, shapes=[(6, 4), (3, 2), (5, 7), (1, 1)], P=2),
]


def gen_core_grad(U):
    """Autograd of the reference's multi_scale_deformable_attn_pytorch (what MSDeformAttnFunction.backward /
    ms_deform_attn_backward compute, MOTR/models/ops/functions/ms_deform_attn_func.py:33-41), fp32 and fp64."""
    for c in GRAD_CASES:
        value, loc, w, go = syn.make_core_grad_inputs(c["seed"], c["B"], c["Q"], c["H"], c["D"], c["shapes"], c["P"])
        out = {}
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            v, l, a = (t.to(dt).clone().requires_grad_(True) for t in (value, loc, w))
            o = U.multi_scale_deformable_attn_pytorch(v, c["shapes"], l, a)
            o.backward(go.to(dt))
            out.update({f"grad_value_{tag}": v.grad, f"grad_loc_{tag}": l.grad, f"grad_w_{tag}": a.grad})
        save(c["name"], dict(c, shapes=[list(s) for s in c["shapes"]], checksum=syn.checksum(value, loc, w, go),
                             source="autograd of ultralytics/nn/modules/utils.py:41-78"), **out)


def gen_posemb(T, U):
    g = torch.Generator().manual_seed(5)
    pos = torch.randn(64, 4, generator=g) * 3.0
    x = torch.cat([torch.rand(60, 4, generator=g), torch.tensor([[0.0, 1.0, 1e-7, 1 - 1e-7], [-0.5, 1.5, 0.5, 2e-5]])])
    save("posemb", dict(seed=5), pos=pos, emb=T.pos2posemb(pos), x=x, inv=U.inverse_sigmoid(x))


def gen_msda(T):
    spec = syn.DecoderSpec()
    sd = syn.make_decoder_state(spec, WEIGHT_SEED)
    for c in MSDA_CASES:
        m = T.MSDeformAttn(spec.d_model, spec.n_levels, spec.n_heads, spec.n_points).eval()
        m.load_state_dict(syn.sub_state(sd, "layers.0.cross_attn."))
        q, refer, feats, _ = syn.make_module_inputs(c["seed"], c["B"], c["Q"], spec.d_model, c["shapes"], c["ref_dim"],
                                                    c["ref_levels"])
        mask = None
        if c["mask"]:
            mask = torch.rand(c["B"], feats.shape[1], generator=torch.Generator().manual_seed(c["seed"])) < 0.2
        with torch.no_grad():
            o = m(q, refer, feats, [list(s) for s in c["shapes"]], mask)
        save(c["name"], dict(c, shapes=[list(s) for s in c["shapes"]], weight_seed=WEIGHT_SEED,
                             checksum=syn.checksum(q, refer, feats)), out=o)


def gen_layers(T):
    spec = syn.DecoderSpec()
    sd = syn.make_decoder_state(spec, WEIGHT_SEED)
    for c in LAYER_CASES:
        cls = getattr(T, c["cls"])
        m = cls(spec.d_model, spec.n_heads, spec.d_ffn, 0.0, torch.nn.ReLU(), spec.n_levels, spec.n_points).eval()
        m.load_state_dict(syn.sub_state(sd, "layers.1."))
        q, refer, feats, qpos = syn.make_module_inputs(c["seed"], c["B"], c["Q"], spec.d_model, c["shapes"], 4, 1)
        refer = refer[:, :, 0]
        with torch.no_grad():
            o = m(q, refer, feats, [list(s) for s in c["shapes"]], None, None, qpos)
        save(c["name"], dict(c, shapes=[list(s) for s in c["shapes"]], weight_seed=WEIGHT_SEED,
                             checksum=syn.checksum(q, refer, feats, qpos)), out=o)


MASKED_CASES = [   # attn_mask / padding_mask at layer and decoder level (transformer.py:637-645, models/utils/ops.py:363-375)
    dict(name="layer_motr_masked", kind="layer", cls="MOTRDecoderLayer", seed=71, B=2, Q=50, shapes=syn.PYRAMIDS["tiny"], float_mask=False),
    dict(name="layer_deformable_masked_f", kind="layer", cls="DeformableTransformerDecoderLayer", seed=72, B=1, Q=37,
         shapes=syn.PYRAMIDS["tiny"], float_mask=True),
    dict(name="decoder_motr_masked", kind="decoder", mode="motr", seed=73, B=2, Q=48, shapes=syn.PYRAMIDS["tiny"], nc=1),
    dict(name="decoder_deformable_masked", kind="decoder", mode="deformable", seed=74, B=1, Q=40, shapes=syn.PYRAMIDS["tiny"], nc=1),
]


def masks_for(c):
    attn, pad = syn.make_denoising_masks(c["seed"], c["B"], c["Q"], syn.level_sizes(c["shapes"]))
    if c.get("float_mask"):   # additive float mask: nn.MultiheadAttention adds it to the scores
        attn = torch.zeros(attn.shape).masked_fill(attn, float("-inf"))
    return attn, pad


def gen_masked(T):
    for c in MASKED_CASES:
        spec = syn.DecoderSpec(nc=c.get("nc", 1))
        sd = syn.make_decoder_state(spec, WEIGHT_SEED)
        shapes = [list(s) for s in c["shapes"]]
        attn, pad = masks_for(c)
        if c["kind"] == "layer":
            m = getattr(T, c["cls"])(spec.d_model, spec.n_heads, spec.d_ffn, 0.0, torch.nn.ReLU(), spec.n_levels,
                                     spec.n_points).eval()
            m.load_state_dict(syn.sub_state(sd, "layers.1."))
            q, refer, feats, qpos = syn.make_module_inputs(c["seed"], c["B"], c["Q"], spec.d_model, c["shapes"], 4, 1)
            with torch.no_grad():
                o = m(q, refer[:, :, 0], feats, shapes, pad, attn, qpos)
            save(c["name"], dict(c, shapes=shapes, weight_seed=WEIGHT_SEED, checksum=syn.checksum(q, refer, feats, qpos)), out=o)
        else:
            dec, bbox, score, pos = build_ref_decoder(T, spec, sd, c["mode"])
            embed, refer, feats, qpos = syn.make_decoder_inputs(c["seed"], c["B"], c["Q"], spec.d_model, c["shapes"])
            with torch.no_grad():
                if c["mode"] == "motr":
                    b, s_, o = dec(embed, refer, feats, shapes, bbox, score, pos, attn_mask=attn, padding_mask=pad,
                                   track_query_embed=qpos)
                else:
                    b, s_ = dec(embed, refer, feats, shapes, bbox, score, pos, attn_mask=attn, padding_mask=pad)
                    o = torch.zeros(0)
            save(c["name"], dict(c, shapes=shapes, weight_seed=WEIGHT_SEED,
                                 checksum=syn.checksum(embed, refer, feats, qpos)), boxes=b, scores=s_, hs=o)


def build_ref_decoder(T, spec, sd, mode):
    layer_cls = T.MOTRDecoderLayer if mode == "motr" else T.DeformableTransformerDecoderLayer
    layer = layer_cls(spec.d_model, spec.n_heads, spec.d_ffn, 0.0, torch.nn.ReLU(), spec.n_levels, spec.n_points)
    dec_cls = T.MOTRTransformerDecoder if mode == "motr" else T.DeformableTransformerDecoder
    dec = dec_cls(spec.d_model, layer, spec.n_layers).eval()
    dec.load_state_dict(syn.sub_state(sd, "layers.") and {k: v for k, v in sd.items() if k.startswith("layers.")})
    bbox = torch.nn.ModuleList([T.MLP(spec.d_model, spec.d_model, 4, 3) for _ in range(spec.n_layers)]).eval()
    bbox.load_state_dict(syn.sub_state(sd, "dec_bbox_head."))
    score = torch.nn.ModuleList([torch.nn.Linear(spec.d_model, spec.nc) for _ in range(spec.n_layers)]).eval()
    score.load_state_dict(syn.sub_state(sd, "dec_score_head."))
    pos = T.MLP(4, spec.pos_hidden, spec.d_model, 2).eval()
    pos.load_state_dict(syn.sub_state(sd, "query_pos_head."))
    return dec, bbox, score, pos


def gen_decoders(T):
    for c in DECODER_CASES:
        spec = syn.DecoderSpec(nc=c["nc"])
        sd = syn.make_decoder_state(spec, WEIGHT_SEED)
        dec, bbox, score, pos = build_ref_decoder(T, spec, sd, c["mode"])
        embed, refer, feats, qpos = syn.make_decoder_inputs(c["seed"], c["B"], c["Q"], spec.d_model, c["shapes"])
        shapes = [list(s) for s in c["shapes"]]
        with torch.no_grad():
            if c["mode"] == "motr":
                b, s, o = dec(embed, refer, feats, shapes, bbox, score, pos, track_query_embed=qpos)
            else:
                b, s = dec(embed, refer, feats, shapes, bbox, score, pos)
                o = torch.zeros(0)
        save(c["name"], dict(c, shapes=shapes, weight_seed=WEIGHT_SEED,
                             checksum=syn.checksum(embed, refer, feats, qpos)), boxes=b, scores=s, hs=o)


def gen_tracker(head, structures):
    """Drive the reference RuntimeTrackerBase over synthetic multi-frame inputs (repair R2 applied by
    this driver: Instances rebuilt with N = T + n_detect rows each frame)."""
    Instances = structures.Instances
    for name, seed, nd, n_frames, dup in (("tracker_a", 51, 40, 12, 0.3), ("tracker_b", 52, 120, 8, 0.5),
                                          ("tracker_empty", 53, 10, 4, 0.0)):
        g = torch.Generator().manual_seed(seed)
        trk = head.RuntimeTrackerBase()
        ids = torch.zeros(0, dtype=torch.long)
        dis = torch.zeros(0, dtype=torch.long)
        tboxes = torch.zeros(0, 4)
        rec = {}
        for t in range(n_frames):
            T_ = ids.shape[0]
            nb = torch.cat([torch.rand(nd, 2, generator=g) * 0.8 + 0.1, torch.rand(nd, 2, generator=g) * 0.2 + 0.03], 1)
            n_dup = int(dup * nd)
            if n_dup and T_ + nd > 1:  # near-duplicates of other boxes so the IoU>0.8 filter fires
                allb = torch.cat([tboxes, nb], 0)
                src = torch.randint(0, allb.shape[0], (n_dup,), generator=g)
                nb[:n_dup] = allb[src] * (1 + 0.02 * torch.randn(n_dup, 4, generator=g))
            boxes = torch.cat([tboxes * (1 + 0.01 * torch.randn(T_, 4, generator=g)), nb], 0).float()
            if name == "tracker_empty":
                scores = torch.rand(T_ + nd, generator=g) * 0.3  # nothing ever reaches 0.4
            else:
                scores = torch.rand(T_ + nd, generator=g)
                scores[:4] = torch.tensor([0.4, 0.5, 0.39999998, 0.49999997])[: min(4, T_ + nd)]
            ids_in = torch.cat([ids, torch.full((nd,), -1, dtype=torch.long)])
            dis_in = torch.cat([dis, torch.zeros(nd, dtype=torch.long)])
            inst = Instances((1, 1))
            inst.scores = scores.clone()
            inst.obj_idxes = ids_in.clone()[:, None]
            inst.disappear_time = dis_in.clone()[:, None]
            inst.pred_boxes = boxes.clone()
            pre = (int(trk.max_obj_id), int(trk.max_obj_id_pre))  # ints: the reference mutates these tensors in place
            trk.update(inst)
            ids_out, dis_out = inst.obj_idxes[:, 0].clone(), inst.disappear_time[:, 0].clone()
            rec[f"scores_{t}"], rec[f"boxes_{t}"] = scores, boxes
            rec[f"ids_in_{t}"], rec[f"dis_in_{t}"] = ids_in, dis_in
            rec[f"ids_out_{t}"], rec[f"dis_out_{t}"] = ids_out, dis_out
            rec[f"counters_in_{t}"] = torch.tensor(pre)
            rec[f"counters_out_{t}"] = torch.tensor([int(trk.max_obj_id), int(trk.max_obj_id_pre)])
            act = ids_out >= 0
            ids, dis, tboxes = ids_out[act], dis_out[act], boxes[act]
        save(name, dict(seed=seed, n_detect=nd, n_frames=n_frames, source="head.py:1143-1283"), **rec)


def gen_qim(qim, structures):
    import argparse
    spec = syn.DecoderSpec()
    sd = syn.make_decoder_state(spec, WEIGHT_SEED)
    args = argparse.Namespace(merger_dropout=0.1, update_query_pos=False, random_drop=0.1, fp_ratio=0.3)
    m = qim.QueryInteractionModule(args, spec.d_model, spec.qim_hidden, spec.d_model * 2).eval()
    missing, unexpected = m.load_state_dict(syn.sub_state(sd, "track_embed."), strict=False)
    assert not unexpected and all(k.startswith("fsqm") for k in missing), (missing, unexpected)
    for name, seed, T_ in (("qim_a", 61, 23), ("qim_one", 62, 1)):
        g = torch.Generator().manual_seed(seed)
        inst = structures.Instances((1, 1))
        ref_pts = torch.randn(T_, 4, generator=g) * 1.5
        qpos = torch.randn(T_, spec.d_model, generator=g) * 0.5
        emb = torch.randn(T_, spec.d_model, generator=g)
        boxes = torch.rand(T_, 4, generator=g)
        boxes[0] = torch.tensor([0.0, 1.0, 1e-6, 0.5])[:4]
        inst.ref_pts, inst.query_pos, inst.output_embedding, inst.pred_boxes = ref_pts.clone(), qpos.clone(), emb, boxes
        with torch.no_grad():
            o = m._update_track_embedding(inst)
        save(name, dict(seed=seed, T=T_, weight_seed=WEIGHT_SEED), ref_pts=ref_pts, query_pos=qpos, out_embed=emb,
             pred_boxes=boxes, new_query_pos=o.query_pos, new_ref_pts=o.ref_pts)


def fsqm_inputs(seed, n_frames, N, d, n_det, clean):
    """Seeded per-frame inputs of FSQM.online_update (MOTR/models/fsqm.py:155-180): detect queries (embedding, score,
    box) and track queries (obj id, score, box). clean=True stays inside the domain where the reference class is
    self-consistent: every track id names a live slot whose id equals its index, a slot that received a score
    below out_threshold is not tracked again (it is freed consecutive_frames later and its slot re-used), and the id
    pool is never drained down to the -1 entries the reference appends to it; clean=False exercises everything else (recycled
    ids, ids >= N, id -1, low tracked scores)."""
    g = torch.Generator().manual_seed(seed)
    frames = []
    for t in range(n_frames):
        emb = torch.randn(n_det, d, generator=g)
        sc = torch.rand(n_det, generator=g)
        if clean:
            sc = torch.where(torch.rand(n_det, generator=g) < 0.15, 0.75 + 0.2 * sc, 0.6 * sc)
            if t == 0:
                sc[:3] = 0.9                                   # slots 0..2 are live from the first frame on
        box = torch.rand(n_det, 4, generator=g)
        k = int(torch.randint(2, 7, (1,), generator=g))
        if clean:
            k = 0 if t == 0 else k
            tid = torch.randint(0, 3 if t < 4 else 2, (k,), generator=g)   # only slots that are certainly live
            tsc = 0.3 + 0.7 * torch.rand(k, generator=g)
            if t == 4:                                         # slot 2 fades out: one low score, never tracked again
                tid = torch.cat([tid, torch.tensor([2])])
                tsc = torch.cat([tsc, torch.tensor([0.1])])
        else:
            tid = torch.randint(-1, N + 3, (k,), generator=g)
            tsc = torch.rand(k, generator=g)
        tbox = torch.rand(tid.shape[0], 4, generator=g)
        frames.append(dict(emb=emb, sc=sc, box=box, tid=tid, tsc=tsc, tbox=tbox))
    return frames


def gen_fsqm(structures):
    """The reference FSQM class itself (MOTR/models/fsqm.py:7-189) driven over seeded frames; its complete state after
    every online_update is the golden."""
    import importlib
    F = importlib.import_module("MOTR.models.fsqm")
    Instances = structures.Instances
    for name, seed, N, d, n_det, n_frames, clean in (("fsqm_clean", 91, 24, 8, 6, 10, True), ("fsqm_quirks", 92, 8, 8, 6, 16, False)):
        m = F.FSQM(N, d)
        rec = {}
        for t, fr in enumerate(fsqm_inputs(seed, n_frames, N, d, n_det, clean)):
            det = Instances((0, 0), output_embedding=fr["emb"], scores=fr["sc"], pred_boxes=fr["box"])
            trk = Instances((0, 0), obj_idxes=fr["tid"].view(-1, 1), scores=fr["tsc"], pred_boxes=fr["tbox"])
            out = m.online_update(det, trk)
            assert out is trk                                      # fsqm.py:180: the input comes back unchanged
            rec.update({f"ids_{t}": m.ids.clone(), f"conf_{t}": m.confidence.clone(), f"low_{t}": m.consecutive_low_frames.clone(),
                        f"boxes_{t}": m.bounding_boxes.clone(), f"mem_{t}": m.query_memory.clone(),
                        f"pool_{t}": torch.tensor(m.global_id_pool, dtype=torch.long)})
        save(name, dict(seed=seed, N=N, d=d, n_det=n_det, n_frames=n_frames, clean=clean, source="MOTR/models/fsqm.py:7-189"), **rec)


def gen_selection(head):
    """Encoder-side query selection of the reference MYDecoder (head.py:1012-1029, 993-1010, 1031-1113), eval mode,
    first-frame path (no carried tracks): feats, detect embeddings, logit-space boxes, encoder scores."""
    for name, pyr, B, nq, nc, seed in (("select_tiny", "tiny", 2, 48, 1, 81), ("select_c1", "C1", 1, 300, 1, 82),
                                       ("select_kitti_nc5", "KITTI", 1, 300, 5, 83)):
        spec = syn.DecoderSpec(nc=nc)
        shapes = [list(x) for x in syn.PYRAMIDS[pyr]]
        sd = syn.make_selector_state(spec, (256, 512, 512), WEIGHT_SEED)
        maps = syn.make_pyramid_maps(seed, B, shapes)
        m = head.MYDecoder(nc=nc, ch=(256, 512, 512), nq=nq)
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and all(not k.startswith(("input_proj", "enc_")) for k in missing), (missing, unexpected)
        m.eval()
        with torch.no_grad():
            feats, shp = m._get_encoder_input(maps)
            embed, refer, enc_bboxes, enc_scores, _, query_pos = m._get_decoder_input(feats, shp, None, None, None,
                                                                                      is_first=True)
        assert shp == shapes
        # feats is [B, Lv, 256]: the full tensor for the tiny case, every 37th row (+ a checksum) otherwise
        step = 1 if pyr == "tiny" else 37
        save(name, dict(pyramid=pyr, B=B, nq=nq, nc=nc, seed=seed, weight_seed=WEIGHT_SEED, shapes=shapes,
                        checksum=syn.checksum(*maps), feats_row_step=step, feats_checksum=syn.checksum(feats),
                        source="head.py:1012-1113"),
             feats=feats[:, ::step].contiguous(), embed=embed, refer=refer, enc_scores=enc_scores, query_pos=query_pos)


def gen_formats():
    """Output text of the reference's own writers on a seeded track table:
    Detector.write_results (MOTR/submit_dance.py:410-419) and TrackResults.save_txt
    (ultralytics/engine/results.py:475-512)."""
    import importlib
    import os
    import tempfile
    sys.path.insert(0, str(ref_loader.REF / "MOTR"))
    sd = importlib.import_module("MOTR.submit_dance")
    res = importlib.import_module("ultralytics.engine.results")
    g = torch.Generator().manual_seed(61)
    n, img_w, img_h = 40, 1088, 608
    table = torch.zeros(n, 9)
    table[:, 1] = torch.arange(n) // 8                       # 5 frames
    table[:, 2] = torch.randint(-1, 30, (n,), generator=g).float()
    table[:, 3:5] = torch.rand(n, 2, generator=g) * 0.8 + 0.1
    table[:, 5:7] = torch.rand(n, 2, generator=g) * 0.2 + 0.01
    table[:, 7] = torch.rand(n, generator=g)
    table[:, 8] = torch.randint(0, 5, (n,), generator=g).float()
    cx, cy, w, h = table[:, 3], table[:, 4], table[:, 5], table[:, 6]
    xyxy = torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], 1) * torch.tensor(
        [img_w, img_h, img_w, img_h], dtype=torch.float32)
    mot, txt, txt_conf = [], {}, {}
    for f in range(5):
        m = table[:, 1] == f
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "mot.txt")
            sd.Detector.write_results(p, f + 1, xyxy[m].numpy(), table[m, 2].long().numpy())
            mot.append(open(p).read() if os.path.exists(p) else "")
            boxes6 = torch.cat([xyxy[m], table[m, 7:8], table[m, 8:9]], 1)
            r = res.TrackResults(np.zeros((img_h, img_w, 3), np.uint8), "x.jpg", {i: str(i) for i in range(5)},
                                 boxes=boxes6, track_id=table[m, 2].long())
            for conf, store in ((False, txt), (True, txt_conf)):
                q = os.path.join(d, f"l{int(conf)}.txt")
                r.save_txt(q, save_conf=conf)
                store[f] = open(q).read()
    save("formats", dict(seed=61, img_w=img_w, img_h=img_h, mot="".join(mot), save_txt=txt, save_txt_conf=txt_conf,
                         source="MOTR/submit_dance.py:410-419; ultralytics/engine/results.py:475-512"), table=table)


def main():
    assert ref_loader.available(), "needs /root/reference"
    torch.set_num_threads(8)
    T, U = ref_loader.load_decoder_modules()
    print("decoder-side goldens (light loader)")
    gen_core(U)
    gen_core_grad(U)
    gen_posemb(T, U)
    gen_msda(T)
    gen_layers(T)
    gen_decoders(T)
    gen_masked(T)
    print("head-side goldens (full reference import with stubs)")
    head, qim, structures = ref_loader.load_full_reference()
    gen_kat0()
    gen_tracker(head, structures)
    gen_qim(qim, structures)
    gen_fsqm(structures)
    gen_selection(head)
    gen_formats()


if __name__ == "__main__":
    main()
