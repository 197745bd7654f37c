"""Per-frame tracking engine: query assembly -> 6-layer decoder -> heads -> track-query update.

Mirrors the frame semantics of MOTRTrack.forward / _post_process_single_image
(ultralytics/nn/modules/head.py:191-239, 300-321, 492-497) with carried tracks as specified by
"O3" in SURVEY.md §8(c) / oracle/tracker_port.py:
  * queries = [carried tracks ; detect queries] per sequence (head.py:1056-1064, 1108-1109);
    track content embedding = denoising_class_embed[argmax previous logits] (head.py:888-900);
  * after the decode every field has N = T + n_detect rows, ids = cat(prev ids, -1), disappear =
    cat(prev, 0) (repair R2, upstream MOTR/models/motr.py:569-574);
  * RuntimeTrackerBase.update on device (moyolo_track_assign), active selection (ids >= 0,
    MOTR/models/qim.py:184-187) and QIM._update_track_embedding (qim.py:251-301) produce the next
    frame's track queries.
Several independent sequences run in lock-step as one ragged batch; frames of one
sequence are strictly ordered. All state lives on the device; the host reads back one int32 per
sequence and frame (the active-track count that sizes the next frame's launch).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import _lib, executor as ex, ops
from .synthetic import DecoderSpec, level_sizes


class _Holder(torch.nn.Module):
    """Minimal parameter holders so executor packs can be built straight from a state_dict."""


def _linear_from(sd, prefix, dev):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    lin = torch.nn.Linear(w.shape[1], w.shape[0])
    lin.weight.data, lin.bias.data = w.clone(), b.clone()
    return lin.to(dev)


class DecoderWeights:
    """All weights of the frame path, packed once for one precision."""

    def __init__(self, sd: Dict[str, torch.Tensor], spec: DecoderSpec, device, precision: str):
        from . import modules as M
        self.spec, self.precision = spec, precision
        dt = ex.lp_dtype(precision)
        self.dt = dt
        layer = M.MOTRDecoderLayer(spec.d_model, spec.n_heads, spec.d_ffn, 0.0, torch.nn.ReLU(), spec.n_levels,
                                   spec.n_points)
        dec = M.MOTRTransformerDecoder(spec.d_model, layer, spec.n_layers)
        dec.load_state_dict({k: v for k, v in sd.items() if k.startswith("layers.")})
        dec = dec.to(device).eval()
        self.layers = [ex.LayerPack(l, dt) for l in dec.layers]
        self.value_proj = ex.ValueProjPack(dec.layers, dt)
        self.bbox = []
        for i in range(spec.n_layers):
            mlp = M.MLP(spec.d_model, spec.d_model, 4, 3)
            mlp.load_state_dict({k[len(f"dec_bbox_head.{i}."):]: v for k, v in sd.items()
                                 if k.startswith(f"dec_bbox_head.{i}.")})
            self.bbox.append(ex.MlpPack(mlp.to(device), dt))
        last = spec.n_layers - 1
        self.score_w = sd[f"dec_score_head.{last}.weight"].to(device).float().contiguous()
        self.score_b = sd[f"dec_score_head.{last}.bias"].to(device).float().contiguous()
        self.class_embed = sd["denoising_class_embed.weight"].to(device).float().contiguous()
        # QIM (MOTR/models/qim.py:85-105)
        q = "track_embed."
        C = spec.d_model
        f = lambda t: t.to(device).float().contiguous()  # noqa: E731
        w = lambda t: t.to(device).to(dt).contiguous()  # noqa: E731
        self.qim = {
            "qk_w": w(sd[q + "self_attn.in_proj_weight"][:2 * C]), "qk_b": f(sd[q + "self_attn.in_proj_bias"][:2 * C]),
            "v_w": w(sd[q + "self_attn.in_proj_weight"][2 * C:]), "v_b": f(sd[q + "self_attn.in_proj_bias"][2 * C:]),
            "o_w": w(sd[q + "self_attn.out_proj.weight"]), "o_b": f(sd[q + "self_attn.out_proj.bias"]),
            "l1_w": w(sd[q + "linear1.weight"]), "l1_b": f(sd[q + "linear1.bias"]),
            "l2_w": w(sd[q + "linear2.weight"]), "l2_b": f(sd[q + "linear2.bias"]),
            "f1_w": w(sd[q + "linear_feat1.weight"]), "f1_b": f(sd[q + "linear_feat1.bias"]),
            "f2_w": w(sd[q + "linear_feat2.weight"]), "f2_b": f(sd[q + "linear_feat2.bias"]),
        }
        for n in ("norm1", "norm2", "norm_feat"):
            self.qim[n] = (f(sd[q + n + ".weight"]), f(sd[q + n + ".bias"]))


def decode_frame(W: DecoderWeights, x_f32, refer_logit, pos, feats_lp, shapes, n_seq: int, ro, ro_host):
    """MOTRTransformerDecoder.forward in eval mode (transformer.py:676-728) over a ragged batch.

    x_f32 [R, C], refer_logit [R, 4], pos [R, C] fp32; feats_lp [n_seq*Lv, C] in the GEMM dtype.
    Returns boxes [R,4], logits [R,nc], scores [R], labels [R] (int32), hs [R,C] fp32.
    """
    dt, spec = W.dt, W.spec
    C, n_l = spec.d_model, spec.n_layers
    Lv = feats_lp.shape[0] // n_seq
    values = ops.linear(feats_lp, W.value_proj.w, W.value_proj.b, out_dtype=dt, engine=ex._GEMM_ENGINE)
    values = values.view(n_seq, Lv, n_l * C)
    refer = ops.sigmoid(refer_logit)
    x_lp = x_f32 if dt == torch.float32 else ops.add_cast(x_f32, None, dt)
    xq_lp = ops.add_cast(x_f32, pos, dt)
    R = x_f32.shape[0]
    for i, pk in enumerate(W.layers):
        pos_next = pos if i + 1 < n_l else None
        x_f32, x_lp, xq_lp = ex.run_layer(pk, x_f32, x_lp, xq_lp, refer.view(R, 1, 4),
                                          values[:, :, i * C:(i + 1) * C], shapes, n_seq, ro, ro_host, False, None,
                                          pos, pos_next, dt)
        refer = ex.bbox_head(W.bbox[i], x_lp, refer)
    logits, scores, labels = ops.score_head(x_lp, W.score_w, W.score_b)
    return refer, logits, scores, labels, x_f32


def qim_update(W: DecoderWeights, ref_pts, query_pos, out_embed, pred_boxes):
    """QueryInteractionModule._update_track_embedding (MOTR/models/qim.py:251-301) for one sequence's
    active tracks: ref_pts [T,4] logits, query_pos/out_embed [T,C] fp32, pred_boxes [T,4].
    Returns (new_query_pos [T,C] fp32, new_ref_pts [T,4] logits)."""
    T, C = out_embed.shape
    dt, q, H = W.dt, W.qim, 8  # nn.MultiheadAttention(dim_in, 8, ...) qim.py:88
    eng = ex._GEMM_ENGINE
    dev = out_embed.device
    qpos = ops.pos2posemb(ref_pts)                                   # qim.py:255
    qk_lp = ops.add_cast(qpos, out_embed, dt)                        # :271
    tgt_lp = out_embed if dt == torch.float32 else ops.add_cast(out_embed, None, dt)
    qkv = torch.empty(T, 3 * C, dtype=dt, device=dev)
    ops.linear(qk_lp, q["qk_w"], q["qk_b"], out=qkv[:, :2 * C], engine=eng)
    ops.linear(tgt_lp, q["v_w"], q["v_b"], out=qkv[:, 2 * C:], engine=eng)
    ro_host = [0, T]
    ro = torch.tensor(ro_host, dtype=torch.int32, device=dev)
    att = ops.self_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], ro, ro_host, H)
    t2 = ops.linear(att, q["o_w"], q["o_b"], out_dtype=torch.float32, engine=eng)
    tgt_f32, tgt_lp, _ = ops.add_layernorm(t2, out_embed, *q["norm1"], 1e-5, True, True, dt)     # :277-278
    h = ops.linear(tgt_lp, q["l1_w"], q["l1_b"], relu=True, engine=eng)
    t3 = ops.linear(h, q["l2_w"], q["l2_b"], out_dtype=torch.float32, engine=eng)                 # :280
    _, tgt_lp, _ = ops.add_layernorm(t3, tgt_f32, *q["norm2"], 1e-5, False, True, dt)             # :281-282
    h = ops.linear(tgt_lp, q["f1_w"], q["f1_b"], relu=True, engine=eng)
    f2 = ops.linear(h, q["f2_w"], q["f2_b"], out_dtype=torch.float32, engine=eng)                 # :290
    new_qpos, _, _ = ops.add_layernorm(f2, query_pos, *q["norm_feat"], 1e-5, True, False, dt)     # :294-298
    return new_qpos, ops.inverse_sigmoid(pred_boxes)                                              # :300


class TrackEngine:
    """Lock-step tracker for `n_seq` independent sequences on one GPU."""

    def __init__(self, sd, spec: DecoderSpec, shapes, device, precision: str = "bf16", n_detect: int = 300,
                 n_seq: int = 1, score_thresh=0.4, filter_thresh=0.5, miss_tolerance=5, iou_thresh=0.8,
                 weights: Optional[DecoderWeights] = None):
        self.dev = torch.device(device)
        self.spec, self.shapes, self.n_detect, self.n_seq = spec, [list(s) for s in shapes], n_detect, n_seq
        self.Lv = level_sizes(shapes)
        self.W = weights or DecoderWeights(sd, spec, self.dev, precision)
        self.thr = (score_thresh, filter_thresh, miss_tolerance, iou_thresh)
        self.frame_idx = 0
        self.reset()

    # ---- state -------------------------------------------------------------------------------
    def reset(self, seq: Optional[int] = None):
        """is_first semantics (head.py:199-205): drop all tracks, fresh ID counters."""
        C, dev = self.spec.d_model, self.dev
        if seq is None:
            self.t_ref = [torch.zeros(0, 4, device=dev) for _ in range(self.n_seq)]
            self.t_qpos = [torch.zeros(0, C, device=dev) for _ in range(self.n_seq)]
            self.t_label = [torch.zeros(0, dtype=torch.int32, device=dev) for _ in range(self.n_seq)]
            self.t_ids = [torch.zeros(0, dtype=torch.int64, device=dev) for _ in range(self.n_seq)]
            self.t_dis = [torch.zeros(0, dtype=torch.int64, device=dev) for _ in range(self.n_seq)]
            self.counters = torch.zeros(self.n_seq, 2, dtype=torch.int64, device=dev)
            self.frame_idx = 0
        else:
            self.t_ref[seq] = torch.zeros(0, 4, device=dev)
            self.t_qpos[seq] = torch.zeros(0, C, device=dev)
            self.t_label[seq] = torch.zeros(0, dtype=torch.int32, device=dev)
            self.t_ids[seq] = torch.zeros(0, dtype=torch.int64, device=dev)
            self.t_dis[seq] = torch.zeros(0, dtype=torch.int64, device=dev)
            self.counters[seq].zero_()

    def n_tracks(self) -> List[int]:
        return [int(t.shape[0]) for t in self.t_ids]

    # ---- one frame ----------------------------------------------------------------------------
    def step(self, feats: torch.Tensor, det_embed: torch.Tensor, det_refer: torch.Tensor) -> List[Dict[str, torch.Tensor]]:
        """feats [n_seq, Lv, C] (GEMM dtype or fp32), det_embed [n_seq, nd, C] fp32, det_refer
        [n_seq, nd, 4] fp32 logit-space boxes. Returns one dict per sequence with the N = T + nd rows
        of this frame: ids (int64, -1 = no object), boxes (cx,cy,w,h normalised), scores, labels."""
        W, dev, S, nd, C = self.W, self.dev, self.n_seq, self.n_detect, self.spec.d_model
        dt = W.dt
        assert feats.shape[0] == S and feats.shape[1] == self.Lv
        feats_lp = feats.reshape(S * self.Lv, C)
        if feats_lp.dtype != dt:
            feats_lp = ops.add_cast(feats_lp.float().contiguous(), None, dt)
        feats_lp = feats_lp.contiguous()
        det_pos = ops.pos2posemb(det_refer.reshape(S * nd, 4).contiguous()).view(S, nd, C)
        xs, rs, ps, ro_host = [], [], [], [0]
        for s in range(S):
            T = self.t_ids[s].shape[0]
            if T:
                xs.append(W.class_embed[self.t_label[s].long()])      # head.py:888-900
                rs.append(self.t_ref[s])
                ps.append(self.t_qpos[s])
            xs.append(det_embed[s])
            rs.append(det_refer[s])
            ps.append(det_pos[s])
            ro_host.append(ro_host[-1] + T + nd)
        x = torch.cat(xs, 0).float().contiguous()
        refer_logit = torch.cat(rs, 0).float().contiguous()
        pos = torch.cat(ps, 0).contiguous()
        ro = torch.tensor(ro_host, dtype=torch.int32, device=dev)

        boxes, logits, scores, labels, hs = decode_frame(W, x, refer_logit, pos, feats_lp, self.shapes, S, ro, ro_host)

        outs = []
        st, ft, mt, it = self.thr
        for s in range(S):
            a, b = ro_host[s], ro_host[s + 1]
            N = b - a
            ids = torch.cat([self.t_ids[s], torch.full((nd,), -1, dtype=torch.int64, device=dev)])      # R2
            dis = torch.cat([self.t_dis[s], torch.zeros(nd, dtype=torch.int64, device=dev)])
            ws = torch.empty(ops.track_workspace_bytes(N), dtype=torch.uint8, device=dev)
            ops.track_assign(scores[a:b], boxes[a:b], ids, dis, self.counters[s], ws, st, ft, mt, it)
            outs.append({"ids": ids, "boxes": boxes[a:b], "scores": scores[a:b], "labels": labels[a:b],
                         "logits": logits[a:b]})
            # active selection + gather of everything the next frame needs
            fields = [refer_logit[a:b], pos[a:b], hs[a:b], boxes[a:b], labels[a:b], ids, dis]
            dst = [torch.empty_like(f) for f in fields]
            n_act = torch.empty(1, dtype=torch.int32, device=dev)
            idx = torch.empty(N, dtype=torch.int32, device=dev)
            ops.track_compact(ids, fields, dst, n_act, idx)
            k = int(n_act.item())  # the one host read-back per sequence and frame
            if k == 0:
                self.reset_tracks_only(s)
                continue
            c_ref, c_pos, c_hs, c_box, c_lab, c_ids, c_dis = (d[:k] for d in dst)
            new_qpos, new_ref = qim_update(W, c_ref, c_pos, c_hs, c_box)
            self.t_ref[s], self.t_qpos[s], self.t_label[s] = new_ref, new_qpos, c_lab
            self.t_ids[s], self.t_dis[s] = c_ids, c_dis
        self.frame_idx += 1
        return outs

    def reset_tracks_only(self, s: int):
        C, dev = self.spec.d_model, self.dev
        self.t_ref[s] = torch.zeros(0, 4, device=dev)
        self.t_qpos[s] = torch.zeros(0, C, device=dev)
        self.t_label[s] = torch.zeros(0, dtype=torch.int32, device=dev)
        self.t_ids[s] = torch.zeros(0, dtype=torch.int64, device=dev)
        self.t_dis[s] = torch.zeros(0, dtype=torch.int64, device=dev)


class SequenceTracker(TrackEngine):
    """Single-sequence convenience wrapper: 2-D inputs, one dict out."""

    def __init__(self, sd, spec, shapes, device, precision="bf16", n_detect=300, **kw):
        super().__init__(sd, spec, shapes, device, precision, n_detect, 1, **kw)

    def step(self, feats, det_embed, det_refer):  # type: ignore[override]
        return super().step(feats[None], det_embed[None], det_refer[None])[0]
