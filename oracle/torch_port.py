"""ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product package `moyolo_b200`.

CPU restatement (plain PyTorch fp32/fp64 tensor ops, functional, weights passed as a dict with the
reference's state_dict key names) of the reference's decoder hot path. Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it.

Parity status: PINNED. `oracle/make_golden.py` (run in the build container, where /root/reference
exists) executes the reference's own modules, imported unmodified, on seeded inputs and stores their
outputs under tests/golden/; tests/test_oracle_vs_golden.py checks every function below against
those files. The reference's own known-answer case MOTR/models/ops/test.py:21-60 is golden #0.

Each function cites the reference lines it restates (paths relative to the reference root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def inverse_sigmoid(x: Tensor, eps: float = 1e-5) -> Tensor:
    """ultralytics/nn/modules/utils.py:34-38 (== MOTR/util/misc.py:532-536)."""
    x = x.clamp(0, 1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def pos2posemb(pos: Tensor, num_pos_feats: int = 64, temperature: float = 10000) -> Tensor:
    """ultralytics/nn/modules/transformer.py:183-190 (== MOTR/models/qim.py:117-124)."""
    i = torch.arange(num_pos_feats, dtype=pos.dtype, device=pos.device)
    dim_t = temperature ** (2 * (i // 2) / num_pos_feats)
    e = (pos * (2 * math.pi))[..., None] / dim_t
    out = torch.stack((e[..., 0::2].sin(), e[..., 1::2].cos()), dim=-1)
    return out.flatten(-3)


def msda_core_gridsample(value: Tensor, shapes: Sequence[Sequence[int]], loc: Tensor, weights: Tensor) -> Tensor:
    """multi_scale_deformable_attn_pytorch, ultralytics/nn/modules/utils.py:41-78 — the same op
    sequence the reference executes (per-level NCHW view, F.grid_sample bilinear / zeros /
    align_corners=False, weighted sum over levels*points). This is the leg timed as cpu_baseline."""
    B, _, H, D = value.shape
    Q, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    grids = 2 * loc - 1
    sampled = []
    start = 0
    for lvl, (h, w) in enumerate(shapes):
        v = value[:, start:start + h * w]                                    # [B, hw, H, D]
        start += h * w
        v = v.permute(0, 2, 3, 1).reshape(B * H, D, h, w)                     # [B*H, D, h, w]
        g = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(B * H, Q, P, 2)
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    s = torch.stack(sampled, dim=-2).reshape(B * H, D, Q, L * P)              # [B*H, D, Q, L*P]
    a = weights.permute(0, 2, 1, 3, 4).reshape(B * H, 1, Q, L * P)
    out = (s * a).sum(-1)                                                     # [B*H, D, Q]
    return out.reshape(B, H * D, Q).transpose(1, 2).contiguous()


def msda_core_gather(value: Tensor, shapes: Sequence[Sequence[int]], loc: Tensor, weights: Tensor) -> Tensor:
    """Same function as above, restated from the arithmetic instead of through grid_sample:
    pixel = loc*size - 0.5, floor, four bounds-checked corners with zero padding
    (MOTR/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:33-84,285-291). Independent cross-check."""
    B, Lv, H, D = value.shape
    Q, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    out = torch.zeros(B, Q, H, D, dtype=value.dtype)
    bidx = torch.arange(B).view(B, 1, 1, 1).expand(B, Q, H, P)
    hidx = torch.arange(H).view(1, 1, H, 1).expand(B, Q, H, P)
    start = 0
    for lvl, (h, w) in enumerate(shapes):
        x = loc[:, :, :, lvl, :, 0] * w - 0.5                                 # [B, Q, H, P]
        y = loc[:, :, :, lvl, :, 1] * h - 0.5
        x0, y0 = torch.floor(x), torch.floor(y)
        fx, fy = x - x0, y - y0
        a = weights[:, :, :, lvl]                                             # [B, Q, H, P]
        for dy, dx, cw in ((0, 0, (1 - fy) * (1 - fx)), (0, 1, (1 - fy) * fx), (1, 0, fy * (1 - fx)),
                           (1, 1, fy * fx)):
            xi, yi = (x0 + dx).long(), (y0 + dy).long()
            ok = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
            pos = start + yi.clamp(0, h - 1) * w + xi.clamp(0, w - 1)
            v = value[bidx, pos, hidx]                                        # [B, Q, H, P, D]
            out += ((cw * a * ok)[..., None] * v).sum(3)
        start += h * w
    return out.reshape(B, Q, H * D)


def msdeform_attn_forward(p: Dict[str, Tensor], query: Tensor, refer_bbox: Tensor, value: Tensor,
                          shapes: Sequence[Sequence[int]], n_heads: int, n_levels: int, n_points: int,
                          value_mask: Optional[Tensor] = None, core=msda_core_gridsample,
                          my_softmax: bool = False) -> Tensor:
    """MSDeformAttn.forward, ultralytics/nn/modules/transformer.py:246-287.
    p keys: sampling_offsets/attention_weights/value_proj/output_proj .weight/.bias (:214-217).
    my_softmax: MOTRMSDeformAttn's optional normalisation exp(x) / (1 + sum exp(x)) (`_custom_softmax`,
    transformer.py:239-244, selected at :369-371). UNPINNED: the reference method lacks `self`, so the reference
    itself raises when the flag is set; this restates the arithmetic of :241-244."""
    B, Q, C = query.shape
    Lv = value.shape[1]
    assert sum(h * w for h, w in shapes) == Lv                                # :262
    v = F.linear(value, p["value_proj.weight"], p["value_proj.bias"])          # :264
    if value_mask is not None:
        v = v.masked_fill(value_mask[..., None], 0.0)                          # :265-266 (True rows -> 0)
    v = v.view(B, Lv, n_heads, C // n_heads)
    off = F.linear(query, p["sampling_offsets.weight"], p["sampling_offsets.bias"])
    off = off.view(B, Q, n_heads, n_levels, n_points, 2)                       # :268
    att = F.linear(query, p["attention_weights.weight"], p["attention_weights.bias"])
    att = att.view(B, Q, n_heads, n_levels * n_points)
    if my_softmax:
        e = torch.exp(att)                                                     # :241
        att = e / (1 + e.sum(-1, keepdim=True))                                # :242-244
    else:
        att = F.softmax(att, -1)                                               # :269-271
    att = att.view(B, Q, n_heads, n_levels, n_points)
    d = refer_bbox.shape[-1]
    if d == 2:                                                                 # :276-279
        norm = torch.tensor([[w, h] for h, w in shapes], dtype=query.dtype)
        locs = refer_bbox[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    elif d == 4:                                                               # :280-282
        locs = refer_bbox[:, :, None, :, None, :2] + off / n_points * refer_bbox[:, :, None, :, None, 2:] * 0.5
    else:
        raise ValueError(f"Last dim of reference_points must be 2 or 4, but got {d}.")
    out = core(v, shapes, locs, att)                                           # :285
    return F.linear(out, p["output_proj.weight"], p["output_proj.bias"])       # :286


def mha_forward(p: Dict[str, Tensor], q_in: Tensor, k_in: Tensor, v_in: Tensor, n_heads: int,
                attn_mask: Optional[Tensor] = None, prefix: str = "self_attn.") -> Tensor:
    """nn.MultiheadAttention as called at transformer.py:637-641 / qim.py:276 (batch-first here).
    q/k/v projected with the three row blocks of in_proj_weight, q scaled by 1/sqrt(Dh), softmax over
    keys, out_proj. q_in/k_in/v_in: [B, N, C]."""
    B, N, C = q_in.shape
    W, b = p[prefix + "in_proj_weight"], p[prefix + "in_proj_bias"]
    Dh = C // n_heads
    q = F.linear(q_in, W[:C], b[:C]).view(B, N, n_heads, Dh).transpose(1, 2)
    k = F.linear(k_in, W[C:2 * C], b[C:2 * C]).view(B, -1, n_heads, Dh).transpose(1, 2)
    v = F.linear(v_in, W[2 * C:], b[2 * C:]).view(B, -1, n_heads, Dh).transpose(1, 2)
    s = (q * (1.0 / math.sqrt(Dh))) @ k.transpose(-1, -2)
    if attn_mask is not None:
        s = s.masked_fill(attn_mask, float("-inf")) if attn_mask.dtype == torch.bool else s + attn_mask
    o = (F.softmax(s, -1) @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(o, p[prefix + "out_proj.weight"], p[prefix + "out_proj.bias"])


def _ln(p, name, x):
    return F.layer_norm(x, (x.shape[-1],), p[name + ".weight"], p[name + ".bias"], 1e-5)


def _sub(p: Dict[str, Tensor], prefix: str) -> Dict[str, Tensor]:
    return {k[len(prefix):]: v for k, v in p.items() if k.startswith(prefix)}


def decoder_layer_forward(p: Dict[str, Tensor], embed: Tensor, refer_bbox: Tensor, feats: Tensor, shapes,
                          n_heads: int, n_levels: int, n_points: int, padding_mask=None, attn_mask=None,
                          query_pos: Optional[Tensor] = None, core=msda_core_gridsample) -> Tensor:
    """DeformableTransformerDecoderLayer.forward (transformer.py:431-450) == MOTRDecoderLayer.forward
    (:627-652) with dropout 0: post-norm self-attention, deformable cross-attention, ReLU FFN."""
    qk = embed if query_pos is None else embed + query_pos
    embed = _ln(p, "norm1", embed + mha_forward(p, qk, qk, embed, n_heads, attn_mask))          # :637-641
    xq = embed if query_pos is None else embed + query_pos
    t = msdeform_attn_forward(_sub(p, "cross_attn."), xq, refer_bbox.unsqueeze(2), feats, shapes, n_heads,
                              n_levels, n_points, padding_mask, core)                           # :644-645
    embed = _ln(p, "norm2", embed + t)                                                          # :646-647
    t = F.linear(F.relu(F.linear(embed, p["linear1.weight"], p["linear1.bias"])), p["linear2.weight"],
                 p["linear2.bias"])                                                             # :576-577
    return _ln(p, "norm3", embed + t)                                                           # :578-579


def mlp_forward(p: Dict[str, Tensor], x: Tensor, n: int, prefix: str) -> Tensor:
    """MLP.forward, transformer.py:158-161."""
    for i in range(n):
        x = F.linear(x, p[f"{prefix}layers.{i}.weight"], p[f"{prefix}layers.{i}.bias"])
        if i < n - 1:
            x = F.relu(x)
    return x


def decoder_forward(sd: Dict[str, Tensor], embed: Tensor, refer_logit: Tensor, feats: Tensor, shapes,
                    n_heads: int, n_levels: int, n_points: int, n_layers: int, mode: str = "motr",
                    query_pos: Optional[Tensor] = None, eval_idx: int = -1, attn_mask=None, padding_mask=None,
                    core=msda_core_gridsample):
    """Eval-mode DeformableTransformerDecoder.forward (transformer.py:465-510, mode="deformable":
    query_pos = pos_mlp(refer) per layer, :491) and MOTRTransformerDecoder.forward (:676-728,
    mode="motr": one fixed track_query_embed, also returns the last embedding).
    sd keys: layers.{i}.*, dec_bbox_head.{i}.layers.{j}.*, dec_score_head.{i}.*, query_pos_head.layers.{j}.*"""
    eval_idx = eval_idx if eval_idx >= 0 else n_layers + eval_idx
    out = embed
    refer = refer_logit.sigmoid()                                                               # :482/:690
    boxes = scores = None
    for i in range(n_layers):
        pos = query_pos if mode == "motr" else mlp_forward(sd, refer, 2, "query_pos_head.")
        out = decoder_layer_forward(_sub(sd, f"layers.{i}."), out, refer, feats, shapes, n_heads, n_levels,
                                    n_points, padding_mask, attn_mask, pos, core)
        refined = torch.sigmoid(mlp_forward(sd, out, 3, f"dec_bbox_head.{i}.") + inverse_sigmoid(refer))  # :709
        if i == eval_idx:                                                                       # :717-721
            scores = F.linear(out, sd[f"dec_score_head.{i}.weight"], sd[f"dec_score_head.{i}.bias"])
            boxes = refined
            break
        refer = refined
    return boxes[None], scores[None], out


def qim_update(sd: Dict[str, Tensor], ref_pts: Tensor, query_pos: Tensor, out_embed: Tensor, pred_boxes: Tensor,
               n_heads: int = 8, prefix: str = "track_embed."):
    """QueryInteractionModule._update_track_embedding, MOTR/models/qim.py:251-301 with
    update_query_pos=False (MOTR/main.py:170) and dropout inactive (eval).
    ref_pts [T,4] logits, query_pos [T,C], out_embed [T,C], pred_boxes [T,4] -> (query_pos', ref_pts')."""
    p = _sub(sd, prefix)
    if ref_pts.shape[0] == 0:
        return query_pos, ref_pts
    qpos = pos2posemb(ref_pts)                                                                  # :255
    qk = (qpos + out_embed)[None]                                                               # :271
    tgt = out_embed
    tgt = _ln(p, "norm1", tgt + mha_forward(p, qk, qk, tgt[None], n_heads)[0])                  # :276-278
    t2 = F.linear(F.relu(F.linear(tgt, p["linear1.weight"], p["linear1.bias"])), p["linear2.weight"],
                  p["linear2.bias"])                                                            # :280
    tgt = _ln(p, "norm2", tgt + t2)                                                             # :281-282
    f2 = F.linear(F.relu(F.linear(tgt, p["linear_feat1.weight"], p["linear_feat1.bias"])),
                  p["linear_feat2.weight"], p["linear_feat2.bias"])                             # :290
    new_query_pos = _ln(p, "norm_feat", query_pos + f2)                                         # :294-298
    return new_query_pos, inverse_sigmoid(pred_boxes[:, :4].detach().clone())                   # :300


def generate_anchors(shapes, grid_size: float = 0.05, dtype=torch.float32, eps: float = 1e-2):
    """ultralytics/nn/modules/head.py:993-1010: cell-centre anchors (cx, cy, w, h) with w = h = 0.05 * 2^level in
    logit space, `inf` where any coordinate falls outside (eps, 1 - eps). Returns ([1, Lv, 4], [1, Lv, 1] bool)."""
    anchors = []
    for i, (h, w) in enumerate(shapes):
        gy, gx = torch.meshgrid(torch.arange(end=h, dtype=dtype), torch.arange(end=w, dtype=dtype), indexing="ij")
        grid_xy = torch.stack([gx, gy], -1)
        valid_wh = torch.tensor([h, w], dtype=dtype)  # sic: the reference divides (x, y) by (h, w), head.py:1000-1001
        grid_xy = (grid_xy.unsqueeze(0) + 0.5) / valid_wh
        wh = torch.ones_like(grid_xy) * grid_size * (2.0 ** i)
        anchors.append(torch.cat([grid_xy, wh], -1).view(-1, h * w, 4))
    anchors = torch.cat(anchors, 1)
    valid = ((anchors > eps) * (anchors < 1 - eps)).all(-1, keepdim=True)
    anchors = torch.log(anchors / (1 - anchors))
    anchors = anchors.masked_fill(~valid, float("inf"))
    return anchors, valid


def encoder_input(sd: Dict[str, Tensor], maps) -> tuple:
    """head.py:1012-1029 `_get_encoder_input`: per level 1x1 conv (no bias) + BatchNorm2d (eval, eps 1e-5,
    head.py:839-840), flattened to [B, H*W, C] and concatenated over levels."""
    feats, shapes = [], []
    for l, m in enumerate(maps):
        p = f"input_proj.{l}."
        y = F.conv2d(m, sd[p + "0.weight"])
        y = F.batch_norm(y, sd[p + "1.running_mean"], sd[p + "1.running_var"], sd[p + "1.weight"], sd[p + "1.bias"],
                         False, 0.0, 1e-5)
        h, w = y.shape[2:]
        feats.append(y.flatten(2).permute(0, 2, 1))
        shapes.append([int(h), int(w)])
    return torch.cat(feats, 1), shapes


def query_selection(sd: Dict[str, Tensor], feats: Tensor, shapes, num_queries: int) -> Dict[str, Tensor]:
    """head.py:1031-1113 `_get_decoder_input`, the detect-query part (is_first / no carried tracks, eval, no
    denoising, learnt_init_query False): enc_output on valid_mask * feats (:1039), score head (:1041), bbox head +
    anchors (:1044), top-k by max-class logit (:1048), gathers (:1053, :1092, :1104)."""
    bs = feats.shape[0]
    anchors, valid = generate_anchors(shapes, dtype=feats.dtype)
    x = valid * feats
    features = F.layer_norm(F.linear(x, sd["enc_output.0.weight"], sd["enc_output.0.bias"]), (x.shape[-1],),
                            sd["enc_output.1.weight"], sd["enc_output.1.bias"], 1e-5)
    scores = F.linear(features, sd["enc_score_head.weight"], sd["enc_score_head.bias"])
    bboxes = mlp_forward(sd, features, 3, "enc_bbox_head.") + anchors
    topk = torch.topk(scores.max(-1).values, num_queries, dim=1).indices
    bi = torch.arange(bs).unsqueeze(-1).repeat(1, num_queries).view(-1)
    ti = topk.view(-1)
    return {"features": features, "scores": scores, "topk": topk,
            "refer": bboxes[bi, ti].view(bs, num_queries, -1), "embed": features[bi, ti].view(bs, num_queries, -1),
            "enc_scores": scores[bi, ti].view(bs, num_queries, -1)}
