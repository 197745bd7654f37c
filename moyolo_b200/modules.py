"""Drop-in operator API of the reference decoder path, executed by libmoyolo_b200.

Same class names, constructor/forward signatures, parameter names/shapes (checkpoint compatible,
SURVEY.md §8(b)) and exceptions as ultralytics/nn/modules/transformer.py:149-161 (MLP), :183-190
(pos2posemb), :193-287 (MSDeformAttn), :290-391 (MOTRMSDeformAttn), :394-450
(DeformableTransformerDecoderLayer), :453-510 (DeformableTransformerDecoder), :515-652
(MOTRDecoderLayer), :663-728 (MOTRTransformerDecoder). `nn.Module`s here only own parameters;
every FLOP of `forward` runs in the CUDA extension (inference scope: no autograd through the ops).
"""
from __future__ import annotations

import copy
import math
from typing import Optional

import torch
import torch.nn as nn
from torch.nn.init import constant_, xavier_uniform_

from . import executor as ex
from . import ops

__all__ = ("MLP", "pos2posemb", "inverse_sigmoid", "multi_scale_deformable_attn", "MSDeformAttn",
           "MOTRMSDeformAttn", "DeformableTransformerDecoderLayer", "DeformableTransformerDecoder",
           "MOTRDecoderLayer", "MOTRTransformerDecoder")


def _clones(module, n):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


def inverse_sigmoid(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """ultralytics/nn/modules/utils.py:34-38 on device (eps is fixed at 1e-5 in the kernel)."""
    if eps != 1e-5:
        raise NotImplementedError("inverse_sigmoid kernel is specialised for eps=1e-5")
    return ops.inverse_sigmoid(x.float())


def pos2posemb(pos: torch.Tensor, num_pos_feats: int = 64, temperature: int = 10000) -> torch.Tensor:
    """transformer.py:183-190."""
    return ops.pos2posemb(pos.float(), num_pos_feats, float(temperature)).to(pos.dtype)


def multi_scale_deformable_attn(value, value_spatial_shapes, sampling_locations, attention_weights):
    """Same contract as multi_scale_deformable_attn_pytorch (utils.py:41-78)."""
    shapes = [[int(h), int(w)] for h, w in (value_spatial_shapes.tolist() if torch.is_tensor(value_spatial_shapes)
                                             else value_spatial_shapes)]
    return ops.msda_sampled(value, shapes, sampling_locations, attention_weights)


class MLP(nn.Module):
    """transformer.py:149-161; parameter names `layers.{i}.weight/bias`."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        dt = ex.lp_dtype(ex.get_default_precision())
        pk = ex.cached_pack(self, "mlp", ex.MlpPack, dt)
        lead = x.shape[:-1]
        h = x.reshape(-1, x.shape[-1])
        h = h.contiguous() if h.dtype == dt else ops.add_cast(h.float().contiguous(), None, dt)
        for lin in pk.hidden:
            h = ops.linear(h, lin.w, lin.b, relu=True)
        y = ops.linear(h, pk.last_lp.w, pk.last_lp.b, out_dtype=torch.float32)
        return y.view(*lead, -1).to(x.dtype)


class MSDeformAttn(nn.Module):
    """Multi-scale deformable attention, transformer.py:193-287."""

    _my_softmax_active = False

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4, my_softmax=False):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError(f'd_model must be divisible by n_heads, but got {d_model} and {n_heads}')
        self.im2col_step = 64
        self.my_softmax = my_softmax
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self.precision: Optional[str] = None  # None -> executor default
        # True: forward() builds an autograd graph (training through the op): the four linears, the
        # softmax and the location arithmetic are torch ops on the GPU, the gather is MSDeformAttnFunction (forward AND
        # backward kernels of libmoyolo_b200), everything fp32. False (default): the fused inference kernels.
        self.differentiable = False
        self._reset_parameters()

    def _reset_parameters(self):
        # transformer.py:221-237: zero offset weights, a ring of per-head directions scaled by the
        # point index as the offset bias, zero attention logits, xavier value/output projections.
        constant_(self.sampling_offsets.weight.data, 0.)
        ang = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        ring = torch.stack([ang.cos(), ang.sin()], -1)
        ring = ring / ring.abs().max(-1, keepdim=True)[0]
        ring = ring.view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
        ring = ring * torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, -1, 1)
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(ring.reshape(-1))
        constant_(self.attention_weights.weight.data, 0.)
        constant_(self.attention_weights.bias.data, 0.)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.)

    def _dt(self):
        return ex.lp_dtype(self.precision or ex.get_default_precision())

    def forward(self, query, refer_bbox, value, value_shapes, value_mask=None):
        """query [bs, Q, C]; refer_bbox [bs, Q, 1|n_levels, 2|4]; value [bs, Lv, C]; value_shapes
        [[H, W], ...]; value_mask [bs, Lv] bool, True rows are zeroed (transformer.py:265-266)."""
        bs, len_q = query.shape[:2]
        len_v = value.shape[1]
        assert sum(s[0] * s[1] for s in value_shapes) == len_v
        num_points = refer_bbox.shape[-1]
        if num_points not in (2, 4):
            raise ValueError(f'Last dim of reference_points must be 2 or 4, but got {num_points}.')
        if torch.is_grad_enabled():
            if self.differentiable:
                return self._forward_autograd(query, refer_bbox, value, value_shapes, value_mask)
            if query.requires_grad or value.requires_grad or refer_bbox.requires_grad:
                raise RuntimeError("moyolo_b200.MSDeformAttn runs its inference kernels, which build no autograd graph, "
                                   "but an input requires grad: set module.differentiable = True for the training path")
        dt = self._dt()
        pk = ex.cached_pack(self, "msda", ex.MsdaPack, dt)
        C = self.d_model
        q2 = query.reshape(bs * len_q, C)
        v2 = value.reshape(bs * len_v, C)
        q_lp = q2.contiguous() if q2.dtype == dt else ops.add_cast(q2.float().contiguous(), None, dt)
        v_lp = v2.contiguous() if v2.dtype == dt else ops.add_cast(v2.float().contiguous(), None, dt)
        zero_rows = None
        if value_mask is not None:
            zero_rows = value_mask.reshape(-1).to(torch.uint8).contiguous()
        val = ops.linear(v_lp, pk.value.w, pk.value.b, out_dtype=dt, zero_rows=zero_rows, engine=ex._GEMM_ENGINE)
        refer = refer_bbox.reshape(bs * len_q, refer_bbox.shape[2], num_points).float().contiguous()
        out = ex.msda_forward(pk, q_lp, refer, val.view(bs, len_v, C), value_shapes, bs, None, dt)
        return out.view(bs, len_q, C).to(query.dtype)


    def _forward_autograd(self, query, refer_bbox, value, value_shapes, value_mask=None):
        """transformer.py:262-286 as a differentiable graph; the core (utils.py:41-78) is the library's gather with
        its own backward kernel (msda_ext.MSDeformAttnFunction)."""
        import torch.nn.functional as F
        from .msda_ext import MSDeformAttnFunction
        if not query.is_cuda:
            raise RuntimeError("Not implemented on the CPU")
        bs, len_q, C = query.shape
        len_v = value.shape[1]
        H, L, P = self.n_heads, self.n_levels, self.n_points
        v = F.linear(value.float(), self.value_proj.weight, self.value_proj.bias)
        if value_mask is not None:
            v = v.masked_fill(value_mask[..., None], 0.0)
        v = v.view(bs, len_v, H, C // H)
        q = query.float()
        off = F.linear(q, self.sampling_offsets.weight, self.sampling_offsets.bias).view(bs, len_q, H, L, P, 2)
        att = F.linear(q, self.attention_weights.weight, self.attention_weights.bias).view(bs, len_q, H, L * P)
        if self.my_softmax and type(self).__name__ == "MOTRMSDeformAttn":
            e = att.exp()
            att = e / (1.0 + e.sum(-1, keepdim=True))                  # transformer.py:239-244
        else:
            att = F.softmax(att, -1)
        att = att.view(bs, len_q, H, L, P)
        rb = refer_bbox.float()
        if rb.shape[-1] == 2:
            norm = torch.as_tensor([[w, h] for h, w in value_shapes], dtype=torch.float32, device=query.device)
            loc = rb[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
        else:
            loc = rb[:, :, None, :, None, :2] + off / P * rb[:, :, None, :, None, 2:] * 0.5
        shapes_t = torch.as_tensor([list(x) for x in value_shapes], dtype=torch.long, device=query.device)
        lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
        core = MSDeformAttnFunction.apply(v.contiguous(), shapes_t, lsi, loc.contiguous(), att.contiguous(),
                                          self.im2col_step)
        return F.linear(core, self.output_proj.weight, self.output_proj.bias).to(query.dtype)


class MOTRMSDeformAttn(MSDeformAttn):
    """transformer.py:290-391: MSDeformAttn plus the optional exp/(1+sum exp) normalisation."""


class _DecoderLayerBase(nn.Module):
    """Shared parameter layout of both decoder layer classes (SURVEY.md §8(b) checkpoint keys)."""

    def _build(self, d_model, n_heads, d_ffn, dropout, act, n_levels, n_points):
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.act = act
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)
        self._p_drop = dropout
        self.precision: Optional[str] = None

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def _check_mode(self, *inputs):
        if self.training and self._p_drop > 0:
            raise NotImplementedError("moyolo_b200 decoder layers are inference kernels: dropout>0 in "
                                      "training mode is not implemented")
        # The decoder-layer kernels build no autograd graph (packed, detached weights): refuse what looks like a
        # training step instead of silently returning outputs that give the parameters no gradient.
        if torch.is_grad_enabled() and (any(torch.is_tensor(t) and t.requires_grad for t in inputs) or
                                        (self.training and any(p.requires_grad for p in self.parameters()))):
            raise RuntimeError("moyolo_b200 decoder layers run inference kernels that build no autograd graph, but "
                               "gradients were requested (an input requires grad, or the layer is in training mode "
                               "with trainable parameters): call under torch.no_grad() / .eval(), or train through "
                               "MSDeformAttn(differentiable=True)")

    def _forward_impl(self, embed, refer_bbox, feats, shapes, padding_mask, attn_mask, query_pos):
        self._check_mode(embed, refer_bbox, feats, query_pos)
        dt = ex.lp_dtype(self.precision or ex.get_default_precision())
        pk = ex.cached_pack(self, "layer", ex.LayerPack, dt)
        bs, Q, C = embed.shape
        Lv = feats.shape[1]
        assert sum(s[0] * s[1] for s in shapes) == Lv
        ro, ro_host = ex.cached_dense_row_offsets(bs, Q, embed.device)
        x_f32 = embed.reshape(bs * Q, C).float().contiguous()
        pos = None if query_pos is None else query_pos.to(embed.dtype).reshape(bs * Q, C).float().contiguous()
        x_lp = x_f32 if dt == torch.float32 else ops.add_cast(x_f32, None, dt)
        xq_lp = x_lp if pos is None else ops.add_cast(x_f32, pos, dt)
        f2 = feats.reshape(bs * Lv, C)
        f_lp = f2.contiguous() if f2.dtype == dt else ops.add_cast(f2.float().contiguous(), None, dt)
        zero_rows = None if padding_mask is None else padding_mask.reshape(-1).to(torch.uint8).contiguous()
        val = ops.linear(f_lp, pk.msda.value.w, pk.msda.value.b, out_dtype=dt, zero_rows=zero_rows,
                         engine=ex._GEMM_ENGINE)
        refer = refer_bbox.reshape(bs * Q, 1, refer_bbox.shape[-1]).float().contiguous()
        mask = _prep_attn_mask(attn_mask, Q, embed.device)
        x_f32, _, _ = ex.run_layer(pk, x_f32, x_lp, xq_lp, refer, val.view(bs, Lv, C), shapes, bs, ro, ro_host, True,
                                   mask, pos, None, dt)
        return x_f32.view(bs, Q, C).to(embed.dtype)


def _prep_attn_mask(attn_mask, Q, device):
    if attn_mask is None:
        return None
    if attn_mask.dim() != 2 or attn_mask.shape != (Q, Q):
        raise NotImplementedError("only 2-D [Q, Q] attention masks are implemented")
    if attn_mask.dtype == torch.bool:  # True = masked out, nn.MultiheadAttention semantics
        m = torch.zeros(Q, Q, dtype=torch.float32, device=device)
        return m.masked_fill_(attn_mask.to(device), float("-inf")).contiguous()
    return attn_mask.to(device=device, dtype=torch.float32).contiguous()


class DeformableTransformerDecoderLayer(_DecoderLayerBase):
    """transformer.py:394-450."""

    def __init__(self, d_model=256, n_heads=8, d_ffn=1024, dropout=0., act=nn.ReLU(), n_levels=4, n_points=4):
        super().__init__()
        self._build(d_model, n_heads, d_ffn, dropout, act, n_levels, n_points)

    def forward(self, embed, refer_bbox, feats, shapes, padding_mask=None, attn_mask=None, query_pos=None):
        return self._forward_impl(embed, refer_bbox, feats, shapes, padding_mask, attn_mask, query_pos)


class MOTRDecoderLayer(_DecoderLayerBase):
    """transformer.py:515-652 (the live layer of MYDecoder, head.py:844)."""

    def __init__(self, d_model=256, n_heads=8, d_ffn=1024, dropout=0.1, act=nn.ReLU(), n_levels=4, n_points=4,
                 self_cross=True, my_softmax=False, local_self_attn=False, extra_track_attn=False):
        super().__init__()
        self.self_cross = self_cross
        self.local_self_attn = local_self_attn
        self.my_softmax = my_softmax
        self._build(d_model, n_heads, d_ffn, dropout, act, n_levels, n_points)
        self.extra_track_attn = extra_track_attn
        if extra_track_attn:  # parameters kept for checkpoint compatibility; unused by forward (:627-652)
            self.update_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
            self.dropout5 = nn.Dropout(dropout)
            self.norm4 = nn.LayerNorm(d_model)

    def forward(self, embed, refer_bbox, feats, shapes, padding_mask=None, attn_mask=None, track_query_pos=None):
        return self._forward_impl(embed, refer_bbox, feats, shapes, padding_mask, attn_mask, track_query_pos)


class _Dims:
    """d_model / n_heads / d_ffn / n_levels / n_points / nc of a decoder, as ClusterDecoder.supports reads them."""

    def __init__(self, d_model, n_heads, d_ffn, n_levels, n_points, nc):
        self.d_model, self.n_heads, self.d_ffn, self.n_levels, self.n_points, self.nc = d_model, n_heads, d_ffn, \
            n_levels, n_points, nc


class _DecoderBase(nn.Module):
    def __init__(self, hidden_dim, decoder_layer, num_layers, eval_idx=-1):
        super().__init__()
        self.layers = _clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.hidden_dim = hidden_dim
        self.eval_idx = eval_idx if eval_idx >= 0 else num_layers + eval_idx
        self.precision: Optional[str] = None

    def _run(self, embed, refer_bbox, feats, shapes, bbox_head, score_head, pos_mlp, attn_mask, padding_mask,
             fixed_pos):
        """Shared body of transformer.py:465-510 and :676-728.

        fixed_pos is None -> query_pos = pos_mlp(refer_bbox) recomputed per layer (:491);
        otherwise the same positional embedding is used by every layer (:705-707).
        """
        for l in self.layers:
            l._check_mode(embed, refer_bbox, feats, fixed_pos)
        prec = self.precision or ex.get_default_precision()
        dt = ex.lp_dtype(prec)
        bs, Q, C = embed.shape
        Lv = feats.shape[1]
        dev = embed.device
        assert sum(s[0] * s[1] for s in shapes) == Lv
        packs = [ex.cached_pack(l, "layer", ex.LayerPack, dt) for l in self.layers]
        bpacks = [ex.cached_pack(h, "mlp", ex.MlpPack, dt) for h in bbox_head]
        vp = ex.cached_pack(self.layers, "valueproj", ex.ValueProjPack, dt)
        ro, ro_host = ex.cached_dense_row_offsets(bs, Q, dev)
        R = bs * Q

        x_f32 = embed.reshape(R, C).float().contiguous()
        refer = ops.sigmoid(refer_bbox.reshape(R, refer_bbox.shape[-1]).float())  # :482 / :690
        f2 = feats.reshape(bs * Lv, C)
        f_lp = f2.contiguous() if f2.dtype == dt else ops.add_cast(f2.float().contiguous(), None, dt)
        zero_rows = None if padding_mask is None else padding_mask.reshape(-1).to(torch.uint8).contiguous()
        n_l = len(self.layers)
        values = ops.linear(f_lp, vp.w, vp.b, out_dtype=dt, zero_rows=zero_rows, engine=ex._GEMM_ENGINE)
        values = values.view(bs, Lv, n_l * C)
        mask = _prep_attn_mask(attn_mask, Q, dev)

        if fixed_pos is not None:
            pos = fixed_pos.to(embed.dtype).reshape(R, C).float().contiguous()
            pos_pack = None
        else:
            pos_pack = ex.cached_pack(pos_mlp, "mlp", ex.MlpPack, dt)
            pos = ex.pos_mlp_forward(pos_pack, refer, dt)
        x_lp = x_f32 if dt == torch.float32 else ops.add_cast(x_f32, None, dt)
        xq_lp = ops.add_cast(x_f32, pos, dt)

        n_out = n_l if self.training else 1
        dec_bboxes = torch.empty(n_out, bs, Q, 4, dtype=torch.float32, device=dev)
        dec_cls = []
        # Eval, one fixed positional embedding (the MOTR decoder), no masks: every layer up to eval_idx, the box
        # refinements and the score head run as ONE cluster kernel (csrc/decoder_cluster.cu) when the frame fits.
        if (ex.CLUSTER_DECODER and not self.training and fixed_pos is not None and mask is None and zero_rows is None and
                refer.shape[-1] == 4 and
                ex.ClusterDecoder.supports(dt, _Dims(C, packs[0].n_heads, packs[0].ffn1.w.shape[0], packs[0].msda.n_levels,
                                                     packs[0].msda.n_points, score_head[self.eval_idx].weight.shape[0]))):
            n_run = self.eval_idx + 1
            head = score_head[self.eval_idx]
            sw, sb = head.weight.detach().float().contiguous(), head.bias.detach().float().contiguous()
            cd = ex.ClusterDecoder(packs[:n_run], bpacks[:n_run], shapes, sw, sb)
            m_rows = cd.tile_rows(R, bs, Q)
            if m_rows:
                nc = sw.shape[0]
                x_out = torch.empty(R, C, dtype=torch.float32, device=dev)
                logits = torch.empty(R, nc, dtype=torch.float32, device=dev)
                kv = torch.empty(2, R, 2 * C, dtype=dt, device=dev)
                bar = torch.zeros(1, dtype=torch.int32, device=dev)
                status = torch.zeros(1, dtype=torch.int32, device=dev)
                refer_out = [None] * (n_run - 1) + [dec_bboxes[0].view(R, 4)]
                cd.run(x_f32, pos, refer.view(R, 4).contiguous(), values, ro, bs, R, m_rows, x_out, kv, bar, refer_out,
                       logits=logits, status=status)
                out_dt = embed.dtype
                return dec_bboxes.to(out_dt), logits.view(1, bs, Q, nc).to(out_dt), x_out.view(bs, Q, C).to(out_dt)
        for i, pk in enumerate(packs):
            pos_next = pos if (fixed_pos is not None and i + 1 < n_l) else None
            x_f32, x_lp, xq_next = ex.run_layer(pk, x_f32, x_lp, xq_lp, refer.view(R, 1, -1),
                                                values[:, :, i * C:(i + 1) * C], shapes, bs, ro, ro_host, True, mask,
                                                pos, pos_next, dt)
            if self.training:
                # :712-716 — dec_bboxes[i] re-evaluates the same expression on the un-detached previous
                # box, which is numerically the refined box itself
                refined = ex.bbox_head(bpacks[i], x_lp, refer, out=dec_bboxes[i]).view(R, 4)
                dec_cls.append(self._scores(score_head[i], x_lp, bs, Q))
            elif i == self.eval_idx:
                ex.bbox_head(bpacks[i], x_lp, refer, out=dec_bboxes[0])
                dec_cls.append(self._scores(score_head[i], x_lp, bs, Q))
                break
            else:
                refined = ex.bbox_head(bpacks[i], x_lp, refer)
            refer = refined
            if fixed_pos is None and i + 1 < n_l:
                pos = ex.pos_mlp_forward(pos_pack, refer, dt)
                xq_lp = ops.add_cast(x_f32, pos, dt)
            else:
                xq_lp = xq_next
        out_dt = embed.dtype
        return dec_bboxes.to(out_dt), torch.stack(dec_cls).to(out_dt), x_f32.view(bs, Q, C).to(out_dt)

    @staticmethod
    def _scores(head, x_lp, bs, Q):
        w = head.weight.detach().float().contiguous()
        b = head.bias.detach().float().contiguous()
        logits, _, _ = ops.score_head(x_lp, w, b, want_scores=False)
        return logits.view(bs, Q, -1)


class DeformableTransformerDecoder(_DecoderBase):
    """transformer.py:453-510."""

    def forward(self, embed, refer_bbox, feats, shapes, bbox_head, score_head, pos_mlp, attn_mask=None,
                padding_mask=None):
        boxes, cls, _ = self._run(embed, refer_bbox, feats, shapes, bbox_head, score_head, pos_mlp, attn_mask,
                                  padding_mask, None)
        return boxes, cls


class MOTRTransformerDecoder(_DecoderBase):
    """transformer.py:663-728 (the live decoder: one fixed track_query_embed for all layers)."""

    def __init__(self, hidden_dim, decoder_layer, num_layers, eval_idx=-1):
        super().__init__(hidden_dim, decoder_layer, num_layers, eval_idx)
        self.bbox_embed = None

    def forward(self, embed, refer_bbox, feats, shapes, bbox_head, score_head, pos_mlp, attn_mask=None,
                padding_mask=None, track_query_embed=None):
        if track_query_embed is None:
            raise AttributeError("'NoneType' object has no attribute 'dtype'")  # transformer.py:635
        return self._run(embed, refer_bbox, feats, shapes, bbox_head, score_head, pos_mlp, attn_mask, padding_mask,
                         track_query_embed)
