"""Encoder-side query selection of MYDecoder on device.

Mirrors `MYDecoder._get_encoder_input`, `_generate_anchors` and the detect-query part of `_get_decoder_input`
(ultralytics/nn/modules/head.py:1012-1029, 993-1010, 1031-1113), eval mode, no denoising, `learnt_init_query=False`:

    neck maps [S, C_l, H_l, W_l]  --1x1 conv + BatchNorm (folded) per level-->  feats [S, Lv, 256]
    features = LayerNorm(Linear(valid_mask * feats));  logits = score_head(features)
    top-k(num_queries) by max-class logit  ->  det_embed = features[topk], enc_scores = logits[topk],
    det_refer = enc_bbox_head(features[topk]) + anchors[topk]          (logit-space boxes)

Differences in schedule, not in values: BatchNorm is folded into the conv weights; the class-score head runs in the
epilogue of the enc_output GEMM; the 3-layer box MLP is evaluated on the selected rows only (it is row-wise, the
reference computes it for all Lv rows and keeps num_queries of them). Maps are taken channels-last
([S, H_l, W_l, C_l], i.e. `x.permute(0, 2, 3, 1)` of the reference's NCHW tensors) so that positions are GEMM rows.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from . import _lib
from . import executor as ex
from . import ops
from .synthetic import DecoderSpec, level_sizes


class QuerySelector:
    def __init__(self, sd: Dict[str, torch.Tensor], spec: DecoderSpec, shapes, ch: Sequence[int], device,
                 precision: str = "bf16", n_queries: int = 300, n_seq: int = 1):
        self.spec, self.shapes, self.ch = spec, [list(s) for s in shapes], list(ch)
        self.dev, self.nq, self.S = torch.device(device), n_queries, n_seq
        self.dt = ex.lp_dtype(precision)
        self.Lv = level_sizes(shapes)
        self.starts = [0]
        for h, w in self.shapes:
            self.starts.append(self.starts[-1] + h * w)
        dev, dt, C = self.dev, self.dt, spec.d_model
        f = lambda t: t.to(dev).float().contiguous()  # noqa: E731
        # 1x1 conv (no bias) + BatchNorm2d in eval mode == Linear with folded scale / shift (head.py:839-840)
        self.proj = []
        for l, c in enumerate(self.ch):
            p = f"input_proj.{l}."
            scale = sd[p + "1.weight"].double() / torch.sqrt(sd[p + "1.running_var"].double() + 1e-5)
            w = sd[p + "0.weight"].double().view(C, c) * scale[:, None]
            b = sd[p + "1.bias"].double() - sd[p + "1.running_mean"].double() * scale
            self.proj.append((w.float().to(dev).to(dt).contiguous(), f(b)))
        self.enc_w = sd["enc_output.0.weight"].to(dev).to(dt).contiguous()
        self.enc_b, self.enc_g, self.enc_beta = f(sd["enc_output.0.bias"]), f(sd["enc_output.1.weight"]), f(sd["enc_output.1.bias"])
        self.score_w, self.score_b = f(sd["enc_score_head.weight"]), f(sd["enc_score_head.bias"])
        self.box_hidden = [(sd[f"enc_bbox_head.layers.{j}.weight"].to(dev).to(dt).contiguous(),
                            f(sd[f"enc_bbox_head.layers.{j}.bias"])) for j in (0, 1)]
        self.box_w3, self.box_b3 = f(sd["enc_bbox_head.layers.2.weight"]), f(sd["enc_bbox_head.layers.2.bias"])
        # static per-shape data and the workspace (no allocation per frame: graph-capturable)
        S, Lv, nq, nc = n_seq, self.Lv, n_queries, spec.nc
        self.invalid = ops.anchor_invalid(self.shapes, Lv, dev).repeat(S).contiguous()       # [S*Lv] uint8
        z = torch.zeros
        self.features = z(S, Lv, C, device=dev)
        self.logits = z(S, Lv, nc, device=dev)
        self.max_logit = z(S, Lv, device=dev)
        self.idx = z(S, nq, dtype=torch.int32, device=dev)
        self.sel_lp = z(S * nq, C, dtype=dt, device=dev)
        self.h1 = z(S * nq, C, dtype=dt, device=dev)
        self.h2 = z(S * nq, C, dtype=dt, device=dev)
        self.enc_scores = z(S, nq, nc, device=dev)
        if dt == torch.float32:
            self.masked = z(S * Lv, C, device=dev)
            self.t = z(S * Lv, C, device=dev)

    def map_shapes(self) -> List[tuple]:
        """Shapes of the channels-last input maps, one per level."""
        return [(self.S, h, w, c) for (h, w), c in zip(self.shapes, self.ch)]

    def run(self, maps: Sequence[torch.Tensor], feats: torch.Tensor, det_embed: torch.Tensor,
            det_refer: torch.Tensor) -> None:
        """maps[l]: [S, H_l, W_l, C_l] of the GEMM dtype, contiguous. Writes feats [S, Lv, 256] (GEMM dtype),
        det_embed [S, nq, 256] fp32 and det_refer [S, nq, 4] fp32 (logit space); `self.enc_scores` / `self.idx`
        hold the selected class logits and pyramid positions."""
        self.project(maps, feats)
        self.select(feats, det_embed, det_refer)

    def project(self, maps: Sequence[torch.Tensor], feats: torch.Tensor) -> None:
        """input_proj (head.py:1012-1029): per level and sequence one GEMM with the folded conv + BatchNorm."""
        eng = ex._GEMM_ENGINE
        for l, (h, w) in enumerate(self.shapes):
            pw, pb = self.proj[l]
            for s in range(self.S):
                ops.linear(maps[l][s].view(h * w, self.ch[l]), pw, pb, out=feats[s, self.starts[l]:self.starts[l + 1]],
                           engine=eng)

    def select(self, feats: torch.Tensor, det_embed: torch.Tensor, det_refer: torch.Tensor) -> None:
        """enc_output + scores over all positions, top-k, gathers, box head + anchors (head.py:1031-1113)."""
        S, Lv, C, nq = self.S, self.Lv, self.spec.d_model, self.nq
        eng = ex._GEMM_ENGINE
        rows = feats.view(S * Lv, C)
        if self.dt == torch.bfloat16 and eng != _lib.GEMM_SIMT:
            ops.enc_output_scores(rows, self.enc_w, self.enc_b, self.enc_g, self.enc_beta, 1e-5, self.invalid,
                                  self.score_w, self.score_b, out_f32=self.features.view(S * Lv, C),
                                  logits=self.logits.view(S * Lv, -1), max_logit=self.max_logit.view(-1))
        else:
            x = rows if rows.dtype == torch.float32 else rows.float()
            ops.mask_rows(x, self.invalid, out=self.masked)
            m = self.masked if self.dt == torch.float32 else self.masked.to(self.dt)
            ops.linear(m, self.enc_w, self.enc_b, out=self.t, engine=eng)
            ops.add_layernorm(self.t, None, self.enc_g, self.enc_beta, 1e-5, out_f32=self.features.view(S * Lv, C))
            ops.score_head(self.features.view(S * Lv, C), self.score_w, self.score_b, want_scores=False,
                           out=(self.logits.view(S * Lv, -1), None, None), max_logit=self.max_logit.view(-1))
        ops.topk(self.max_logit, nq, out=self.idx)                                           # head.py:1048
        ops.select_gather(self.features, self.logits, self.idx, det_embed, embed_lp=self.sel_lp,
                          enc_scores=self.enc_scores)                                        # :1092, :1104
        (w0, b0), (w1, b1) = self.box_hidden
        ops.linear(self.sel_lp, w0, b0, relu=True, out=self.h1, engine=eng)
        ops.linear(self.h1, w1, b1, relu=True, out=self.h2, engine=eng)
        ops.anchor_box(self.h2, self.box_w3, self.box_b3, self.idx.view(-1), self.shapes, Lv, out=det_refer.view(-1, 4))
