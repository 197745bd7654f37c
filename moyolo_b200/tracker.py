"""Per-frame tracking engine: query assembly -> 6-layer decoder -> heads -> track-query update.

Mirrors the frame semantics of MOTRTrack.forward / _post_process_single_image
(ultralytics/nn/modules/head.py:191-239, 300-321, 492-497) with carried tracks as specified by
"O3" in SURVEY.md §8(c) / oracle/tracker_port.py:
  * queries = [carried tracks ; detect queries] per sequence (head.py:1056-1064, 1108-1109);
    track content embedding = denoising_class_embed[argmax previous logits] (head.py:888-900);
  * after the decode every field has N = T + n_detect rows, ids = cat(prev ids, -1), disappear =
    cat(prev, 0) (repair R2, upstream MOTR/models/motr.py:569-574);
  * RuntimeTrackerBase.update on device (moyolo_track_assign_batched), active selection (ids >= 0,
    MOTR/models/qim.py:184-187) and QIM._update_track_embedding (qim.py:251-301) produce the next
    frame's track queries.

Several independent sequences run in lock-step as one ragged batch; frames of one
sequence are strictly ordered. ALL state lives on the device in fixed-capacity arrays with the
per-sequence track counts in device memory, so the launch shapes of a frame depend only on the
padded row count: each padded size is captured once as a CUDA graph and replayed. The host reads
back one int32 per sequence and frame (the active-track count that selects the next graph).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import executor as ex
from . import ops
from .synthetic import DecoderSpec, level_sizes


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

    x_f32 [R, C], refer_logit [R, 4], pos [R, C] fp32; feats_lp [n_seq*Lv, C] in the GEMM dtype; ro
    device row offsets [n_seq+1]; ro_host a host list that bounds them (grid sizing only).
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


def qim_update(W: DecoderWeights, ref_pts, query_pos, out_embed, ro, ro_host, seg_len=None):
    """QueryInteractionModule._update_track_embedding (MOTR/models/qim.py:251-301) for the active tracks
    of every sequence at once: rows of sequence s are [ro[s], ro[s] + seg_len[s]) (the rest of the
    [R, .] buffers is padding). ref_pts [R,4] logits, query_pos / out_embed [R,C] fp32.
    Returns new_query_pos [R,C] fp32 (ref_pts <- inverse_sigmoid(pred_boxes) is fused in the write-back)."""
    R, C = out_embed.shape
    dt, q, H = W.dt, W.qim, 8  # nn.MultiheadAttention(dim_in, 8, ...) qim.py:88
    eng = ex._GEMM_ENGINE
    dev = out_embed.device
    qpos = ops.pos2posemb(ref_pts)                                   # qim.py:255
    qk_lp = ops.add_cast(qpos, out_embed, dt)                        # :271
    tgt_lp = out_embed if dt == torch.float32 else ops.add_cast(out_embed, None, dt)
    qkv = torch.empty(R, 3 * C, dtype=dt, device=dev)
    ops.linear(qk_lp, q["qk_w"], q["qk_b"], out=qkv[:, :2 * C], engine=eng)
    ops.linear(tgt_lp, q["v_w"], q["v_b"], out=qkv[:, 2 * C:], engine=eng)
    att = ops.self_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], ro, ro_host, H, seg_len=seg_len)
    t2 = ops.linear(att, q["o_w"], q["o_b"], out_dtype=torch.float32, engine=eng)
    tgt_f32, tgt_lp, _ = ops.add_layernorm(t2, out_embed, *q["norm1"], 1e-5, True, True, dt)     # :277-278
    h = ops.linear(tgt_lp, q["l1_w"], q["l1_b"], relu=True, engine=eng)
    t3 = ops.linear(h, q["l2_w"], q["l2_b"], out_dtype=torch.float32, engine=eng)                 # :280
    _, tgt_lp, _ = ops.add_layernorm(t3, tgt_f32, *q["norm2"], 1e-5, False, True, dt)             # :281-282
    h = ops.linear(tgt_lp, q["f1_w"], q["f1_b"], relu=True, engine=eng)
    f2 = ops.linear(h, q["f2_w"], q["f2_b"], out_dtype=torch.float32, engine=eng)                 # :290
    new_qpos, _, _ = ops.add_layernorm(f2, query_pos, *q["norm_feat"], 1e-5, True, False, dt)     # :294-298
    return new_qpos


class _FramePlan:
    """Static buffers (and optionally the captured CUDA graph) of one padded frame size."""
    __slots__ = ("rows_pad", "graph", "ids", "boxes", "scores", "labels", "logits", "n_active", "ro", "n_launch")


class TrackEngine:
    """Lock-step tracker for `n_seq` independent sequences on one GPU."""

    def __init__(self, sd, spec: DecoderSpec, shapes, device, precision: str = "bf16", n_detect: int = 300,
                 n_seq: int = 1, score_thresh=0.4, filter_thresh=0.5, miss_tolerance=5, iou_thresh=0.8,
                 weights: Optional[DecoderWeights] = None, cap: int = 512, bucket: int = 32,
                 use_graphs: bool = True):
        self.dev = torch.device(device)
        self.spec, self.shapes, self.n_detect, self.n_seq = spec, [list(s) for s in shapes], n_detect, n_seq
        self.Lv = level_sizes(shapes)
        self.W = weights or DecoderWeights(sd, spec, self.dev, precision)
        self.thr = (score_thresh, filter_thresh, miss_tolerance, iou_thresh)
        self.cap, self.bucket, self.use_graphs = cap, bucket, use_graphs
        S, C, dev = n_seq, spec.d_model, self.dev
        # device-resident track state (fixed capacity)
        self.n_tracks = torch.zeros(S, dtype=torch.int32, device=dev)
        self.t_ref = torch.zeros(S, cap, 4, device=dev)
        self.t_qpos = torch.zeros(S, cap, C, device=dev)
        self.t_label = torch.zeros(S, cap, dtype=torch.int32, device=dev)
        self.t_ids = torch.zeros(S, cap, dtype=torch.int64, device=dev)
        self.t_dis = torch.zeros(S, cap, dtype=torch.int64, device=dev)
        self.counters = torch.zeros(S, 2, dtype=torch.int64, device=dev)
        # static frame inputs (graphs read these)
        self.feats_in = torch.zeros(S, self.Lv, C, dtype=self.W.dt, device=dev)
        self.det_embed_in = torch.zeros(S, n_detect, C, device=dev)
        self.det_refer_in = torch.zeros(S, n_detect, 4, device=dev)
        self._count_host = torch.zeros(S, dtype=torch.int32).pin_memory()
        self._T = [0] * S
        self._plans: Dict[int, _FramePlan] = {}
        self.frame_idx = 0

    # ---- state -------------------------------------------------------------------------------
    def reset(self, seq: Optional[int] = None):
        """is_first semantics (head.py:199-205): drop all tracks, fresh ID counters."""
        if seq is None:
            self.n_tracks.zero_()
            self.counters.zero_()
            self._T = [0] * self.n_seq
            self.frame_idx = 0
        else:
            self.n_tracks[seq] = 0
            self.counters[seq].zero_()
            self._T[seq] = 0

    def n_tracks_host(self) -> List[int]:
        return list(self._T)

    n_tracks_list = n_tracks_host

    def track_ids(self, s: int) -> torch.Tensor:
        return self.t_ids[s, :self._T[s]]

    def track_disappear(self, s: int) -> torch.Tensor:
        return self.t_dis[s, :self._T[s]]

    # ---- one frame ----------------------------------------------------------------------------
    def _body(self, rows_pad: int) -> _FramePlan:
        """Every launch of one frame for a padded row count (graph-capturable: no host sync, shapes
        depend on rows_pad only, counts are read from device memory by the kernels)."""
        W, dev, S, nd, C = self.W, self.dev, self.n_seq, self.n_detect, self.spec.d_model
        R = rows_pad
        x = torch.empty(R, C, device=dev)
        refer_logit = torch.empty(R, 4, device=dev)
        pos = torch.empty(R, C, device=dev)
        ids = torch.empty(R, dtype=torch.int64, device=dev)
        dis = torch.empty(R, dtype=torch.int64, device=dev)
        ro = torch.empty(S + 1, dtype=torch.int32, device=dev)
        ops.frame_assemble(S, nd, C, self.cap, self.n_tracks, self.t_ref, self.t_qpos, self.t_label, self.t_ids,
                           self.t_dis, W.class_embed, self.det_embed_in, self.det_refer_in, x, refer_logit, pos, ids,
                           dis, ro, R)
        # host-side bound of the device offsets, used for grid sizing only: sum_s ceil(N_s/16) <= R/16 + S
        ro_host = [16 * i for i in range(S)] + [16 * S + R]
        feats_lp = self.feats_in.view(S * self.Lv, C)
        boxes, logits, scores, labels, hs = decode_frame(W, x, refer_logit, pos, feats_lp, self.shapes, S, ro, ro_host)
        st, ft, mt, it = self.thr
        ws = torch.empty(S * ops.track_workspace_bytes(R), dtype=torch.uint8, device=dev)
        ops.track_assign_batched(scores, boxes, ids, dis, self.counters, ro, S, R, ws, st, ft, mt, it)
        n_active = torch.empty(S, dtype=torch.int32, device=dev)
        active_index = torch.empty(R, dtype=torch.int32, device=dev)
        c_ref = torch.zeros(R, 4, device=dev)
        c_pos = torch.zeros(R, C, device=dev)
        c_hs = torch.zeros(R, C, device=dev)
        c_box = torch.zeros(R, 4, device=dev)
        ops.frame_compact(S, C, self.cap, ro, ids, dis, labels, refer_logit, pos, hs, boxes, n_active, active_index,
                          c_ref, c_pos, c_hs, c_box, self.t_label, self.t_ids, self.t_dis)
        new_qpos = qim_update(W, c_ref, c_pos, c_hs, ro, ro_host, seg_len=n_active)
        ops.frame_writeback(S, C, self.cap, ro, n_active, new_qpos, c_box, self.t_qpos, self.t_ref, self.n_tracks)
        p = _FramePlan()
        p.rows_pad, p.graph = R, None
        p.ids, p.boxes, p.scores, p.labels, p.logits, p.n_active, p.ro = ids, boxes, scores, labels, logits, n_active, ro
        return p

    def _state_snapshot(self):
        return [t.clone() for t in (self.n_tracks, self.t_ref, self.t_qpos, self.t_label, self.t_ids, self.t_dis,
                                    self.counters)]

    def _state_restore(self, snap):
        for dst, src in zip((self.n_tracks, self.t_ref, self.t_qpos, self.t_label, self.t_ids, self.t_dis,
                             self.counters), snap):
            dst.copy_(src)

    def _plan(self, rows_pad: int) -> _FramePlan:
        p = self._plans.get(rows_pad)
        if p is None:
            snap = self._state_snapshot()
            torch.cuda.synchronize(self.dev)
            side = torch.cuda.Stream(self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):  # one eager pass: loads modules, sets kernel attributes
                self._body(rows_pad)
            torch.cuda.current_stream(self.dev).wait_stream(side)
            torch.cuda.synchronize(self.dev)
            self._state_restore(snap)
            g = torch.cuda.CUDAGraph()
            before = ops.LAUNCHES
            with torch.cuda.graph(g):
                p = self._body(rows_pad)
            p.graph = g
            p.n_launch = ops.LAUNCHES - before  # kernels replayed by every graph launch
            ops.LAUNCHES = before
            self._state_restore(snap)  # capture does not execute, but keep the invariant explicit
            self._plans[rows_pad] = p
        return p

    def prepare(self, max_tracks_per_seq: int) -> int:
        """Pre-capture the frame graphs for every padded size up to `max_tracks_per_seq` tracks per
        sequence (done at start-up so no capture happens in the frame loop). Returns #graphs."""
        if not self.use_graphs:
            return 0
        lo = self.n_seq * self.n_detect
        hi = self.n_seq * (self.n_detect + max_tracks_per_seq)
        r = self._round(lo)
        while r <= self._round(hi):
            self._plan(r)
            r += self.bucket
        return len(self._plans)

    def _round(self, rows: int) -> int:
        b = self.bucket
        return (rows + b - 1) // b * b

    def load_inputs(self, feats: torch.Tensor, det_embed: torch.Tensor, det_refer: torch.Tensor) -> None:
        """Copy (device or pinned-host) frame inputs into the static input buffers."""
        if feats.data_ptr() != self.feats_in.data_ptr():
            if feats.dtype == self.feats_in.dtype or not feats.is_cuda:
                self.feats_in.copy_(feats.reshape(self.feats_in.shape), non_blocking=True)
            else:  # dtype conversion on device through the library's own cast kernel
                src = feats.reshape(-1).float().contiguous()
                self.feats_in.view(-1).copy_(ops.add_cast(src, None, self.feats_in.dtype))
        if det_embed.data_ptr() != self.det_embed_in.data_ptr():
            self.det_embed_in.copy_(det_embed.reshape(self.det_embed_in.shape), non_blocking=True)
        if det_refer.data_ptr() != self.det_refer_in.data_ptr():
            self.det_refer_in.copy_(det_refer.reshape(self.det_refer_in.shape), non_blocking=True)

    def step(self, feats: torch.Tensor, det_embed: torch.Tensor, det_refer: torch.Tensor) -> List[Dict[str, torch.Tensor]]:
        """feats [n_seq, Lv, C] (GEMM dtype or fp32), det_embed [n_seq, nd, C] fp32, det_refer
        [n_seq, nd, 4] fp32 logit-space boxes. Returns one dict per sequence with the N = T + nd rows
        of this frame: ids (int64, -1 = no object), boxes (cx,cy,w,h normalised), scores, labels, logits.
        The returned tensors are views of per-size static buffers: consume or copy them before the next
        frame of the same padded size."""
        S, nd = self.n_seq, self.n_detect
        self.load_inputs(feats, det_embed, det_refer)
        rows = sum(self._T) + S * nd
        rows_pad = self._round(rows)
        if self.use_graphs:
            p = self._plan(rows_pad)
            p.graph.replay()
            ops.LAUNCHES += p.n_launch
        else:
            p = self._body(rows_pad)
        # the one host read-back of the frame: active-track counts (they size the next frame)
        self._count_host.copy_(p.n_active, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        outs, off = [], 0
        for s in range(S):
            n = self._T[s] + nd
            outs.append({"ids": p.ids[off:off + n], "boxes": p.boxes[off:off + n], "scores": p.scores[off:off + n],
                         "labels": p.labels[off:off + n], "logits": p.logits[off:off + n]})
            off += n
        self._T = [int(v) for v in self._count_host.tolist()]
        self.frame_idx += 1
        return outs


class SequenceTracker(TrackEngine):
    """Single-sequence convenience wrapper: 2-D inputs, one dict out."""

    def __init__(self, sd, spec, shapes, device, precision="bf16", n_detect=300, **kw):
        super().__init__(sd, spec, shapes, device, precision, n_detect, 1, **kw)

    def step(self, feats, det_embed, det_refer):  # type: ignore[override]
        return super().step(feats[None], det_embed[None], det_refer[None])[0]
