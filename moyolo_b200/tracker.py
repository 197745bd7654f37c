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

Execution model (DESIGN.md "Frame pipeline"):
  * Several independent sequences run in lock-step as one ragged batch; frames of
    one sequence are strictly ordered.
  * ALL state lives on the device in fixed-capacity arrays with the per-sequence track counts in
    device memory, and every intermediate of a frame lives in a pre-allocated per-plan workspace, so
    the launches of a frame depend only on the padded row count: each (padded size, input slot) is
    captured once as a CUDA graph and replayed. Inside the graph the all-layers value projection and
    every layer's box head run on side branches next to the main chain.
  * The host runs AHEAD of the device: `submit()` enqueues frame t while frame t-1 is still running,
    picking the padded size from the newest track counts it already knows plus a margin. If the real
    row count does not fit, the device aborts that frame (and any later one) without touching the
    track state and the host re-launches it with the exact size (moyolo_frame_assemble, `ctrl`).
    Inputs go through a 2-deep device ring filled on a copy stream, so the host->device copy of frame
    t+1 overlaps the compute of frame t; results leave as ONE packed [rows, 8] copy per frame and as
    rows appended on the device to a resident track table.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional

import torch

from . import _lib
from . import executor as ex
from . import ops
from .synthetic import DecoderSpec, level_sizes

CTRL_ABORT, CTRL_FRAME, CTRL_CURSOR, CTRL_TABLE_OVERFLOW, CTRL_ABORT_ROWS, CTRL_TRACK_OVERFLOW = 0, 1, 2, 3, 4, 5


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
            "qkv_w": w(sd[q + "self_attn.in_proj_weight"]), "qkv_b": f(sd[q + "self_attn.in_proj_bias"]),
            "o_w": w(sd[q + "self_attn.out_proj.weight"]), "o_b": f(sd[q + "self_attn.out_proj.bias"]),
            "l1_w": w(sd[q + "linear1.weight"]), "l1_b": f(sd[q + "linear1.bias"]),
            "l2_w": w(sd[q + "linear2.weight"]), "l2_b": f(sd[q + "linear2.bias"]),
            "f1_w": w(sd[q + "linear_feat1.weight"]), "f1_b": f(sd[q + "linear_feat1.bias"]),
            "f2_w": w(sd[q + "linear_feat2.weight"]), "f2_b": f(sd[q + "linear_feat2.bias"]),
        }
        for n in ("norm1", "norm2", "norm_feat"):
            self.qim[n] = (f(sd[q + n + ".weight"]), f(sd[q + n + ".bias"]))


class FrameWorkspace:
    """Every intermediate tensor of one frame for a padded row count R (allocated once, outside any
    graph capture; kernels write through `out=` so a frame performs no allocation)."""

    def __init__(self, R: int, S: int, spec: DecoderSpec, dt: torch.dtype, dev, d_qim: int, rows_per_seq: int = 0):
        C, F, nl = spec.d_model, spec.d_ffn, spec.n_layers
        f32 = dict(dtype=torch.float32, device=dev)
        lp = dict(dtype=dt, device=dev)
        z = torch.zeros
        self.R = R
        # frame inputs built by frame_assemble
        self.x = z(R, C, **f32)              # residual stream
        self.refer_logit = z(R, 4, **f32)
        self.pos = z(R, C, **f32)
        self.ids0 = z(R, dtype=torch.int64, device=dev)    # as assembled (previous frame's ids | -1)
        self.dis0 = z(R, dtype=torch.int64, device=dev)
        self.ids = z(R, dtype=torch.int64, device=dev)     # after this frame's ID assignment
        self.dis = z(R, dtype=torch.int64, device=dev)
        self.ro = z(S + 1, dtype=torch.int32, device=dev)
        self.refer = [z(R, 4, **f32) for _ in range(nl + 1)]   # sigmoid(refer) and the refined boxes per layer
        # decoder layer
        self.x_lp = self.x if dt == torch.float32 else z(R, C, **lp)
        self.xq_lp = z(R, C, **lp)
        self.qkv = z(R, 3 * C, **lp)
        self.att = z(R, C, **lp)
        self.t = z(R, C, **f32)              # GEMM output feeding an add+LayerNorm
        self.x1 = z(R, C, **f32)
        self.x1q_lp = z(R, C, **lp)
        self.ol = z(R, ex.offlog_width(spec), **f32)
        self.g = z(R, C, **lp)
        self.x2 = z(R, C, **f32)
        self.x2_lp = self.x2 if dt == torch.float32 else z(R, C, **lp)
        self.h = z(R, F, **lp)
        # box head (side branch)
        self.bh1 = z(R, C, **lp)
        self.bh2 = z(R, C, **lp)
        # heads + track update
        self.logits = z(R, spec.nc, **f32)
        self.scores = z(R, **f32)
        self.labels = z(R, dtype=torch.int32, device=dev)
        self.rows_per_seq = min(R, rows_per_seq) if rows_per_seq > 0 else R   # bound of one sequence's rows
        self.assign_ws = z(S * ops.track_workspace_bytes(self.rows_per_seq), dtype=torch.uint8, device=dev)
        self.n_active = z(S, dtype=torch.int32, device=dev)
        self.active_index = z(R, dtype=torch.int32, device=dev)
        self.c_ref = z(R, 4, **f32)
        self.c_pos = z(R, C, **f32)
        self.c_hs = z(R, C, **f32)
        self.c_box = z(R, 4, **f32)
        # QIM
        self.q_pos = z(R, C, **f32)
        self.q_qk_lp = z(R, C, **lp)
        self.q_tgt_lp = self.c_hs if dt == torch.float32 else z(R, C, **lp)
        self.q_tgt = z(R, C, **f32)
        self.q_tgt_lp2 = self.q_tgt if dt == torch.float32 else z(R, C, **lp)
        self.q_tgt_lp3 = z(R, C, **lp)
        self.q_h = z(R, d_qim, **lp)
        self.q_new = z(R, C, **f32)
        # cluster decoder (csrc/decoder_cluster.cu): self-attention K|V scratch (double-buffered by layer parity),
        # grid barrier counter, status flag
        self.kv = z(2, R, 2 * C, **lp) if dt == torch.bfloat16 else None
        self.grid_bar = z(1, dtype=torch.int32, device=dev)
        self.dc_status = z(1, dtype=torch.int32, device=dev)


def qim_update_ws(W: DecoderWeights, spec: DecoderSpec, ws: "FrameWorkspace", ro_host) -> None:
    """QueryInteractionModule._update_track_embedding (MOTR/models/qim.py:251-301) for the active
    tracks of every sequence at once: rows of sequence s are [ro[s], ro[s] + n_active[s]) of the
    compact buffers (the rest is padding). ref_pts <- inverse_sigmoid(pred_boxes) (qim.py:299) is
    fused into frame_writeback."""
    C = spec.d_model
    dt, q, H = W.dt, W.qim, 8  # nn.MultiheadAttention(dim_in, 8, ...) qim.py:88
    eng = ex._GEMM_ENGINE
    # q = k = tgt + pos2posemb(ref_pts), v = tgt (qim.py:255, 271): operands written by frame_compact
    ex.qkv_proj(ws.q_qk_lp, ws.q_tgt_lp, q["qkv_w"], q["qkv_b"], ws.qkv, C, eng)
    ops.self_attention(ws.qkv[:, :C], ws.qkv[:, C:2 * C], ws.qkv[:, 2 * C:], ws.ro, ro_host, H, out=ws.att,
                       seg_len=ws.n_active)
    if ex.fused_epilogues(dt, C) and q["l1_w"].shape[0] % 64 == 0:
        ops.linear_add_layernorm(ws.att, q["o_w"], q["o_b"], ws.c_hs, *q["norm1"], 1e-5, out_f32=ws.q_tgt,   # :277-278
                                 out_lp=ws.q_tgt_lp2)
        if ops.ffn_fused_supported(dt, C, q["l1_w"].shape[0]):   # each FFN block of the QIM in one launch
            ops.ffn_add_layernorm(ws.q_tgt_lp2, q["l1_w"], q["l1_b"], q["l2_w"], q["l2_b"], ws.q_h, ws.q_tgt,
                                  *q["norm2"], 1e-5, out_lp=ws.q_tgt_lp3)                                    # :280-282
            ops.ffn_add_layernorm(ws.q_tgt_lp3, q["f1_w"], q["f1_b"], q["f2_w"], q["f2_b"], ws.q_h, ws.c_pos,
                                  *q["norm_feat"], 1e-5, out_f32=ws.q_new)                                   # :290-298
            return
        ops.linear(ws.q_tgt_lp2, q["l1_w"], q["l1_b"], relu=True, out=ws.q_h, engine=eng)
        ops.linear_add_layernorm(ws.q_h, q["l2_w"], q["l2_b"], ws.q_tgt, *q["norm2"], 1e-5,                  # :280-282
                                 out_lp=ws.q_tgt_lp3)
        ops.linear(ws.q_tgt_lp3, q["f1_w"], q["f1_b"], relu=True, out=ws.q_h, engine=eng)
        ops.linear_add_layernorm(ws.q_h, q["f2_w"], q["f2_b"], ws.c_pos, *q["norm_feat"], 1e-5,              # :290-298
                                 out_f32=ws.q_new)
        return
    ops.linear(ws.att, q["o_w"], q["o_b"], out=ws.t, engine=eng)
    ops.add_layernorm(ws.t, ws.c_hs, *q["norm1"], 1e-5, out_f32=ws.q_tgt,                       # :277-278
                      out_lp=None if dt == torch.float32 else ws.q_tgt_lp2)
    ops.linear(ws.q_tgt_lp2, q["l1_w"], q["l1_b"], relu=True, out=ws.q_h, engine=eng)
    ops.linear(ws.q_h, q["l2_w"], q["l2_b"], out=ws.t, engine=eng)                               # :280
    ops.add_layernorm(ws.t, ws.q_tgt, *q["norm2"], 1e-5, want_f32=False, out_lp=ws.q_tgt_lp3)   # :281-282
    ops.linear(ws.q_tgt_lp3, q["f1_w"], q["f1_b"], relu=True, out=ws.q_h, engine=eng)
    ops.linear(ws.q_h, q["f2_w"], q["f2_b"], out=ws.t, engine=eng)                               # :290
    ops.add_layernorm(ws.t, ws.c_pos, *q["norm_feat"], 1e-5, out_f32=ws.q_new)                  # :294-298


def qim_update(weights: DecoderWeights, ref_pts: torch.Tensor, query_pos: torch.Tensor, out_embed: torch.Tensor,
               pred_boxes: torch.Tensor):
    """QueryInteractionModule._update_track_embedding (MOTR/models/qim.py:251-301) for one set of T tracks through
    the kernels the frame uses (`qim_update_ws`): ref_pts [T,4] logits, query_pos / out_embed [T,C] fp32,
    pred_boxes [T,4] -> (new query_pos [T,C] fp32, new ref_pts [T,4] = inverse_sigmoid(pred_boxes), qim.py:299)."""
    spec, dt, dev = weights.spec, weights.dt, ref_pts.device
    T, C = out_embed.shape
    R = max(16, (T + 15) // 16 * 16)
    ws = FrameWorkspace(R, 1, spec, dt, dev, weights.qim["l1_w"].shape[0])
    ws.c_hs[:T].copy_(out_embed)
    ws.c_pos[:T].copy_(query_pos)
    ops.pos2posemb(ref_pts.float().contiguous(), out=ws.q_pos[:T])                  # qim.py:255
    ops.add_cast(ws.c_hs[:T], ws.q_pos[:T], dt, out=ws.q_qk_lp[:T])                 # q = k = tgt + query_pos (:271)
    if dt != torch.float32:
        ops.add_cast(ws.c_hs[:T], None, dt, out=ws.q_tgt_lp[:T])
    ws.ro.copy_(torch.tensor([0, T], dtype=torch.int32))
    ws.n_active.fill_(T)
    qim_update_ws(weights, spec, ws, [0, 16 + R])
    return ws.q_new[:T].clone(), ops.inverse_sigmoid(pred_boxes.float().contiguous())


class _FramePlan:
    """Workspace (and optionally the captured CUDA graph) of one (padded frame size, input slot)."""
    __slots__ = ("rows_pad", "slot", "graph", "ws", "n_launch", "desc", "info", "frame_rows")


class TrackEngine:
    """Lock-step tracker for `n_seq` independent sequences on one GPU."""

    DEPTH = 4  # host result ring = host_lag + 2 (set per instance)

    def __init__(self, sd, spec: DecoderSpec, shapes, device, precision: str = "bf16", n_detect: int = 300,
                 n_seq: int = 1, score_thresh=0.4, filter_thresh=0.5, miss_tolerance=5, iou_thresh=0.8,
                 weights: Optional[DecoderWeights] = None, cap: int = 512, bucket: int = 64, margin: int = 32,
                 use_graphs: bool = True, table_rows: int = 1 << 18, branches: bool = True, selector=None,
                 value_ahead: Optional[bool] = None, gather_probe: Optional[ops.GatherProbe] = None,
                 cluster_decoder: Optional[bool] = None, static_tracks: Optional[int] = None, on_full: str = "raise",
                 host_lag: int = 2):
        # host_lag: how many frames the enqueueing host may run ahead of the newest frame whose results it has read
        # (>= 2). The device side is unchanged (two input / value slots, frame t's copy waits for frame t-2); a larger
        # lag only lets the host queue more launches, so a host stall shorter than (host_lag - 1) frame times no
        # longer drains the device queue. The speculative padded size then has to cover host_lag - 1 frames of births
        # (margin per stale frame); a wrong guess is re-run as before.
        if host_lag < 2:
            raise ValueError("host_lag must be >= 2")
        self.host_lag = int(host_lag)
        self.DEPTH = self.host_lag + 2
        # Fixed-size query memory (the stated purpose of MOTR/models/fsqm.py:8-10): with
        # `static_tracks=N` every sequence owns N track slots and EVERY frame runs with the same N + n_detect rows per
        # sequence -- one CUDA graph serves all frames, the host never needs the track counts to size a launch (no
        # speculative size, no abort / re-launch path). Unused slots are padding rows that no real row ever reads
        # (row counts live in device memory), so results are bit-identical to the dynamic engine while the tracks fit.
        # When more than N objects are alive: on_full="raise" (default) reports it, on_full="drop" applies FSQM's rule
        # "memory full -> the new query is not injected" (fsqm.py:78-81): surplus newborn objects are reported in the
        # frame they appear in but not carried.
        if on_full not in ("raise", "drop"):
            raise ValueError("on_full must be 'raise' or 'drop'")
        self.static_tracks, self.on_full = static_tracks, on_full
        if static_tracks is not None:
            cap = int(static_tracks)
        self.dev = torch.device(device)
        self.gather_probe = gather_probe   # instrumentation of every gather launch (bench.py roofline leg)
        self.launches = 0                  # kernels launched by this engine's frames (graph replays included)
        self.spec, self.shapes, self.n_detect, self.n_seq = spec, [list(s) for s in shapes], n_detect, n_seq
        self.Lv = level_sizes(shapes)
        self.W = weights or DecoderWeights(sd, spec, self.dev, precision)
        self.thr = (score_thresh, filter_thresh, miss_tolerance, iou_thresh)
        # whole decoder as one cluster kernel where the frame fits the device (see _cluster_rows)
        self._cd = None
        want_cd = ex.CLUSTER_DECODER if cluster_decoder is None else bool(cluster_decoder)
        if want_cd and ex.ClusterDecoder.supports(self.W.dt, spec):
            self._cd = ex.ClusterDecoder(self.W.layers, self.W.bbox, self.shapes, self.W.score_w, self.W.score_b)
        elif cluster_decoder:
            raise ValueError("cluster_decoder=True needs bf16, d_model 256, 8 heads, d_ffn 1024, 3 levels x 4 points")
        self.cap, self.bucket, self.margin, self.use_graphs = cap, bucket, margin, use_graphs
        self.branches = branches
        # split value projection (see _body): bf16 tcgen05 path with the persistent kernel's shape constraints
        self._vp_split = (self.W.dt == torch.bfloat16 and spec.d_model == 256 and ex._GEMM_ENGINE != _lib.GEMM_SIMT and
                          n_seq * self.Lv >= 4096 and os.environ.get("MOYOLO_VP_SPLIT", "1") != "0")
        self._vp_ctas = int(os.environ.get("MOYOLO_VP_CTAS", "72" if n_seq <= 2 else "0"))
        self._vp_gate = os.environ.get("MOYOLO_VP_GATE", "1") != "0"   # gate the ahead projection on the previous tail
        self._vp_gate_layer = int(os.environ.get("MOYOLO_VP_GATE_LAYER", str(spec.n_layers)))  # n_layers = the tail
        S, C, dev = n_seq, spec.d_model, self.dev
        # device-resident track state (fixed capacity)
        self.n_tracks = torch.zeros(S, dtype=torch.int32, device=dev)
        self.t_ref = torch.zeros(S, cap, 4, device=dev)
        self.t_qpos = torch.zeros(S, cap, C, device=dev)
        self.t_label = torch.zeros(S, cap, dtype=torch.int32, device=dev)
        self.t_ids = torch.zeros(S, cap, dtype=torch.int64, device=dev)
        self.t_dis = torch.zeros(S, cap, dtype=torch.int64, device=dev)
        self.counters = torch.zeros(S, 2, dtype=torch.int64, device=dev)
        self.ctrl = torch.zeros(8, dtype=torch.int32, device=dev)
        self.seq_ids = torch.arange(S, dtype=torch.int32, device=dev)
        self.table = torch.zeros(table_rows, 9, device=dev)
        # 2-deep input ring (graphs read these) and the all-layers value tensor shared by every plan
        # `selector` (moyolo_b200.selector.QuerySelector): the frame then starts from the neck's
        # channels-last maps; feats / detect queries are produced inside the frame graph.
        self.selector = selector
        if selector is None:
            self.feats_in = [torch.zeros(S, self.Lv, C, dtype=self.W.dt, device=dev) for _ in range(2)]
            self.det_embed_in = [torch.zeros(S, n_detect, C, device=dev) for _ in range(2)]
            self.det_refer_in = [torch.zeros(S, n_detect, 4, device=dev) for _ in range(2)]
            self._ring = [[self.feats_in[i], self.det_embed_in[i], self.det_refer_in[i]] for i in range(2)]
        else:
            assert selector.S == S and selector.nq == n_detect and selector.dt == self.W.dt
            # Query selection AHEAD of the frame (see _prelude): input projection, value projection and the top-k
            # selection of frame t+1 run as their own small graph on a side stream while frame t is still decoding
            # (the latency-bound decoder leaves most SMs idle); their outputs are then double-buffered by frame parity.
            self._sel_ahead = (use_graphs and self.W.dt == torch.bfloat16 and value_ahead is not False and
                               os.environ.get("MOYOLO_SEL_AHEAD", "1") != "0")
            nb = 2 if self._sel_ahead else 1
            fi = [torch.zeros(S, self.Lv, C, dtype=self.W.dt, device=dev) for _ in range(nb)]
            de = [torch.zeros(S, n_detect, C, device=dev) for _ in range(nb)]
            dr = [torch.zeros(S, n_detect, 4, device=dev) for _ in range(nb)]
            self.feats_in, self.det_embed_in, self.det_refer_in = [fi[i % nb] for i in range(2)], \
                [de[i % nb] for i in range(2)], [dr[i % nb] for i in range(2)]
            self._ring = [[torch.zeros(shp, dtype=self.W.dt, device=dev) for shp in selector.map_shapes()]
                          for _ in range(2)]
        # Value projection AHEAD of the frame (bf16 tall path, frame inputs = feats): it is not captured in the frame
        # graph but launched at submit time on its own stream, gated on an event the PREVIOUS frame's graph records
        # when its last decoder layer has been issued -- so it runs on the SMs the previous frame's tail (ID
        # assignment, QIM: a chain of tiny kernels) leaves idle instead of in front of this frame's first gather.
        # Needs the all-layers value tensor double-buffered by frame parity.
        self._vp_ahead = (selector is None and self._vp_split and os.environ.get("MOYOLO_VP_AHEAD", "1") != "0" and
                          value_ahead is not False)
        if selector is None:
            self._sel_ahead = False
        if value_ahead and not (self._vp_ahead or self._sel_ahead):
            raise ValueError("value_ahead needs the bf16 tcgen05 path (and S*Lv >= 4096 when the inputs are feats)")
        self._pre_graphs = [None, None]   # per input slot: captured prelude graph (selection-ahead mode)
        n_val = 2 if (self._vp_ahead or self._sel_ahead) else 1
        self.values_buf = [torch.zeros(S, self.Lv, spec.n_layers * C, dtype=self.W.dt, device=dev) for _ in range(n_val)]
        self.values = self.values_buf[0]
        # streams / events
        self._main = torch.cuda.current_stream(dev)
        self._copy = torch.cuda.Stream(dev)
        self._s_val = torch.cuda.Stream(dev)
        self._s_box = torch.cuda.Stream(dev)
        self._out = torch.cuda.Stream(dev)      # result copies (device -> pinned host), off the main stream
        self._ev_done = [torch.cuda.Event() for _ in range(self.DEPTH)]
        self._ev_copy = [torch.cuda.Event() for _ in range(2)]
        self._ev_vp = [torch.cuda.Event() for _ in range(2)]                 # value projection of slot done
        # recorded INSIDE the frame graphs, waited on from outside: raw events of the library (an external
        # event-record node while capturing, a plain record otherwise)
        self._ev_tail = [_lib.lib().moyolo_event_create() for _ in range(2)]
        # pinned host rings: [n_active (S) | ctrl (8)] and the packed frame rows
        self._max_rows = self._round(S * (n_detect + cap))
        self._static_rows = self._max_rows if static_tracks is not None else None
        self._h_info = torch.zeros(self.DEPTH, S + 8, dtype=torch.int32).pin_memory()
        self._h_rows = torch.zeros(self.DEPTH, self._max_rows, 8).pin_memory()
        self._plans: Dict[tuple, _FramePlan] = {}
        # native submission (moyolo_frame_submit): raw handles of the streams / events above. torch creates
        # the underlying cudaEvent lazily on the first record, so every event is recorded once here.
        self._ev_scratch = torch.cuda.Event()
        self._ev_graph = torch.cuda.Event()
        for e in (*self._ev_done, *self._ev_copy, self._ev_scratch, self._ev_graph, *self._ev_vp):
            e.record(self._main)
        torch.cuda.synchronize(dev)
        self._native = use_graphs and os.environ.get("MOYOLO_NATIVE_SUBMIT", "1") != "0"
        self._h_info_np = self._h_info.numpy()
        self._host_reset()

    def __del__(self):
        try:
            for e in getattr(self, "_ev_tail", []):
                _lib.lib().moyolo_event_destroy(e)
        except Exception:  # interpreter shutdown
            pass

    # ---- host-side bookkeeping ---------------------------------------------------------------
    def _host_reset(self):
        self._T = [0] * self.n_seq          # exact track counts after frame `_known`
        self._known = -1                    # newest frame whose counts the host has read
        self._growth: List[int] = []        # net track growth of the last harvested frames
        self._next = 0                      # next frame index to submit
        self._inflight: List[dict] = []     # submitted, not yet harvested (oldest first)
        self._T_before: Dict[int, List[int]] = {0: [0] * self.n_seq}  # frame -> track counts it started with
        self._last_plan: Optional[_FramePlan] = None
        self.frame_idx = 0
        self.aborts = 0

    def reset(self, seq: Optional[int] = None):
        """is_first semantics (head.py:199-205): drop all tracks, fresh ID counters (and, for a full
        reset, an empty track table and frame counter)."""
        self.drain()
        if seq is None:
            self.n_tracks.zero_()
            self.counters.zero_()
            self.ctrl.zero_()
            self._host_reset()
        else:
            self.n_tracks[seq] = 0
            self.counters[seq].zero_()
            self._T[seq] = 0
            # collect() of the next frame slices its packed rows with the counts that frame STARTED with
            self._T_before[self._next] = list(self._T)

    def set_seq_ids(self, ids) -> None:
        """Global sequence index of every lock-step slot (the `seq` column of the emitted rows)."""
        self.drain()
        self.seq_ids.copy_(torch.as_tensor(list(ids), dtype=torch.int32))

    def n_tracks_host(self) -> List[int]:
        self.drain()
        return list(self._T)

    n_tracks_list = n_tracks_host

    def track_ids(self, s: int) -> torch.Tensor:
        self.drain()
        return self.t_ids[s, :self._T[s]]

    def track_disappear(self, s: int) -> torch.Tensor:
        self.drain()
        return self.t_dis[s, :self._T[s]]

    def track_table(self) -> torch.Tensor:
        """Rows [seq, frame, id, cx, cy, w, h, score, cls] of every tracked object emitted so far
        (device tensor, ordered by frame, then slot, then query order)."""
        self.drain()
        info = self.ctrl.cpu()
        if int(info[CTRL_TABLE_OVERFLOW]):
            raise RuntimeError(f"moyolo_b200: track table capacity ({self.table.shape[0]} rows) exceeded; "
                               "construct TrackEngine with a larger table_rows")
        return self.table[:int(info[CTRL_CURSOR])]

    def track_table_device(self):
        """(table buffer [table_rows, 9], device int32 row count, device int32 overflow flag) WITHOUT waiting for the
        in-flight frames: work enqueued on the current stream after this call sees the table as of the last
        submitted frame (sharding.run_sharded packs it for the final gather with no host read). The caller must
        drain() afterwards and repeat if `aborts` changed (a speculative frame was re-run)."""
        torch.cuda.current_stream(self.dev).wait_event(self._ev_done[(self._next - 1) % self.DEPTH])
        return (self.table, self.ctrl[CTRL_CURSOR:CTRL_CURSOR + 1], self.ctrl[CTRL_TABLE_OVERFLOW:CTRL_TABLE_OVERFLOW + 1])

    # ---- one frame: every launch ----------------------------------------------------------------
    def _body(self, p: _FramePlan) -> None:
        """All launches of one frame for plan `p` (graph-capturable: no host sync, no allocation, shapes
        depend on rows_pad only; counts are read from device memory by the kernels)."""
        W, S, nd, C, ws, R = self.W, self.n_seq, self.n_detect, self.spec.d_model, p.ws, p.rows_pad
        dt, spec, n_l = W.dt, self.spec, self.spec.n_layers
        cur = torch.cuda.current_stream(self.dev)
        feats = self.feats_in[p.slot].view(S * self.Lv, C)
        fork = self.branches
        if self.selector is not None and not self._sel_ahead:  # input projection of the neck maps -> feats (head.py:1012-1029)
            self.selector.project(self._ring[p.slot], self.feats_in[p.slot])
        # side branch: value projection of ALL layers (transformer.py:264; feats is the same tensor in every layer,
        # transformer.py:705). With the persistent tcgen05 kernel it is issued as two launches: layer 0's column
        # slice on every SM (the first gather waits for nothing else), then layers 1.. on a CTA budget that leaves
        # SMs free for the latency-bound main chain running next to it (they are needed one layer later).
        values = self.values_buf[p.slot % len(self.values_buf)]
        values2d = values.view(S * self.Lv, n_l * C)
        split = self._vp_split and n_l > 1
        self._ev_v0 = None

        def value_proj():
            if self._vp_ahead or self._sel_ahead:   # launched at submit time (see _value_ahead), not part of the frame graph
                return
            if split:
                ops.linear_tall(feats, W.value_proj.w[:C], W.value_proj.b[:C], values2d[:, :C])
                if fork:
                    self._ev_v0 = torch.cuda.Event()
                    self._ev_v0.record(torch.cuda.current_stream(self.dev))
                ops.linear_tall(feats, W.value_proj.w[C:], W.value_proj.b[C:], values2d[:, C:], max_ctas=self._vp_ctas)
            else:
                ops.linear(feats, W.value_proj.w, W.value_proj.b, out=values2d, engine=ex._GEMM_ENGINE)

        if fork:
            self._s_val.wait_stream(cur)
            with torch.cuda.stream(self._s_val):
                value_proj()
        else:
            value_proj()
        if self.selector is not None and not self._sel_ahead:  # detect queries of this frame (head.py:1031-1113)
            self.selector.select(self.feats_in[p.slot], self.det_embed_in[p.slot], self.det_refer_in[p.slot])
        ops.frame_assemble(S, nd, C, self.cap, self.n_tracks, self.t_ref, self.t_qpos, self.t_label, self.t_ids,
                           self.t_dis, W.class_embed, self.det_embed_in[p.slot], self.det_refer_in[p.slot], ws.x,
                           ws.refer_logit, ws.pos, ws.ids0, ws.dis0, ws.ro, R, ctrl=self.ctrl,
                           refer_sig=ws.refer[0],                            # transformer.py:690
                           x_lp=None if dt == torch.float32 else ws.x_lp, xq_lp=ws.xq_lp)
        # host-side bound of the device offsets, used for grid sizing only: sum_s ceil(N_s/16) <= R/16 + S
        ro_host = [16 * i for i in range(S)] + [16 * S + R]
        # class-score head inside the last layer's FFN2 GEMM+LayerNorm launch (bf16 fused-epilogue path, nc <= 8)
        score_fused = ops.SCORE_FUSED and ex.fused_epilogues(dt, C) and spec.nc <= 8 and \
            not ops.ffn_fused_supported(dt, C, W.layers[-1].ffn1.w.shape[0])
        m_rows = self._cluster_rows(R)
        if m_rows:
            # ONE launch for all layers (self-attention, deformable attention, FFN, LayerNorms, box refinement) and the
            # score head: csrc/decoder_cluster.cu. The residual stream is updated in place (a cluster reads and writes
            # only its own rows of ws.x).
            if fork and not (self._vp_ahead or self._sel_ahead):
                cur.wait_stream(self._s_val)   # all layers' value slices
            self._cd.run(ws.x, ws.pos, ws.refer[0], values, ws.ro, S, R, m_rows, ws.x, ws.kv, ws.grid_bar,
                         ws.refer[1:], x_lp_out=ws.x_lp, logits=ws.logits, scores=ws.scores, labels=ws.labels,
                         status=ws.dc_status)
            score_fused = True   # (scores, labels and logits are already there)
        else:
            self._layer_chain(p, values, ro_host, score_fused, fork, cur)
        if self._vp_ahead and (m_rows or self._vp_gate_layer >= n_l):  # the next frame's value projection may start now: only the tail is left
            _lib.check(_lib.lib().moyolo_event_record(self._ev_tail[p.slot], cur.cuda_stream))
        boxes = ws.refer[n_l]
        if not score_fused:
            ops.score_head(ws.x_lp, W.score_w, W.score_b, out=(ws.logits, ws.scores, ws.labels))
        st, ft, mt, it = self.thr
        # ID assignment (head.py:1232-1243) + active-track selection/compaction (qim.py:184-187) in one launch. With
        # side branches it does not touch the refined boxes (frame_writeback gathers them through the selection), so
        # neither it nor the QIM update waits for the last layer's box head running on the side branch.
        ops.frame_assign_compact(S, C, self.cap, R, ws.ro, ws.scores, ws.ids0, ws.dis0, self.counters, ws.ids, ws.dis,
                                 ws.labels, ws.refer_logit, ws.pos, ws.x, None if fork else boxes, ws.n_active,
                                 ws.active_index, ws.c_ref, ws.c_pos, ws.c_hs, None if fork else ws.c_box, self.t_label,
                                 self.t_ids, self.t_dis, st, ft, mt, ctrl=self.ctrl, q_qk_lp=ws.q_qk_lp,   # qim.py:255, 271
                                 q_tgt_lp=None if dt == torch.float32 else ws.q_tgt_lp)

        def side():
            # neither the duplicate filter (it only moves the ID counters, head.py:1268-1283) nor the frame's result
            # rows / track-table append are needed by the QIM update
            ops.track_suppress_batched(boxes, ws.ids, self.counters, ws.ro, S, ws.rows_per_seq, ws.assign_ws, it,
                                       ctrl=self.ctrl)
            ops.frame_emit(S, R, ws.ro, ws.ids, boxes, ws.scores, ws.labels, ws.n_active, ws.active_index,
                           self.seq_ids, p.frame_rows, self.table, self.ctrl)

        if fork:
            self._s_box.wait_stream(cur)   # (the last box head is already queued on this branch)
            with torch.cuda.stream(self._s_box):
                side()
        else:
            side()
        self._qim_update(ws, ro_host)
        if fork:
            cur.wait_stream(self._s_box)   # final boxes (and the rest of the side branch)
        # write-back also stores what the host reads back after the frame: [n_active | ctrl]
        ops.frame_writeback(S, C, self.cap, ws.ro, ws.n_active, ws.q_new, None if fork else ws.c_box, self.t_qpos,
                            self.t_ref, self.n_tracks, ctrl=self.ctrl, info=p.info, boxes=boxes if fork else None,
                            active_index=ws.active_index if fork else None)

    def _cluster_rows(self, rows_pad: int) -> int:
        """Tile height (32 / 64 rows) of the one-launch cluster decoder for a frame of rows_pad rows, 0 = the frame
        does not fit the co-resident clusters / key staging of this device (launch-chained schedule then)."""
        if self._cd is None or self.gather_probe is not None:
            return 0
        return self._cd.tile_rows(rows_pad, self.n_seq, rows_pad - (self.n_seq - 1) * self.n_detect)

    def _layer_chain(self, p: "_FramePlan", values, ro_host, score_fused: bool, fork: bool, cur) -> None:
        """Launch-chained schedule of the decoder layers (7 launches per layer on the main chain + a 3-launch box
        head side branch): frames that do not fit the cluster decoder, fp32, and the instrumented roofline leg."""
        W, S, C, ws, R = self.W, self.n_seq, self.spec.d_model, p.ws, p.rows_pad
        dt, n_l = W.dt, self.spec.n_layers
        for i, pk in enumerate(W.layers):
            last = i + 1 == n_l
            value_view = values[:, :, i * C:(i + 1) * C]

            def before_gather(i=i):
                if not fork:
                    return
                if i == 0:
                    if self._ev_v0 is not None:
                        cur.wait_event(self._ev_v0)   # layer 0's value slice only
                    else:
                        cur.wait_stream(self._s_val)
                else:
                    if i == 1 and self._ev_v0 is not None:
                        cur.wait_stream(self._s_val)  # the other layers' slices
                    cur.wait_stream(self._s_box)   # refer[i] comes from the box head of layer i-1

            if self._vp_ahead and i == self._vp_gate_layer:   # experiment knob: release the next projection earlier
                _lib.check(_lib.lib().moyolo_event_record(self._ev_tail[p.slot], cur.cuda_stream))
            ex.run_layer_ws(pk, ws, ws.refer[i].view(R, 1, 4), value_view, self.shapes, S, ws.ro, ro_host,
                            ws.pos, None if last else ws.pos, dt, before_gather,
                            score=(W.score_w, W.score_b, ws.logits, ws.scores, ws.labels) if (last and score_fused) else None,
                            gather_probe=self.gather_probe)
            # box refinement of this layer (transformer.py:709): only the NEXT layer's gather needs it
            # (the last layer's box head runs next to the score head, which only needs the layer output)
            if fork:
                self._s_box.wait_stream(cur)
                with torch.cuda.stream(self._s_box):
                    ex.bbox_head_ws(W.bbox[i], ws, ws.refer[i], ws.refer[i + 1])
            else:
                ex.bbox_head_ws(W.bbox[i], ws, ws.refer[i], ws.refer[i + 1])

    def _qim_update(self, ws: FrameWorkspace, ro_host) -> None:
        qim_update_ws(self.W, self.spec, ws, ro_host)

    # ---- plans ------------------------------------------------------------------------------------
    def _state_snapshot(self):
        return [t.clone() for t in (self.n_tracks, self.t_ref, self.t_qpos, self.t_label, self.t_ids, self.t_dis,
                                    self.counters, self.ctrl)]

    def _state_restore(self, snap):
        for dst, src in zip((self.n_tracks, self.t_ref, self.t_qpos, self.t_label, self.t_ids, self.t_dis,
                             self.counters, self.ctrl), snap):
            dst.copy_(src)

    def _plan(self, rows_pad: int, slot: int) -> _FramePlan:
        p = self._plans.get((rows_pad, slot))
        if p is not None:
            return p
        p = _FramePlan()
        p.rows_pad, p.slot, p.graph, p.n_launch, p.desc = rows_pad, slot, None, 0, None
        other = self._plans.get((rows_pad, slot ^ 1))
        # both input slots of one size share a workspace: frames are serialised on the main stream
        p.ws = other.ws if other is not None else FrameWorkspace(rows_pad, self.n_seq, self.spec, self.W.dt, self.dev,
                                                                 self.W.qim["l1_w"].shape[0], self.n_detect + self.cap)
        # what the host reads back is double-buffered per input slot (frame parity): the result copies run on their
        # own stream while the next frame's graph (other slot) is already executing
        p.info = torch.zeros(self.n_seq + 8, dtype=torch.int32, device=self.dev)
        p.frame_rows = torch.zeros(rows_pad, 8, device=self.dev)
        if self.use_graphs:
            torch.cuda.synchronize(self.dev)
            snap = self._state_snapshot()
            side = torch.cuda.Stream(self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):  # one eager pass: loads modules, sets kernel attributes
                self._body(p)
            torch.cuda.current_stream(self.dev).wait_stream(side)
            torch.cuda.synchronize(self.dev)
            self._state_restore(snap)
            g = torch.cuda.CUDAGraph()
            before = ops.launch_count()
            with torch.cuda.graph(g):
                self._body(p)
            p.graph = g
            p.desc = self._make_desc(p)
            p.n_launch = ops.launch_count() - before  # kernel nodes replayed by every graph launch
            self._state_restore(snap)  # capture does not execute, but keep the invariant explicit
            torch.cuda.synchronize(self.dev)
        self._plans[(rows_pad, slot)] = p
        return p

    def _make_desc(self, p: _FramePlan):
        """Pre-filled moyolo_frame_submit_t of a plan: only the per-frame fields change at submit time."""
        d = _lib.FrameSubmit()
        d.copy_stream, d.main_stream, d.main_stream_valid = self._copy.cuda_stream, self._main.cuda_stream, 1
        d.ev_scratch = self._ev_scratch.cuda_event
        d.ev_copy = self._ev_copy[p.slot].cuda_event
        d.graph_exec = p.graph.raw_cuda_graph_exec()
        d.n_inputs = 3
        for i, t in enumerate(self._ring[p.slot]):
            d.in_dst[i] = t.data_ptr()
            d.in_bytes[i] = t.numel() * t.element_size()
        d.out_src[0], d.out_bytes[0] = p.info.data_ptr(), p.info.numel() * 4
        d.out_src[1], d.out_bytes[1] = p.frame_rows.data_ptr(), p.rows_pad * 8 * 4
        d.out_stream, d.out_stream_valid, d.ev_graph = self._out.cuda_stream, 1, self._ev_graph.cuda_event
        if self._sel_ahead:
            d.vp_valid, d.vp_stream = 2, self._s_val.cuda_stream
            d.ev_vp = self._ev_vp[p.slot].cuda_event
            d.pre_graph_exec = self._prelude(p.slot).raw_cuda_graph_exec()
        if self._vp_ahead:
            S, C, n_l = self.n_seq, self.spec.d_model, self.spec.n_layers
            d.vp_valid, d.vp_stream = 1, self._s_val.cuda_stream
            d.ev_vp = self._ev_vp[p.slot].cuda_event
            d.vp_x, d.vp_ldx = self.feats_in[p.slot].data_ptr(), C
            d.vp_w, d.vp_bias = self.W.value_proj.w.data_ptr(), self.W.value_proj.b.data_ptr()
            d.vp_y, d.vp_ldy = self.values_buf[p.slot].data_ptr(), n_l * C
            d.vp_M, d.vp_N, d.vp_max_ctas = S * self.Lv, n_l * C, self._vp_ctas
        return d

    def _submit_native(self, t: int, rows_pad: int, feats, det_embed, det_refer, want_rows: bool,
                       sync_inputs: bool) -> dict:
        """Input copies, graph launch, result copies and event hand-offs of one frame as ONE C call."""
        slot, h = t % 2, t % self.DEPTH
        p = self._plan(rows_pad, slot)
        d = p.desc
        d.ev_slot_free = self._ev_done[(t - 2) % self.DEPTH].cuda_event if t >= 2 else None
        d.sync_inputs = 1 if sync_inputs else 0
        d.ev_done = self._ev_done[h].cuda_event
        d.ev_tail_prev = self._ev_tail[(t - 1) % 2] if (self._vp_ahead and self._vp_gate and t > 0) else None
        d.in_src[0], d.in_src[1], d.in_src[2] = feats.data_ptr(), det_embed.data_ptr(), det_refer.data_ptr()
        d.out_dst[0] = self._h_info[h].data_ptr()
        d.out_dst[1] = self._h_rows[h].data_ptr()
        d.n_outputs = 2 if want_rows else 1
        _lib.check(_lib.lib().moyolo_frame_submit(ctypes.byref(d)))
        self.launches += p.n_launch + (1 if self._vp_ahead else 0) + (self._pre_launches if self._sel_ahead else 0)
        self._last_plan = p
        return {"frame": t, "plan": p, "rows_pad": rows_pad, "want_rows": want_rows}

    def _native_ok(self, feats, det_embed, det_refer) -> bool:
        return self._native and all(x.dtype == r.dtype and x.is_contiguous() and x.numel() == r.numel() and
                                    (x.is_cuda or x.is_pinned())
                                    for x, r in zip((feats, det_embed, det_refer), self._ring[0]))

    def prepare(self, max_tracks_per_seq: int) -> int:
        """Pre-capture the frame graphs for every padded size up to `max_tracks_per_seq` tracks per
        sequence (done at start-up so no capture happens in the frame loop). Returns #graphs."""
        if not self.use_graphs:
            return 0
        self.drain()
        if self._static_rows is not None:
            self._plan(self._static_rows, 0)
            self._plan(self._static_rows, 1)
            return len(self._plans)
        lo = self.n_seq * self.n_detect
        hi = self.n_seq * (self.n_detect + max_tracks_per_seq) + self.margin * (self.host_lag - 1)
        r = self._round(lo)
        while r <= self._round(hi):
            self._plan(r, 0)
            self._plan(r, 1)
            r += self.bucket
        return len(self._plans)

    def _round(self, rows: int) -> int:
        b = self.bucket
        return (rows + b - 1) // b * b

    # ---- pipelined frame loop ---------------------------------------------------------------------
    def _load_inputs(self, slot: int, feats, det_embed, det_refer, frame: int, sync_inputs: bool) -> None:
        """Copy (device or pinned-host) frame inputs into ring slot `slot` on the copy stream, after the
        frame that last read this slot (frame - 2) has finished."""
        cs = self._copy
        if frame >= 2:
            cs.wait_event(self._ev_done[(frame - 2) % self.DEPTH])
        if sync_inputs:  # inputs still being produced on the caller's stream
            cs.wait_stream(self._main)
        for t in (feats, det_embed, det_refer):
            if t.is_cuda:
                t.record_stream(cs)  # the caller may drop its reference before the async copy has run
        with torch.cuda.stream(cs):
            for src, dst in zip((feats, det_embed, det_refer), self._ring[slot]):
                if src.dtype == dst.dtype or not src.is_cuda:
                    dst.copy_(src.reshape(dst.shape), non_blocking=True)
                else:  # dtype conversion on device through the library's own cast kernel
                    ops.add_cast(src.reshape(-1).float().contiguous(), None, dst.dtype, out=dst.view(-1))
            self._ev_copy[slot].record(cs)
        self._value_ahead(slot, frame)

    def _prelude(self, slot: int) -> torch.cuda.CUDAGraph:
        """Selection-ahead mode: everything of a frame that depends on its INPUTS only -- input projection of the neck
        maps (head.py:1012-1029), the all-layers value projection (transformer.py:264) and the encoder-side query
        selection (head.py:1031-1113) -- captured once per input slot as its own graph."""
        g = self._pre_graphs[slot]
        if g is not None:
            return g
        S, C, n_l, sel = self.n_seq, self.spec.d_model, self.spec.n_layers, self.selector

        def run():
            cur = torch.cuda.current_stream(self.dev)
            sel.project(self._ring[slot], self.feats_in[slot])
            self._s_box.wait_stream(cur)   # value projection next to the selection chain
            with torch.cuda.stream(self._s_box):
                ops.linear(self.feats_in[slot].view(S * self.Lv, C), self.W.value_proj.w, self.W.value_proj.b,
                           out=self.values_buf[slot].view(S * self.Lv, n_l * C), engine=ex._GEMM_ENGINE)
            sel.select(self.feats_in[slot], self.det_embed_in[slot], self.det_refer_in[slot])
            cur.wait_stream(self._s_box)

        torch.cuda.synchronize(self.dev)
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):   # eager pass: loads modules, sets kernel attributes
            run()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        before = ops.launch_count()
        with torch.cuda.graph(g):
            run()
        self._pre_launches = ops.launch_count() - before
        torch.cuda.synchronize(self.dev)
        self._pre_graphs[slot] = g
        return g

    def _value_ahead(self, slot: int, frame: int) -> None:
        """Value projection (or, in selection-ahead mode, the whole prelude graph) of the frame whose inputs were
        just enqueued for ring slot `slot` (python path; the native submission does the same inside
        moyolo_frame_submit)."""
        if self._sel_ahead:
            g = self._prelude(slot)
            vs = self._s_val
            vs.wait_event(self._ev_copy[slot])
            with torch.cuda.stream(vs):
                g.replay()
                self._ev_vp[slot].record(vs)
            self.launches += self._pre_launches
            self._main.wait_event(self._ev_vp[slot])
            return
        if not self._vp_ahead:
            return
        vs = self._s_val
        vs.wait_event(self._ev_copy[slot])
        if frame > 0 and self._vp_gate:
            _lib.check(_lib.lib().moyolo_stream_wait_event(vs.cuda_stream, self._ev_tail[(frame - 1) % 2]))
        S, C, n_l = self.n_seq, self.spec.d_model, self.spec.n_layers
        with torch.cuda.stream(vs):
            ops.linear_tall(self.feats_in[slot].view(S * self.Lv, C), self.W.value_proj.w, self.W.value_proj.b,
                            self.values_buf[slot].view(S * self.Lv, n_l * C), max_ctas=self._vp_ctas)
            self._ev_vp[slot].record(vs)
        self.launches += 1
        self._main.wait_event(self._ev_vp[slot])

    def _launch(self, frame: int, rows_pad: int, want_rows: bool) -> dict:
        slot = frame % 2
        p = self._plan(rows_pad, slot)
        main = self._main
        h = frame % self.DEPTH
        with torch.cuda.stream(main):
            main.wait_event(self._ev_copy[slot])
            if self.use_graphs:
                p.graph.replay()
                self.launches += p.n_launch
            else:
                before = ops.launch_count()
                self._body(p)
                self.launches += ops.launch_count() - before
            # the frame's host-visible results: [active-track counts | control block] and, optionally, the
            # packed rows -- one small device->host copy each
            self._h_info[h].copy_(p.info, non_blocking=True)
            if want_rows:
                self._h_rows[h, :rows_pad].copy_(p.frame_rows, non_blocking=True)
            self._ev_done[h].record(main)
        self._last_plan = p
        return {"frame": frame, "plan": p, "rows_pad": rows_pad, "want_rows": want_rows}

    def _harvest(self, upto: int, block: bool) -> None:
        """Read the results of in-flight frames <= upto (oldest first). A frame whose speculative size
        did not fit is re-run (with every later in-flight frame) using the exact counts."""
        while self._inflight and self._inflight[0]["frame"] <= upto:
            rec = self._inflight[0]
            ev = self._ev_done[rec["frame"] % self.DEPTH]
            if not block and not ev.query():
                return
            ev.synchronize()
            info = self._h_info_np[rec["frame"] % self.DEPTH]
            if int(info[self.n_seq + CTRL_ABORT]) != 0:
                self._recover()
                continue
            if int(info[self.n_seq + CTRL_TRACK_OVERFLOW]) != 0 and self.on_full == "raise":
                raise RuntimeError(f"moyolo_b200: a sequence carried more than cap={self.cap} active tracks in frame "
                                   f"{rec['frame']}; the surplus tracks lost their identity. Construct TrackEngine with "
                                   "a larger cap and re-run the sequence")
            self._inflight.pop(0)
            grown = int(info[:self.n_seq].sum()) - sum(self._T)
            self._growth = (self._growth + [grown])[-8:]   # net births of the last frames: sizes the speculation
            self._T = info[:self.n_seq].tolist()
            self._known = rec["frame"]
            self._T_before[rec["frame"] + 1] = list(self._T)
            self._T_before.pop(rec["frame"] - self.DEPTH, None)

    def _recover(self) -> None:
        """The oldest in-flight frame aborted on the device (its padded size was too small): nothing it or
        any later frame did touched the track state. Re-launch them in order with exact sizes."""
        torch.cuda.synchronize(self.dev)
        redo = self._inflight
        self._inflight = []
        self.ctrl[CTRL_ABORT:CTRL_ABORT + 1].zero_()
        self.aborts += 1
        for rec in redo:
            rows = sum(self._T) + self.n_seq * self.n_detect
            if self.host_lag > 2:   # the two input slots have moved on to later frames: load this frame's inputs again
                self._load_inputs(rec["frame"] % 2, *rec["keep"], rec["frame"], False)
            new = self._launch(rec["frame"], self._round(rows), rec["want_rows"])
            new["keep"] = rec["keep"]
            self._inflight.append(new)
            self._harvest(rec["frame"], block=True)

    def submit(self, feats: torch.Tensor, det_embed: torch.Tensor, det_refer: torch.Tensor,
               want_rows: bool = True, sync_inputs: bool = False) -> int:
        """Enqueue the next frame without waiting for it. feats [n_seq, Lv, C] (GEMM dtype or fp32;
        device or pinned host), det_embed [n_seq, nd, C] fp32, det_refer [n_seq, nd, 4] fp32 logit-space
        boxes. With a `selector` the three arguments are instead the neck's channels-last maps
        [n_seq, H_l, W_l, C_l] of the three pyramid levels (GEMM dtype). Unless sync_inputs is set the tensors must be complete when submit() is called (their copy
        runs on the engine's copy stream, which is then not ordered after the caller's stream), and they
        must not be modified until `host_lag` (default 2) more frames have been submitted. Returns the frame index
        for `collect`."""
        t = self._next
        self._harvest(t - self.host_lag, block=True)    # at most host_lag frames in flight
        self._harvest(t - 1, block=False)   # use the newest counts if they are already here
        if self._static_rows is not None:   # fixed-size query memory: the same launch for every frame
            rows_pad = self._static_rows
        else:
            rows = sum(self._T) + self.n_seq * self.n_detect
            stale = t - 1 - self._known     # frames whose births the host has not seen yet
            # per stale frame: the largest net growth of the last 8 frames + 4, at most `margin` (a wrong guess is
            # only slower, never wrong: the device aborts the frame and it is re-run with the exact size)
            per = self.margin if len(self._growth) < 3 else min(self.margin, max(self._growth) + 4)
            rows_pad = self._round(rows + max(per, 0) * stale)
            rows_pad = min(rows_pad, self._max_rows)
        if self._native_ok(feats, det_embed, det_refer):
            rec = self._submit_native(t, rows_pad, feats, det_embed, det_refer, want_rows, sync_inputs)
        else:
            self._load_inputs(t % 2, feats, det_embed, det_refer, t, sync_inputs)
            rec = self._launch(t, rows_pad, want_rows)
        rec["keep"] = (feats, det_embed, det_refer)   # alive until the frame has finished (and for a re-run)
        self._inflight.append(rec)
        self._next = t + 1
        self.frame_idx = self._next
        return t

    def drain(self) -> None:
        """Wait for every submitted frame (host counts become exact)."""
        self._harvest(self._next - 1, block=True)

    def collect(self, frame: int) -> List[Dict[str, torch.Tensor]]:
        """Host-side results of a submitted frame (submitted with want_rows=True; at most DEPTH-2 frames
        back): per sequence a dict of CPU tensors ids (int64, -1 = no object), boxes (cx,cy,w,h normalised),
        scores, labels for its N = T + n_detect rows, in query order. Views of a pinned ring: copy
        them if they must outlive the next two submits."""
        self._harvest(frame, block=True)
        rows = self._h_rows[frame % self.DEPTH]
        Tb = self._T_before.get(frame)   # exact once the previous frame has been harvested
        if Tb is None or frame < self._next - (self.DEPTH - 2):
            raise RuntimeError("collect(): frame results are no longer available")
        outs, off = [], 0
        for s in range(self.n_seq):
            n = Tb[s] + self.n_detect
            r = rows[off:off + n]
            outs.append({"ids": r[:, 0].long(), "boxes": r[:, 1:5], "scores": r[:, 5], "labels": r[:, 6].int()})
            off += n
        return outs

    def step(self, feats: torch.Tensor, det_embed: torch.Tensor, det_refer: torch.Tensor) -> List[Dict[str, torch.Tensor]]:
        """Synchronous frame: submit + wait. Returns one dict per sequence with the N = T + nd rows of this
        frame as DEVICE tensors: ids (int64, -1 = no object), boxes (cx,cy,w,h normalised), scores, labels,
        logits. They are views of the plan's workspace: consume or copy them before the next frame."""
        S, nd = self.n_seq, self.n_detect
        self.drain()
        T_in = list(self._T)
        t = self.submit(feats, det_embed, det_refer, want_rows=False, sync_inputs=True)
        self._harvest(t, block=True)
        ws = self._last_plan.ws
        outs, off = [], 0
        boxes = ws.refer[self.spec.n_layers]
        for s in range(S):
            n = T_in[s] + nd
            outs.append({"ids": ws.ids[off:off + n], "boxes": boxes[off:off + n], "scores": ws.scores[off:off + n],
                         "labels": ws.labels[off:off + n], "logits": ws.logits[off:off + n]})
            off += n
        return outs


class SequenceTracker(TrackEngine):
    """Single-sequence convenience wrapper: 2-D inputs, one dict out."""

    def __init__(self, sd, spec, shapes, device, precision="bf16", n_detect=300, **kw):
        super().__init__(sd, spec, shapes, device, precision, n_detect, 1, **kw)

    def step(self, feats, det_embed, det_refer):  # type: ignore[override]
        return super().step(feats[None], det_embed[None], det_refer[None])[0]
