"""ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product package `moyolo_b200`.

CPU restatement of the track-query update of DecoderTracker, in numpy float32 scalars / Python loops:

  * `TrackerPort.update`  — RuntimeTrackerBase.update, ultralytics/nn/modules/head.py:1201-1283, with
    `_filter_tracks` (:1155-1171) and `_calculate_iou` (:1173-1196). PINNED: tests/golden/tracker_*.npz
    were produced by the reference's own RuntimeTrackerBase class (oracle/make_golden.py).
  * `track_sequence_port` — the carried-track frame loop "O3" (SURVEY.md §8(c)). The reference ships
    this path broken (`head.is_first` is never cleared, head.py:106,115; forcing it crashes at :1235),
    so O3 is a SPECIFICATION composed only of reference functions plus three repairs:
      R1 honour `is_first` (ultralytics/nn/tasks.py:513 already passes it);
      R2 after decoding N = T + n_detect queries every Instances field has N rows:
         obj_idxes = cat(prev_ids, -1 x n_detect), disappear_time = cat(prev, 0 x n_detect)
         (upstream Instances.cat([active, init]), MOTR/models/motr.py:569-574);
      R3 next-frame tracks = rows with obj_idxes >= 0 (MOTR/models/qim.py:184-187) passed through
         `_update_track_embedding` (qim.py:251-301).
    Parity of O3 as a whole is therefore UNPINNED by any runnable reference (each of its parts is
    pinned separately: decoder, tracker update, QIM update).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import torch_port as tp

f32 = np.float32


def _iou(b1, b2):
    """head.py:1173-1196 in fp32 scalar arithmetic; boxes are read as (x, y, w, h)."""
    if abs(f32(b1[0] - b2[0])) > f32(f32(0.5) * min(b1[0], b2[0])):
        return f32(0)
    if abs(f32(b1[1] - b2[1])) > f32(f32(0.5) * min(b1[1], b2[1])):
        return f32(0)
    ix1, iy1 = max(b1[0], b2[0]), max(b1[1], b2[1])
    ix2 = min(f32(b1[0] + b1[2]), f32(b2[0] + b2[2]))
    iy2 = min(f32(b1[1] + b1[3]), f32(b2[1] + b2[3]))
    dx, dy = f32(ix2 - ix1), f32(iy2 - iy1)
    inter = f32((dx if dx > 0 else f32(0)) * (dy if dy > 0 else f32(0)))
    a1, a2 = f32(b1[2] * b1[3]), f32(b2[2] * b2[3])
    with np.errstate(divide="ignore", invalid="ignore"):
        return f32(inter / f32(f32(a1 + a2) - inter))


class TrackerPort:
    """RuntimeTrackerBase (head.py:1143-1283): thresholds :1146, counters max_obj_id / max_obj_id_pre."""

    def __init__(self, score_thresh=0.4, filter_score_thresh=0.5, miss_tolerance=5, iou_thresh=0.8):
        self.score_thresh, self.filter_score_thresh = f32(score_thresh), f32(filter_score_thresh)
        self.miss_tolerance, self.iou_thresh = miss_tolerance, f32(iou_thresh)
        self.max_obj_id = 0
        self.max_obj_id_pre = 0

    def clear(self):  # :1198-1199 (a fresh instance is created on is_first, head.py:200-202)
        self.max_obj_id = 0

    def filter_keep(self, boxes: np.ndarray) -> np.ndarray:
        """_filter_tracks, :1155-1171: greedy, index order, drop j when IoU(i, j) > thr."""
        n = boxes.shape[0]
        keep = np.ones(n, dtype=bool)
        for i in range(n):
            if keep[i]:
                for j in range(i + 1, n):
                    if keep[j] and _iou(boxes[i], boxes[j]) > self.iou_thresh:
                        keep[j] = False
        return keep

    def update(self, scores: np.ndarray, boxes: np.ndarray, obj_idxes: np.ndarray, disappear_time: np.ndarray):
        """In-place on obj_idxes / disappear_time (int64 [N]); returns the kept-active index list
        (what the reference returns and then discards, head.py:493-497)."""
        scores = scores.astype(np.float32)
        boxes = boxes.astype(np.float32)
        for i in range(len(scores)):                                          # :1232-1243
            if obj_idxes[i] == -1 and scores[i] >= self.score_thresh:
                obj_idxes[i] = self.max_obj_id
                self.max_obj_id += 1
            elif obj_idxes[i] >= 0 and scores[i] < self.filter_score_thresh:
                disappear_time[i] += 1
                if disappear_time[i] >= self.miss_tolerance:
                    obj_idxes[i] = -1
        active = np.nonzero(obj_idxes >= 0)[0]                                # :1245-1250
        if active.size == 0:
            return active
        keep = self.filter_keep(boxes[active])                                # :1256-1262
        kept = active[keep]
        ids = obj_idxes[kept].copy()                                          # the copy made by Instances.__getitem__
        tmp = 0
        for i in range(len(ids)):                                             # :1268-1275
            if ids[i] > self.max_obj_id_pre:
                ids[i] = self.max_obj_id_pre + tmp + 1
                tmp += 1
        self.max_obj_id = int(ids.max()) + 1                                  # :1280-1281
        self.max_obj_id_pre = self.max_obj_id - 1
        return kept


def track_sequence_port(sd: Dict[str, torch.Tensor], frames, shapes, n_heads: int, n_levels: int, n_points: int,
                        n_layers: int, nc: int, core=tp.msda_core_gridsample, record_embed: bool = False,
                        device="cpu", autocast_dtype=None):
    """O3 frame loop on the CPU (see module docstring). With device="cuda" the same PyTorch ops run eagerly on the
    GPU (bench.py's `gpu_eager_baseline`: what the reference's modules do after `.cuda()`, optionally under
    torch.autocast(autocast_dtype)); the ID assigner stays the host loop it is in the reference (head.py:1232-1243).

    frames: iterable of (feats [Lv, C], detect_embed [nd, C], detect_refer_logit [nd, 4]) fp32 CPU tensors.
    Query assembly follows head.py:1056-1064,1108-1109 (tracks first): embed = cat(class_embed[argmax
    prev logits], detect_embed); refer = cat(track ref_pts, detect refer); query_pos = cat(track
    query_pos, pos2posemb(detect refer)).
    Returns a list of per-frame dicts {ids, boxes, scores, labels, n_tracks_in, counters}.
    """
    import contextlib
    dev = torch.device(device)
    if dev.type != "cpu":
        sd = {k: v.to(dev) for k, v in sd.items()}
    C = sd["denoising_class_embed.weight"].shape[1]
    trk = TrackerPort()
    t_ref = torch.zeros(0, 4, device=dev)
    t_qpos = torch.zeros(0, C, device=dev)
    t_logits = torch.zeros(0, nc, device=dev)
    t_ids = np.zeros(0, dtype=np.int64)
    t_dis = np.zeros(0, dtype=np.int64)
    out = []
    amp = torch.autocast(dev.type, dtype=autocast_dtype) if autocast_dtype is not None else contextlib.nullcontext()
    with torch.no_grad(), amp:
        for feats, det_embed, det_refer in frames:
            if dev.type != "cpu":
                feats, det_embed, det_refer = feats.to(dev), det_embed.to(dev), det_refer.to(dev)
            T = t_ref.shape[0]
            nd = det_embed.shape[0]
            cls_embed = sd["denoising_class_embed.weight"][t_logits.argmax(-1)] if T else torch.zeros(0, C, device=dev)
            embed = torch.cat([cls_embed, det_embed], 0)[None]
            refer = torch.cat([t_ref, det_refer], 0)[None]
            qpos = torch.cat([t_qpos, tp.pos2posemb(det_refer)], 0)[None]
            boxes, logits, hs = tp.decoder_forward(sd, embed, refer, feats[None], shapes, n_heads, n_levels,
                                                   n_points, n_layers, "motr", qpos, core=core)
            boxes, logits, hs = boxes[0, 0].float(), logits[0, 0].float(), hs[0].float()
            scores = logits.sigmoid().max(-1).values                          # head.py:310
            ids = np.concatenate([t_ids, np.full(nd, -1, dtype=np.int64)])    # R2
            dis = np.concatenate([t_dis, np.zeros(nd, dtype=np.int64)])
            boxes_h, scores_h = boxes.cpu().numpy(), scores.cpu().numpy()     # (host copy, as head.py:1157 does)
            trk.update(scores_h, boxes_h, ids, dis)
            act_np = np.nonzero(ids >= 0)[0]
            act = torch.from_numpy(act_np).to(dev)                            # R3 (qim.py:184-187)
            rec = {"ids": ids.copy(), "boxes": boxes_h.copy(), "scores": scores_h.copy(),
                   "labels": logits.argmax(-1).cpu().numpy().copy(), "n_tracks_in": T,
                   "counters": (trk.max_obj_id, trk.max_obj_id_pre)}
            if record_embed:
                rec["hs"] = hs.cpu().numpy().copy()
            out.append(rec)
            new_qpos, new_ref = tp.qim_update(sd, refer[0][act], qpos[0][act], hs[act], boxes[act], n_heads)
            t_ref, t_qpos, t_logits = new_ref.float(), new_qpos.float(), logits[act]
            t_ids, t_dis = ids[act_np], dis[act_np]
    return out
