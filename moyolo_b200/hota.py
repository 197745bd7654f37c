"""HOTA of the device-resident track table.

The reference evaluates HOTA per video in its validator (ultralytics/models/MOTRtrack/val.py:257-310, 403-432) with
`ultralytics/utils/hota.py`, a locally edited copy of TrackEval's `HOTA` metric. The edited first pass mutates
the per-frame tracker-id arrays IN PLACE (`tracker_ids_t -= min(...)`, `tracker_ids_t -= 1`, hota.py:84-96) and
indexes the global alignment scores by position instead of by id (hota.py:118-119), so its second pass runs on shifted
ids; the validator wraps the call in try/except (val.py:288-327). What this module implements is therefore the
PUBLISHED algorithm that file was copied from (Luiten et al., IJCV 2021; TrackEval `trackeval/metrics/hota.py`): the
parts of the reference file that are unedited -- per-timestep matching and TP/FN/FP/LocA accumulation (hota.py:103-150),
association scores (:152-160), `_compute_final_fields` (:214-228), `combine_sequences` (:167-177), the alpha grid
(:16) and the field names (:17-21) -- are followed line by line, and `tests/test_hota.py` checks the two pure
functions against the reference's own on the CPU box.

Input = the track table `TrackEngine.track_table()` / `sharding.run_sharded()` produce, rows
[seq, frame, id, cx, cy, w, h, score, cls], and a ground-truth table [seq, frame, id, cx, cy, w, h]. The per-frame
IoU matrices (val.py:509-553, x0y0x1y1) are computed for ALL frames of a sequence in one batched tensor op on the
table's device; the assignment (scipy `linear_sum_assignment`, as in the reference) runs on the host.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

ALPHAS = np.arange(0.05, 0.99, 0.05)   # hota.py:16
INTEGER_ARRAY_FIELDS = ["HOTA_TP", "HOTA_FN", "HOTA_FP"]
FLOAT_ARRAY_FIELDS = ["HOTA", "DetA", "AssA", "DetRe", "DetPr", "AssRe", "AssPr", "LocA", "OWTA"]
FLOAT_FIELDS = ["HOTA(0)", "LocA(0)", "HOTALocA(0)"]
_EPS = float(np.finfo("float").eps)


def box_ious(gt_xyxy: torch.Tensor, trk_xyxy: torch.Tensor) -> torch.Tensor:
    """IoU of every (gt, tracker) box pair, boxes as (x0, y0, x1, y1); leading batch dims allowed
    (TrackValidator._calculate_box_ious, val.py:514-553: empty boxes have IoU 0)."""
    a, b = gt_xyxy[..., :, None, :], trk_xyxy[..., None, :, :]
    mn, mx = torch.minimum(a, b), torch.maximum(a, b)
    inter = (mn[..., 2] - mx[..., 0]).clamp(min=0) * (mn[..., 3] - mx[..., 1]).clamp(min=0)
    a1 = (gt_xyxy[..., 2] - gt_xyxy[..., 0]) * (gt_xyxy[..., 3] - gt_xyxy[..., 1])
    a2 = (trk_xyxy[..., 2] - trk_xyxy[..., 0]) * (trk_xyxy[..., 3] - trk_xyxy[..., 1])
    union = a1[..., :, None] + a2[..., None, :] - inter
    ok = (a1[..., :, None] > _EPS) & (a2[..., None, :] > _EPS) & (union > _EPS)
    return torch.where(ok, inter / union.clamp(min=_EPS), torch.zeros_like(inter))


def _cxcywh_to_xyxy(b: torch.Tensor) -> torch.Tensor:
    return torch.cat([b[..., :2] - b[..., 2:4] / 2, b[..., :2] + b[..., 2:4] / 2], -1)


def sequence_data(track_rows: torch.Tensor, gt_rows: torch.Tensor, n_frames: Optional[int] = None) -> Dict:
    """The `data` dict HOTA.eval_sequence takes (val.py:257-310) for ONE sequence: ids relabelled to 0..n-1, per
    timestep the gt / tracker id arrays and their IoU matrix. track_rows [N, >=7] = (seq, frame, id, cx, cy, w, h, ...),
    gt_rows [M, >=7] likewise (same device)."""
    dev = track_rows.device
    tf, gf = track_rows[:, 1].long(), gt_rows[:, 1].long()
    if n_frames is None:
        n_frames = int(max(int(tf.max()) if tf.numel() else -1, int(gf.max()) if gf.numel() else -1)) + 1
    t_ids_u, t_inv = torch.unique(track_rows[:, 2].long(), return_inverse=True)
    g_ids_u, g_inv = torch.unique(gt_rows[:, 2].long(), return_inverse=True)
    # pad every frame to the largest object count and compute all IoU matrices in one op on the table's device
    t_cnt = torch.bincount(tf, minlength=n_frames)
    g_cnt = torch.bincount(gf, minlength=n_frames)
    Tm, Gm = int(t_cnt.max()) if tf.numel() else 0, int(g_cnt.max()) if gf.numel() else 0

    def pad(rows, frames, cnt, width):
        order = torch.sort(frames, stable=True).indices
        start = torch.cumsum(cnt, 0) - cnt
        slot = torch.arange(rows.shape[0], device=dev) - start[frames[order]]
        boxes = torch.zeros(n_frames, max(width, 1), 4, dtype=torch.float64, device=dev)
        boxes[frames[order], slot] = _cxcywh_to_xyxy(rows[order, 3:7].double())
        return boxes, order, slot

    tb, t_order, _ = pad(track_rows, tf, t_cnt, Tm)
    gb, g_order, _ = pad(gt_rows, gf, g_cnt, Gm)
    ious = box_ious(gb, tb).cpu().numpy()                      # [F, Gm, Tm]
    t_lab, g_lab = t_inv[t_order].cpu().numpy(), g_inv[g_order].cpu().numpy()
    t_cnt_h, g_cnt_h = t_cnt.cpu().numpy(), g_cnt.cpu().numpy()
    t_off, g_off = np.concatenate([[0], np.cumsum(t_cnt_h)]), np.concatenate([[0], np.cumsum(g_cnt_h)])
    data = {"num_timesteps": n_frames, "num_gt_ids": int(g_ids_u.numel()), "num_tracker_ids": int(t_ids_u.numel()),
            "num_gt_dets": int(gt_rows.shape[0]), "num_tracker_dets": int(track_rows.shape[0]),
            "gt_ids": [], "tracker_ids": [], "similarity_scores": []}
    for t in range(n_frames):
        data["gt_ids"].append(g_lab[g_off[t]:g_off[t + 1]].astype(int))
        data["tracker_ids"].append(t_lab[t_off[t]:t_off[t + 1]].astype(int))
        data["similarity_scores"].append(ious[t, :g_cnt_h[t], :t_cnt_h[t]])
    return data


def compute_final_fields(res: Dict) -> Dict:
    """hota.py:214-228."""
    res["DetRe"] = res["HOTA_TP"] / np.maximum(1, res["HOTA_TP"] + res["HOTA_FN"])
    res["DetPr"] = res["HOTA_TP"] / np.maximum(1, res["HOTA_TP"] + res["HOTA_FP"])
    res["DetA"] = res["HOTA_TP"] / np.maximum(1, res["HOTA_TP"] + res["HOTA_FN"] + res["HOTA_FP"])
    res["HOTA"] = np.sqrt(res["DetA"] * res["AssA"])
    res["OWTA"] = np.sqrt(res["DetRe"] * res["AssA"])
    res["HOTA(0)"] = res["HOTA"][0]
    res["LocA(0)"] = res["LocA"][0]
    res["HOTALocA(0)"] = res["HOTA(0)"] * res["LocA(0)"]
    return res


def eval_sequence(data: Dict) -> Dict:
    """HOTA of one sequence from the `data` dict of sequence_data (gt / tracker ids 0-based and contiguous)."""
    from scipy.optimize import linear_sum_assignment
    nA = len(ALPHAS)
    res = {f: np.zeros(nA, dtype=np.float64) for f in FLOAT_ARRAY_FIELDS + INTEGER_ARRAY_FIELDS}
    for f in FLOAT_FIELDS:
        res[f] = 0
    if data["num_tracker_dets"] == 0:                                   # hota.py:36-40
        res["HOTA_FN"] = data["num_gt_dets"] * np.ones(nA)
        res["LocA"] = np.ones(nA)
        res["LocA(0)"] = 1.0
        return res
    if data["num_gt_dets"] == 0:                                        # hota.py:41-45
        res["HOTA_FP"] = data["num_tracker_dets"] * np.ones(nA)
        res["LocA"] = np.ones(nA)
        res["LocA(0)"] = 1.0
        return res
    G, T = data["num_gt_ids"], data["num_tracker_ids"]
    potential = np.zeros((G, T))
    gt_count, trk_count = np.zeros((G, 1)), np.zeros((1, T))
    for gt_ids, trk_ids, sim in zip(data["gt_ids"], data["tracker_ids"], data["similarity_scores"]):
        # global association statistics (hota.py:53-66; the id arrays are NOT modified here)
        denom = sim.sum(0)[np.newaxis, :] + sim.sum(1)[:, np.newaxis] - sim
        sim_iou = np.zeros_like(sim)
        mask = denom > 0 + _EPS
        sim_iou[mask] = sim[mask] / denom[mask]
        potential[gt_ids[:, np.newaxis], trk_ids[np.newaxis, :]] += sim_iou
        gt_count[gt_ids] += 1
        trk_count[0, trk_ids] += 1
    with np.errstate(divide="ignore", invalid="ignore"):
        global_alignment = potential / (gt_count + trk_count - potential)     # hota.py:99
    matches = [np.zeros_like(potential) for _ in ALPHAS]
    for gt_ids, trk_ids, sim in zip(data["gt_ids"], data["tracker_ids"], data["similarity_scores"]):
        if len(gt_ids) == 0:                                            # hota.py:105-108
            res["HOTA_FP"] += len(trk_ids)
            continue
        if len(trk_ids) == 0:                                           # hota.py:109-112
            res["HOTA_FN"] += len(gt_ids)
            continue
        score_mat = global_alignment[gt_ids[:, np.newaxis], trk_ids[np.newaxis, :]] * sim
        rows, cols = linear_sum_assignment(-score_mat)                  # hota.py:122-123
        for a, alpha in enumerate(ALPHAS):                              # hota.py:133-145
            ok = sim[rows, cols] >= alpha - _EPS
            ar, ac = rows[ok], cols[ok]
            n = len(ar)
            res["HOTA_TP"][a] += n
            res["HOTA_FN"][a] += len(gt_ids) - n
            res["HOTA_FP"][a] += len(trk_ids) - n
            if n > 0:
                res["LocA"][a] += float(sim[ar, ac].sum())
                matches[a][gt_ids[ar], trk_ids[ac]] += 1
    for a in range(nA):                                                 # hota.py:152-160
        mc = matches[a]
        ass_a = mc / np.maximum(1, gt_count + trk_count - mc)
        res["AssA"][a] = np.sum(mc * ass_a) / np.maximum(1, res["HOTA_TP"][a])
        res["AssRe"][a] = np.sum(mc * (mc / np.maximum(1, gt_count))) / np.maximum(1, res["HOTA_TP"][a])
        res["AssPr"][a] = np.sum(mc * (mc / np.maximum(1, trk_count))) / np.maximum(1, res["HOTA_TP"][a])
    res["LocA"] = np.maximum(1e-10, res["LocA"]) / np.maximum(1e-10, res["HOTA_TP"])     # hota.py:163
    return compute_final_fields(res)


def combine_sequences(all_res: Dict[object, Dict]) -> Dict:
    """hota.py:167-177 (with _BaseMetric._combine_sum / _combine_weighted_av)."""
    res = {f: sum(r[f] for r in all_res.values()) for f in INTEGER_ARRAY_FIELDS}
    for f in ("AssRe", "AssPr", "AssA"):
        res[f] = sum(r[f] * r["HOTA_TP"] for r in all_res.values()) / np.maximum(1.0, res["HOTA_TP"])
    loca = sum(r["LocA"] * r["HOTA_TP"] for r in all_res.values())
    res["LocA"] = np.maximum(1e-10, loca) / np.maximum(1e-10, res["HOTA_TP"])
    return compute_final_fields(res)


def hota_from_tables(track_table: torch.Tensor, gt_table: torch.Tensor, n_frames: Optional[Dict[int, int]] = None,
                     min_score: float = 0.0) -> Dict:
    """Per-sequence and combined HOTA of a track table [N, 9] = (seq, frame, id, cx, cy, w, h, score, cls) against a
    ground-truth table [M, >=7] = (seq, frame, id, cx, cy, w, h). Returns {"sequences": {seq: fields},
    "combined": fields, "summary": {field: mean over alpha}}."""
    if track_table.shape[0] and min_score > 0:
        track_table = track_table[track_table[:, 7] >= min_score]
    seqs = sorted(set(track_table[:, 0].long().unique().tolist()) | set(gt_table[:, 0].long().unique().tolist()))
    per: Dict[int, Dict] = {}
    for s in seqs:
        tr, gr = track_table[track_table[:, 0].long() == s], gt_table[gt_table[:, 0].long() == s]
        per[s] = eval_sequence(sequence_data(tr, gr, None if n_frames is None else n_frames.get(s)))
    combined = combine_sequences(per) if per else None
    summary = None if combined is None else {f: float(np.mean(combined[f])) for f in FLOAT_ARRAY_FIELDS}
    return {"sequences": per, "combined": combined, "summary": summary}
