"""Track output formats: the text the reference writes after the per-frame hot path.

Input is the engine's track table / frame rows (`TrackEngine.track_table()`, rows
[seq, frame, id, cx, cy, w, h, score, cls] with boxes normalised (cx, cy, w, h)):

  * `mot_challenge_lines`  — MOTChallenge rows `frame,id,x1,y1,w,h,1,-1,-1,-1` exactly as
    `Detector.write_results` formats them (MOTR/submit_dance.py:410-419: boxes converted to pixel xyxy first,
    MOTR/submit_dance.py:300-330 / util/box_ops box_cxcywh_to_xyxy, ids < 0 skipped, values printed with
    Python's default float formatting of the float32 scalars);
  * `save_txt_lines`       — `TrackResults.save_txt` rows `track_id cls x y w h [conf]` with `%g` formatting
    (ultralytics/engine/results.py:475-512).

Host-side formatting only (numpy); nothing here launches device work.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Iterable, List

import numpy as np


def _table(rows) -> np.ndarray:
    t = rows.detach().cpu().numpy() if hasattr(rows, "detach") else np.asarray(rows)
    t = np.asarray(t, dtype=np.float32)
    if t.ndim != 2 or t.shape[1] != 9:
        raise ValueError("track table rows must be [N, 9] = (seq, frame, id, cx, cy, w, h, score, cls)")
    return t


def cxcywh_to_xyxy_pixels(boxes: np.ndarray, img_w: int, img_h: int) -> np.ndarray:
    """box_cxcywh_to_xyxy then scale by (w, h, w, h) in float32 (MOTR/util/box_ops.py:10-14,
    MOTR/models/motr.py TrackerPostProcess :336-352)."""
    b = np.asarray(boxes, dtype=np.float32)
    half = np.float32(0.5)
    xyxy = np.stack([b[:, 0] - half * b[:, 2], b[:, 1] - half * b[:, 3], b[:, 0] + half * b[:, 2],
                     b[:, 1] + half * b[:, 3]], axis=1).astype(np.float32)
    return xyxy * np.asarray([img_w, img_h, img_w, img_h], dtype=np.float32)


def mot_challenge_lines(rows, img_w: int, img_h: int, seq: int = None, first_frame: int = 1) -> List[str]:
    """MOTChallenge text rows of one sequence, frame numbers starting at `first_frame` (the reference numbers
    frames from 1, MOTR/submit_dance.py:372-393)."""
    t = _table(rows)
    if seq is not None:
        t = t[t[:, 0] == np.float32(seq)]
    xyxy = cxcywh_to_xyxy_pixels(t[:, 3:7], img_w, img_h)
    out = []
    for r, (x1, y1, x2, y2) in zip(t, xyxy):
        track_id = int(r[2])
        if track_id < 0:
            continue
        w, h = x2 - x1, y2 - y1  # float32 arithmetic, as the reference's numpy rows
        out.append(f"{int(r[1]) + first_frame},{track_id},{x1},{y1},{w},{h},1,-1,-1,-1\n")
    return out


def save_txt_lines(rows, img_w: int, img_h: int, save_conf: bool = False) -> Dict[int, List[str]]:
    """Per frame, the rows TrackResults.save_txt appends: `track_id cls x y w h [conf]`, `%g` formatted.
    The reference holds PIXEL xyxy boxes in its Results object and derives the normalised xywh from them
    (Boxes.xywhn = xyxy2xywh(xyxy) / (w, h, w, h), ultralytics/engine/results.py:590-604, utils/ops.py xyxy2xywh);
    the same float32 round trip is applied here so the printed digits agree. Returns {frame: [line, ...]}."""
    t = _table(rows)
    xyxy = cxcywh_to_xyxy_pixels(t[:, 3:7], img_w, img_h)
    two = np.float32(2)
    xywh = np.stack([(xyxy[:, 0] + xyxy[:, 2]) / two, (xyxy[:, 1] + xyxy[:, 3]) / two, xyxy[:, 2] - xyxy[:, 0],
                     xyxy[:, 3] - xyxy[:, 1]], axis=1).astype(np.float32)
    xywhn = xywh / np.asarray([img_w, img_h, img_w, img_h], dtype=np.float32)
    out: Dict[int, List[str]] = {}
    for r, b in zip(t, xywhn):
        line = (int(r[2]), int(r[8]), *[float(v) for v in b])
        if save_conf:
            line += (float(r[7]),)
        out.setdefault(int(r[1]), []).append(("%g " * len(line)).rstrip() % line + "\n")
    return out


def write_mot_challenge(path, rows, img_w: int, img_h: int, seq: int = None, first_frame: int = 1) -> int:
    lines = mot_challenge_lines(rows, img_w, img_h, seq, first_frame)
    Path(path).write_text("".join(lines))
    return len(lines)


def write_save_txt(directory, rows, img_w: int, img_h: int, stem: str = "frame", save_conf: bool = False) -> int:
    """One `<stem>_<frame>.txt` per frame under `directory` (the predictor's labels/ layout)."""
    d = Path(directory)
    d.mkdir(parents=True, exist_ok=True)
    frames = save_txt_lines(rows, img_w, img_h, save_conf)
    for f, lines in frames.items():
        (d / f"{stem}_{f}.txt").write_text("".join(lines))
    return len(frames)


def iter_sequences(rows) -> Iterable[int]:
    t = _table(rows)
    return sorted({int(s) for s in t[:, 0]})
