"""Sequence-sharding launcher.

Tracking is sequential within a video and independent across videos (the reference resets per
video, ultralytics/models/MOTRtrack/val.py:288-291, and upstream loops videos serially,
MOTR/submit_dance.py:499-504), so videos are assigned to ranks up front (longest-processing-time
first), each rank runs its videos in lock-step on its own GPU with NO collective inside the frame
loop, and one two-phase all_gather at the end collects the fixed-width track rows
    [seq, frame, id, cx, cy, w, h, score, cls]   (float32)
on every rank (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

ROW_WIDTH = 9


def lpt_assign(frame_counts: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of sequence indices to ranks (ties: lower index,
    lower rank), deterministic on every rank."""
    order = sorted(range(len(frame_counts)), key=lambda i: (-int(frame_counts[i]), i))
    loads = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += int(frame_counts[i])
    return out


def lockstep_groups(seq_ids: Sequence[int], frame_counts: Sequence[int], max_in_flight: int) -> List[List[int]]:
    """Split one rank's sequences into lock-step groups of at most `max_in_flight`, longest first, so
    sequences of similar length share a group (a group runs for max(len) frames)."""
    order = sorted(seq_ids, key=lambda i: (-int(frame_counts[i]), i))
    return [order[i:i + max_in_flight] for i in range(0, len(order), max_in_flight)]


def pack_track_rows(seq_idx: int, frame_idx: int, ids: torch.Tensor, boxes: torch.Tensor, scores: torch.Tensor,
                    labels: torch.Tensor) -> torch.Tensor:
    """All N rows of one frame as [N, 9] float32 (inactive rows keep id = -1; filtered at gather time
    so the frame loop stays free of host syncs)."""
    n = ids.shape[0]
    rows = torch.empty(n, ROW_WIDTH, dtype=torch.float32, device=ids.device)
    rows[:, 0] = float(seq_idx)
    rows[:, 1] = float(frame_idx)
    rows[:, 2] = ids.to(torch.float32)
    rows[:, 3:7] = boxes
    rows[:, 7] = scores
    rows[:, 8] = labels.to(torch.float32)
    return rows


def finalize_rows(frame_rows: List[torch.Tensor], device) -> torch.Tensor:
    """Concatenate per-frame rows and keep tracked objects only (id >= 0)."""
    if not frame_rows:
        return torch.zeros(0, ROW_WIDTH, dtype=torch.float32, device=device)
    allr = torch.cat(frame_rows, 0)
    return allr[allr[:, 2] >= 0].contiguous()


def gather_track_rows(local_rows: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Two-phase all_gather: row counts, then rows padded to the per-rank maximum. Every rank returns
    the full table sorted by (seq, frame, original order). Single process: returns the input."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return _sort_rows(local_rows)
    world = dist.get_world_size(group)
    dev = local_rows.device
    count = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    padded = torch.zeros(mx, ROW_WIDTH, dtype=torch.float32, device=dev)
    padded[:local_rows.shape[0]] = local_rows
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    return _sort_rows(torch.cat([b[:c] for b, c in zip(bufs, counts)], 0))


def _sort_rows(rows: torch.Tensor) -> torch.Tensor:
    if rows.shape[0] == 0:
        return rows
    key = rows[:, 0].double() * 1e7 + rows[:, 1].double()
    order = torch.sort(key, stable=True).indices
    return rows[order]


def run_sharded(make_engine, sequences: Sequence[Dict], rank: int, world_size: int, max_in_flight: int = 4,
                group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Track every sequence; return the gathered table on all ranks.

    sequences[i] = {"n_frames": int, "frames": callable t -> (feats, det_embed, det_refer)}.
    make_engine(n_seq) -> moyolo_b200.tracker.TrackEngine (or an object with the same reset /
    set_seq_ids / submit / track_table interface). The engine appends the tracked objects of every
    frame to a device-resident table itself, so the frame loop below only enqueues work.
    Sequences that end early keep stepping on their last frame (lock-step batch); their surplus rows
    are dropped when the table is read.
    """
    counts = [s["n_frames"] for s in sequences]
    mine = lpt_assign(counts, world_size)[rank]
    tables: List[torch.Tensor] = []
    device = None
    for grp in lockstep_groups(mine, counts, max_in_flight):
        eng = make_engine(len(grp))
        eng.reset()
        eng.set_seq_ids(grp)
        for t in range(max(counts[i] for i in grp)):
            batch = [sequences[i]["frames"](min(t, counts[i] - 1)) for i in grp]
            feats = torch.stack([b[0] for b in batch])
            de = torch.stack([b[1] for b in batch])
            dr = torch.stack([b[2] for b in batch])
            device = feats.device
            eng.submit(feats, de, dr, want_rows=False, sync_inputs=True)
        tab = eng.track_table()
        n_frames = torch.as_tensor(counts, dtype=torch.float32, device=tab.device)
        if tab.shape[0]:
            tab = tab[tab[:, 1] < n_frames[tab[:, 0].long()]]
        tables.append(tab.clone())
    local = torch.cat(tables, 0) if tables else torch.zeros(0, ROW_WIDTH, dtype=torch.float32,
                                                            device=device or torch.device("cpu"))
    return gather_track_rows(local, group)
