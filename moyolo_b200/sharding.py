"""Sequence-sharding launcher.

Tracking is sequential within a video and independent across videos (the reference resets per
video, ultralytics/models/MOTRtrack/val.py:288-291, and upstream loops videos serially,
MOTR/submit_dance.py:499-504), so videos are assigned to ranks up front (longest-processing-time
first), each rank runs its videos in lock-step on its own GPU with NO collective inside the frame
loop, and ONE all_gather_into_tensor of fixed-capacity buffers at the end collects the fixed-width track rows
    [seq, frame, id, cx, cy, w, h, score, cls]   (float32)
on every rank (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

ROW_WIDTH = 9


def bind_host_to_gpu(local_rank: int, local_world: int = 1) -> Optional[str]:
    """Pin this process (and the threads it starts later: NCCL proxy, copy helpers) to its own CPU cores: first the
    cores of the NUMA node its GPU hangs off (sysfs: the PCI device's numa_node), so pinned host buffers allocated
    afterwards are local to the GPU's PCIe root, then an equal slice of those cores per local rank, so the
    latency-critical enqueue threads of the ranks never share or migrate between cores. Best effort: returns a
    description, or None when nothing could be changed."""
    import os
    try:
        cpus = set(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return None
    what = []
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node >= 0:
            local = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                local.update(range(int(lo), int(hi or lo) + 1))
            n_nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
            if local & cpus and n_nodes > 1:
                cpus &= local
                what.append(f"numa node {node}")
    except Exception:   # no CUDA device / no sysfs topology: keep the whole affinity mask
        pass
    if local_world > 1:
        order = sorted(cpus)
        per = len(order) // local_world
        if per >= 2:
            cpus = set(order[(local_rank % local_world) * per:(local_rank % local_world + 1) * per])
            what.append(f"cpus {min(cpus)}-{max(cpus)}")
    if not what:
        return None
    try:
        os.sched_setaffinity(0, cpus)
    except OSError:
        return None
    return ", ".join(what)


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


class GatheredTable:
    """Result of `gather_track_rows`: the track rows of every rank, merged on the device by rank offset. Nothing
    here synchronises the host until `rows()` / `count()` is called."""

    def __init__(self, merged: torch.Tensor, total: Optional[torch.Tensor], n_host: Optional[int] = None,
                 overflow: Optional[torch.Tensor] = None):
        self._merged, self._total, self._n, self._overflow = merged, total, n_host, overflow

    def count(self) -> int:
        if self._n is None:
            if self._overflow is not None:
                info = torch.stack([self._total.reshape(()), self._overflow.reshape(())]).cpu()   # ONE host sync
                if int(info[1]) != 0:
                    raise RuntimeError("gather_track_rows: a rank produced more rows than the fixed gather capacity")
                self._n = int(info[0])
            else:
                self._n = int(self._total.cpu())
        return self._n

    def rows(self, sort: bool = False) -> torch.Tensor:
        """[n, 9] rows ordered by rank, then as each rank emitted them (frame, lock-step slot, query order). Ranks
        own disjoint sequences, so this is already grouped per sequence and frame-ordered within one;
        sort=True gives the canonical (seq, frame, original order) ordering, identical for every world size."""
        r = self._merged[:self.count()]
        return _sort_rows(r) if sort else r


def gather_track_rows(local_rows: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                      capacity: Optional[int] = None, count_dev: Optional[torch.Tensor] = None,
                      overflow_dev: Optional[torch.Tensor] = None) -> GatheredTable:
    """All ranks' track rows on every rank with ONE collective and no host synchronisation: each rank contributes
    a fixed-capacity buffer [capacity + 1, 9] whose row 0 is a header (row count, overflow flag),
    `all_gather_into_tensor` moves it, and the ranks' segments are merged by offset on the device (exclusive
    prefix sum of the header counts + one index_copy; fixed shapes, nothing is read back).
    capacity: rows per rank, known to every rank without communication (e.g. frames x sequences x a bound on
    tracked objects per frame). None: a count exchange picks the exact capacity first (one extra small collective
    and one host read). Single process: returns the input.
    count_dev (CUDA, needs `capacity`): device int32 holding this rank's row count -- `local_rows` is then the whole
    table buffer and the host never learns the count before the collective (TrackEngine.track_table_device);
    overflow_dev: device int32, non-zero = this rank's table is incomplete."""
    n_local = int(local_rows.shape[0]) if count_dev is None else 0
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if count_dev is not None:
            return GatheredTable(local_rows, count_dev.reshape(()), None, None if overflow_dev is None else overflow_dev.reshape(()))
        return GatheredTable(local_rows, None, n_local)
    world = dist.get_world_size(group)
    dev = local_rows.device
    if count_dev is not None and (capacity is None or dev.type != "cuda"):
        raise ValueError("gather_track_rows: count_dev needs a CUDA table and a fixed capacity")
    if capacity is None:
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, torch.tensor([n_local], dtype=torch.int64, device=dev), group=group)
        capacity = max(int(counts.max().cpu()), 1)
    cap = int(capacity)
    if dev.type == "cuda":
        # device path: pack kernel -> ONE all_gather_into_tensor -> merge kernel (moyolo_table_pack / _merge): the
        # eager torch version below costs ~250 us of launch gaps on the GPU, this one ~60 us at 8 ranks
        from . import _lib
        bufs = _gather_buffers(dev, world, cap)
        send, recv, merged, info = bufs
        st = torch.cuda.current_stream(dev).cuda_stream
        rows = local_rows if local_rows.is_contiguous() else local_rows.contiguous()
        _lib.check(_lib.lib().moyolo_table_pack(
            rows.data_ptr() if (n_local or count_dev is not None) else None, n_local,
            None if count_dev is None else count_dev.data_ptr(), None if overflow_dev is None else overflow_dev.data_ptr(),
            cap, send.data_ptr(), st))
        dist.all_gather_into_tensor(recv, send, group=group)
        _lib.check(_lib.lib().moyolo_table_merge(recv.data_ptr(), world, cap, merged.data_ptr(), info.data_ptr(), st))
        return GatheredTable(merged, info[0], None, info[1])
    send = torch.zeros(cap + 1, ROW_WIDTH, dtype=torch.float32, device=dev)     # host logic (gloo tests)
    n_send = min(n_local, cap)
    send[0, 0] = float(n_send)
    send[0, 1] = 1.0 if n_local > cap else 0.0
    send[1:1 + n_send] = local_rows[:n_send]
    recv = torch.empty(world * (cap + 1), ROW_WIDTH, dtype=torch.float32, device=dev)   # concatenated along dim 0
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, cap + 1, ROW_WIDTH)
    counts = recv[:, 0, 0].round().long()                                  # [world], on the device
    offsets = torch.cumsum(counts, 0) - counts                             # exclusive prefix sum
    i = torch.arange(cap, device=dev)
    dst = torch.where(i[None, :] < counts[:, None], offsets[:, None] + i[None, :], world * cap)   # invalid -> dump row
    merged = torch.zeros(world * cap + 1, ROW_WIDTH, dtype=torch.float32, device=dev)
    merged.index_copy_(0, dst.reshape(-1), recv[:, 1:].reshape(world * cap, ROW_WIDTH))
    return GatheredTable(merged, counts.sum(), None, recv[:, 0, 1].sum())


_GATHER_BUFFERS = {}


def _gather_buffers(dev, world: int, cap: int):
    """Persistent send / receive / merged / info buffers per (device, world, capacity): the gather allocates nothing."""
    key = (str(dev), world, cap)
    b = _GATHER_BUFFERS.get(key)
    if b is None:
        b = (torch.zeros(cap + 1, ROW_WIDTH, dtype=torch.float32, device=dev),
             torch.empty(world * (cap + 1), ROW_WIDTH, dtype=torch.float32, device=dev),
             torch.zeros(world * cap, ROW_WIDTH, dtype=torch.float32, device=dev),
             torch.zeros(2, dtype=torch.int32, device=dev))
        _GATHER_BUFFERS.clear()   # one configuration at a time
        _GATHER_BUFFERS[key] = b
    return b


def _sort_rows(rows: torch.Tensor) -> torch.Tensor:
    if rows.shape[0] == 0:
        return rows
    key = rows[:, 0].double() * 1e7 + rows[:, 1].double()
    order = torch.sort(key, stable=True).indices
    return rows[order]


def _stack_frames(sequences, counts, grp, t):
    batch = [sequences[i]["frames"](min(t, counts[i] - 1)) for i in grp]
    return tuple(torch.stack([b[k] for b in batch]) for k in range(3))


def run_sharded(make_engine, sequences: Sequence[Dict], rank: int, world_size: int, max_in_flight: int = 4,
                group: Optional[dist.ProcessGroup] = None, batch_fn=None, rows_per_frame: Optional[int] = None,
                sort: bool = True, sync_inputs: bool = True, before_gather=None):
    """Track every sequence; return the gathered table on all ranks (a tensor in canonical (seq, frame) order, or
    with sort=False the lazy `GatheredTable`, which costs no host synchronisation).

    batch_fn(group_seq_ids, t) -> (feats, det_embed, det_refer) already stacked for the lock-step group (avoids
    the per-frame torch.stack); rows_per_frame: bound on tracked objects per frame and sequence -> fixed gather
    capacity, one collective, no count exchange. sync_inputs=False: the frame tensors are complete when they are
    handed over (resident inputs), so their copy need not wait for the caller's stream. before_gather(get_local)
    is called once this rank's table is complete on the stream (bench.py puts its timing mark there); get_local()
    returns this rank's rows (it may synchronise: call it after the run).

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
    groups = lockstep_groups(mine, counts, max_in_flight)
    capacity = None
    if rows_per_frame is not None:   # the same number on every rank: the most loaded rank's frames x the bound
        assign = lpt_assign(counts, world_size)
        capacity = max(1, max(sum(counts[i] for i in a) for a in assign) * int(rows_per_frame))
    # fully asynchronous tail: one lock-step group of equally long sequences on a CUDA engine with a fixed gather
    # capacity -> the gather is enqueued behind the last frame without the host ever reading the row count
    if (world_size > 1 and capacity is not None and len(groups) == 1 and len({counts[i] for i in groups[0]}) == 1):
        grp = groups[0]
        eng = make_engine(len(grp))
        if hasattr(eng, "track_table_device"):
            eng.reset()
            eng.set_seq_ids(grp)
            for t in range(counts[grp[0]]):
                feats, de, dr = batch_fn(grp, t) if batch_fn is not None else _stack_frames(sequences, counts, grp, t)
                eng.submit(feats, de, dr, want_rows=False, sync_inputs=sync_inputs)
            while True:
                aborts = eng.aborts
                buf, n_dev, over_dev = eng.track_table_device()
                if before_gather is not None:
                    before_gather(lambda: eng.track_table())
                table = gather_track_rows(buf, group, capacity, n_dev, over_dev)
                eng.drain()
                if eng.aborts == aborts:     # no speculative frame was re-run after the gather was enqueued
                    break
            return table.rows(sort=True) if sort else table
    for grp in groups:
        eng = make_engine(len(grp))
        eng.reset()
        eng.set_seq_ids(grp)
        for t in range(max(counts[i] for i in grp)):
            feats, de, dr = batch_fn(grp, t) if batch_fn is not None else _stack_frames(sequences, counts, grp, t)
            device = feats.device
            eng.submit(feats, de, dr, want_rows=False, sync_inputs=sync_inputs)
        tab = eng.track_table()
        g_max = max(counts[i] for i in grp)
        if tab.shape[0] and any(counts[i] != g_max for i in grp):   # drop the surplus frames of early-ending sequences
            n_frames = torch.as_tensor(counts, dtype=torch.float32, device=tab.device)
            tab = tab[tab[:, 1] < n_frames[tab[:, 0].long()]]
        tables.append(tab if len(groups) == 1 else tab.clone())   # (make_engine may hand out one engine repeatedly)
    if len(tables) == 1:
        local = tables[0]     # (a view of the engine's table: consumed by the gather before the engine runs again)
    else:
        local = torch.cat(tables, 0) if tables else torch.zeros(0, ROW_WIDTH, dtype=torch.float32,
                                                                device=device or torch.device("cpu"))
    if before_gather is not None:
        before_gather(lambda: local)
    table = gather_track_rows(local, group, capacity)
    return table.rows(sort=True) if sort else table
