"""Fixed-Size Query Memory on the device: drop-in for `MOTR.models.fsqm.FSQM` (MOTR/models/fsqm.py:7-189).

Same constructor, attribute names and methods as the reference class (`query_memory`, `confidence`, `ids`,
`bounding_boxes`, `consecutive_low_frames`, `online_update`, `get_active_queries`, `reset`); the state lives on the
GPU and one `online_update` is ONE kernel launch (csrc/fsqm.cu) instead of three per-element Python loops with
`.item()` reads. Semantics: the specification F1-F3 of oracle/fsqm_port.py (identical to the shipped class wherever
that class is self-consistent; see tests/test_fsqm.py).

`TrackEngine(static_tracks=N)` is the engine-level counterpart: the fixed-size query memory the module exists for
(fsqm.py:8-10) so that every frame has the same query count and runs the same CUDA graph.
"""
from __future__ import annotations

import torch

from . import _lib


class FSQM:
    def __init__(self, max_num_queries: int, feature_dim: int, in_threshold: float = 0.7, out_threshold: float = 0.3,
                 consecutive_frames: int = 3, device="cuda"):
        self.max_num_queries, self.feature_dim = int(max_num_queries), int(feature_dim)
        self.in_threshold, self.out_threshold = float(in_threshold), float(out_threshold)
        self.consecutive_frames = int(consecutive_frames)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("Not implemented on the CPU")
        self.reset()

    def reset(self) -> None:                                             # fsqm.py:182-189
        N, d, dev = self.max_num_queries, self.feature_dim, self.device
        self.query_memory = torch.zeros(N, d, device=dev)
        self.confidence = torch.zeros(N, device=dev)
        self.ids = -torch.ones(N, dtype=torch.long, device=dev)
        self.bounding_boxes = torch.zeros(N, 4, device=dev)
        self.consecutive_low_frames = torch.zeros(N, dtype=torch.int32, device=dev)
        self._pool = torch.cat([torch.arange(N, device=dev), torch.zeros(N, dtype=torch.long, device=dev)])
        self._pool_hdr = torch.tensor([0, N], dtype=torch.int32, device=dev)

    @property
    def global_id_pool(self):
        """The FIFO id pool as a Python list (host read; for inspection only)."""
        head, count = self._pool_hdr.tolist()
        p = self._pool.tolist()
        return [p[(head + i) % len(p)] for i in range(count)]

    def online_update(self, detect_queries, track_queries, frame_idx: int = 0):
        """fsqm.py:155-180. `detect_queries`: an object with output_embedding [n, d], scores [n], pred_boxes [n, 4]
        (the reference's Instances or anything with those attributes); `track_queries`: obj_idxes [k] or [k, 1],
        scores [k], pred_boxes [k, 4]. Returns `track_queries` unchanged, as the reference does (:180)."""
        f32 = lambda t: t.detach().to(self.device, torch.float32).contiguous()  # noqa: E731
        emb, sc, box = f32(detect_queries.output_embedding), f32(detect_queries.scores), f32(detect_queries.pred_boxes)
        tid = track_queries.obj_idxes.detach().to(self.device, torch.long).reshape(-1).contiguous()
        tsc, tbox = f32(track_queries.scores), f32(track_queries.pred_boxes)
        if emb.shape[0] and emb.shape[1] != self.feature_dim:
            raise ValueError("detect_queries.output_embedding has the wrong feature dimension")
        p = lambda t: t.data_ptr() if t.numel() else None  # noqa: E731
        _lib.check(_lib.lib().moyolo_fsqm_update(
            self.max_num_queries, self.feature_dim, self.in_threshold, self.out_threshold, self.consecutive_frames,
            self.query_memory.data_ptr(), self.confidence.data_ptr(), self.ids.data_ptr(), self.bounding_boxes.data_ptr(),
            self.consecutive_low_frames.data_ptr(), self._pool.data_ptr(), self._pool_hdr.data_ptr(), tid.shape[0], p(tid),
            p(tsc), p(tbox), sc.shape[0], p(emb), p(sc), p(box), torch.cuda.current_stream(self.device).cuda_stream))
        return track_queries

    def get_active_queries(self):
        """fsqm.py:134-153: all slots (the reference returns the full memory, not only the live rows)."""
        return {"output_embedding": self.query_memory, "scores": self.confidence, "pred_boxes": self.bounding_boxes,
                "obj_idxes": self.ids.unsqueeze(1)}
