"""ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product package `moyolo_b200`.

Fixed-Size Query Memory (MOTR/models/fsqm.py:7-189), restated twice in plain Python / numpy:

  * `FsqmReference` -- the class EXACTLY as shipped, quirks included. PINNED: tests/golden/fsqm_clean.npz and
    fsqm_quirks.npz hold the complete state of the reference class itself after every online_update on seeded frames
    (oracle/make_golden.py gen_fsqm); tests/test_fsqm.py checks this restatement against both.
  * `FsqmSpec` -- the SPECIFICATION the product implements (moyolo_b200.fsqm, csrc/fsqm.cu): the same state machine
    with three repairs, each where the shipped code contradicts its own docstring:
      F1  a track query updates the slot that HOLDS its id (the reference uses the id as the slot index,
          fsqm.py:127-132, which is the same thing only until the first id is recycled into another slot);
      F2  only live slots age (the reference also ages empty slots and appends their id -1 to the id pool every
          consecutive_frames frames, fsqm.py:104-115, so the pool fills up with -1);
      F3  a slot is freed after `consecutive_frames` CONSECUTIVE frames below out_threshold (the docstring,
          fsqm.py:24): the low-frame counter is cleared when the confidence is at or above the threshold, not by
          every update (fsqm.py:131 clears it even for a low score, so a tracked slot can never be freed).
    On fsqm_clean.npz (the domain where the shipped class is self-consistent) Spec == Reference state for state;
    parity of the spec beyond that domain is UNPINNED by construction and documented as such in DESIGN.md.
"""
from __future__ import annotations

import numpy as np


class FsqmReference:
    """fsqm.py:12-44 (state), :46-100 (inject), :102-115 (remove), :117-132 (update_confidence), :155-180."""

    def __init__(self, max_num_queries, feature_dim, in_threshold=0.7, out_threshold=0.3, consecutive_frames=3):
        self.N, self.d = max_num_queries, feature_dim
        self.in_threshold, self.out_threshold, self.consecutive_frames = in_threshold, out_threshold, consecutive_frames
        self.reset()

    def reset(self):                                                     # :182-189
        self.query_memory = np.zeros((self.N, self.d), np.float32)
        self.confidence = np.zeros(self.N, np.float32)
        self.ids = -np.ones(self.N, np.int64)
        self.bounding_boxes = np.zeros((self.N, 4), np.float32)
        self.consecutive_low_frames = np.zeros(self.N, np.int32)
        self.global_id_pool = list(range(self.N))

    def update_confidence(self, tid, tsc, tbox):                         # :117-132
        for i in range(len(tid)):
            q = int(tid[i])
            if 0 <= q < self.N:
                self.confidence[q] = tsc[i]
                self.bounding_boxes[q] = tbox[i]
                self.consecutive_low_frames[q] = 0

    def inject_new_queries(self, emb, sc, box):                          # :46-100
        valid = np.nonzero(sc > np.float32(self.in_threshold))[0]
        for i in valid:
            free = np.nonzero(self.ids == -1)[0]
            if len(free) == 0:
                break                                                    # memory full: not injected (:78-81)
            s = free[0]
            self.ids[s] = self.global_id_pool.pop(0)
            self.query_memory[s] = emb[i]
            self.confidence[s] = sc[i]
            self.bounding_boxes[s] = box[i]
            self.consecutive_low_frames[s] = 0

    def remove_inactive_queries(self):                                   # :102-115
        for s in np.nonzero(self.confidence < np.float32(self.out_threshold))[0]:
            self.consecutive_low_frames[s] += 1
            if self.consecutive_low_frames[s] >= self.consecutive_frames:
                self.query_memory[s] = 0
                self.confidence[s] = 0.0
                self.global_id_pool.append(int(self.ids[s]))
                self.ids[s] = -1
                self.bounding_boxes[s] = 0
                self.consecutive_low_frames[s] = 0

    def online_update(self, emb, sc, box, tid, tsc, tbox):               # :155-180
        self.update_confidence(tid, tsc, tbox)
        self.inject_new_queries(emb, sc, box)
        self.remove_inactive_queries()


class FsqmSpec(FsqmReference):
    """Repairs F1-F3 (module docstring); everything else as the reference."""

    def update_confidence(self, tid, tsc, tbox):
        for i in range(len(tid)):
            q = int(tid[i])
            if q < 0:
                continue
            hit = np.nonzero(self.ids == q)[0]                           # F1: the slot holding this id
            if len(hit):
                self.confidence[hit[0]] = tsc[i]
                self.bounding_boxes[hit[0]] = tbox[i]                    # (F3: the low-frame counter is not touched)

    def remove_inactive_queries(self):
        for s in range(self.N):
            if self.ids[s] == -1:                                        # F2: empty slots do not age
                continue
            if self.confidence[s] < np.float32(self.out_threshold):
                self.consecutive_low_frames[s] += 1
                if self.consecutive_low_frames[s] >= self.consecutive_frames:
                    self.query_memory[s] = 0
                    self.confidence[s] = 0.0
                    self.global_id_pool.append(int(self.ids[s]))
                    self.ids[s] = -1
                    self.bounding_boxes[s] = 0
                    self.consecutive_low_frames[s] = 0
            else:
                self.consecutive_low_frames[s] = 0                       # F3
