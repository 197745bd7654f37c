"""Host logic of the sequence-sharding launcher on CPU: LPT assignment, lock-step grouping, and the
two-phase track-row gather across world_size 2 over gloo (the N>1 path; NCCL on the GPU box)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from moyolo_b200 import sharding


def test_lpt_assign_balanced_and_deterministic():
    counts = [500, 120, 480, 300, 90, 310, 450, 60]
    a = sharding.lpt_assign(counts, 4)
    assert sorted(i for r in a for i in r) == list(range(len(counts)))
    loads = [sum(counts[i] for i in r) for r in a]
    assert max(loads) - min(loads) <= max(counts) // 2
    assert a == sharding.lpt_assign(counts, 4)
    assert sharding.lpt_assign([], 2) == [[], []]
    assert sharding.lpt_assign([7], 3) == [[0], [], []]


def test_lockstep_groups():
    counts = [10, 50, 30, 40, 20]
    g = sharding.lockstep_groups([0, 1, 2, 3, 4], counts, 2)
    assert g == [[1, 3], [2, 4], [0]]


class _FakeEngine:
    """Stands in for TrackEngine on CPU: `tracks` every detect row whose first feature is positive."""

    def __init__(self, n_seq):
        self.n_seq = n_seq
        self.reset()

    def reset(self):
        self.rows, self.frame, self.seq_ids = [], 0, list(range(self.n_seq))

    def set_seq_ids(self, ids):
        self.seq_ids = list(ids)

    def submit(self, feats, det_embed, det_refer, want_rows=False, sync_inputs=False):
        for s in range(self.n_seq):
            keep = det_embed[s, :, 0] > 0
            n = int(keep.sum())
            r = torch.zeros(n, sharding.ROW_WIDTH)
            r[:, 0], r[:, 1] = float(self.seq_ids[s]), float(self.frame)
            r[:, 2] = torch.nonzero(keep)[:, 0].float()
            r[:, 3:7] = det_refer[s][keep]
            r[:, 7] = det_embed[s, keep, 0]
            self.rows.append(r)
        self.frame += 1
        return self.frame - 1

    def track_table(self):
        return torch.cat(self.rows, 0) if self.rows else torch.zeros(0, sharding.ROW_WIDTH)


def _sequences():
    seqs = []
    for i, n in enumerate([5, 3, 4, 2, 6]):
        g = torch.Generator().manual_seed(100 + i)
        fr = [(torch.randn(4, 8, generator=g), torch.randn(6, 8, generator=g), torch.rand(6, 4, generator=g))
              for _ in range(n)]
        seqs.append({"n_frames": n, "frames": (lambda t, fr=fr: fr[t])})
    return seqs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        table = sharding.run_sharded(_FakeEngine, _sequences(), rank, world, max_in_flight=2)
        torch.save(table, os.path.join(out_dir, f"t{rank}.pt"))
        # fixed-capacity single-collective variant (no count exchange): same table
        fixed = sharding.run_sharded(_FakeEngine, _sequences(), rank, world, max_in_flight=2, rows_per_frame=6)
        assert torch.equal(fixed, table)
        lazy = sharding.run_sharded(_FakeEngine, _sequences(), rank, world, max_in_flight=2, rows_per_frame=6, sort=False)
        assert lazy.count() == table.shape[0] and torch.equal(lazy.rows(sort=True), table)
        # a capacity that is too small is reported, not silently truncated
        small = sharding.gather_track_rows(torch.ones(5, sharding.ROW_WIDTH), capacity=3)
        try:
            small.count()
            raise AssertionError("overflow not detected")
        except RuntimeError as e:
            assert "capacity" in str(e)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_run_sharded_world2_matches_single_process(tmp_path):
    single = sharding.run_sharded(_FakeEngine, _sequences(), 0, 1, max_in_flight=2)
    assert single.shape[0] > 0
    # early-ending sequences in a lock-step group must not contribute surplus frames
    for i, s in enumerate(_sequences()):
        assert single[single[:, 0] == i][:, 1].max() == s["n_frames"] - 1
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    t0, t1 = torch.load(tmp_path / "t0.pt"), torch.load(tmp_path / "t1.pt")
    assert torch.equal(t0, t1), "every rank must hold the same gathered table"
    assert torch.equal(t0, single), "sharded table differs from the single-process table"


def test_gather_single_process_sorts_rows():
    rows = torch.tensor([[1, 2, 0, 0, 0, 0, 0, 0, 0], [0, 5, 1, 0, 0, 0, 0, 0, 0], [0, 1, 2, 0, 0, 0, 0, 0, 0]],
                        dtype=torch.float32)
    out = sharding.gather_track_rows(rows)
    assert out.count() == 3 and torch.equal(out.rows(), rows)          # single process: untouched, no sort
    assert out.rows(sort=True)[:, :2].tolist() == [[0, 1], [0, 5], [1, 2]]


def test_bind_host_to_gpu_partitions_the_cores():
    """sharding.bind_host_to_gpu: every local rank gets its own equal slice of the allowed cores (best effort; no GPU
    topology on this box, so only the per-rank split applies), and a single rank is left alone."""
    import os
    from moyolo_b200 import sharding
    if not hasattr(os, "sched_getaffinity"):
        pytest.skip("no sched_getaffinity")
    before = os.sched_getaffinity(0)
    try:
        assert sharding.bind_host_to_gpu(0, 1) is None and os.sched_getaffinity(0) == before
        if len(before) < 4:
            pytest.skip("needs >= 4 cores")
        seen = []
        for r in range(2):
            os.sched_setaffinity(0, before)
            what = sharding.bind_host_to_gpu(r, 2)
            assert what and "cpus" in what
            seen.append(os.sched_getaffinity(0))
        assert seen[0].isdisjoint(seen[1]) and len(seen[0]) == len(seen[1]) == len(before) // 2
        assert (seen[0] | seen[1]) <= before
    finally:
        os.sched_setaffinity(0, before)
