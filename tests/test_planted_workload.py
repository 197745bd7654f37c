"""CPU checks of the planted-margin tracking workloads (moyolo_b200.synthetic.TRACKING_WORKLOADS) with the oracle:
every score keeps a wide margin to the 0.4 / 0.5 thresholds (head.py:1146), births follow the planted objectness
channel, non-persistent classes die after miss_tolerance frames, persistent classes are carried, and the label
of every born track is the planted class (head.py:888-900 then restarts it from that class embedding)."""
import numpy as np
import pytest
import torch

from moyolo_b200 import synthetic as syn
from oracle.tracker_port import track_sequence_port


@pytest.mark.parametrize("name", ["tiny", "tiny5"])
def test_planted_margins_births_deaths(name):
    spec, shapes, sd, plant = syn.tracking_workload(name, 7)
    nd, nf = 64, 14
    g = syn.PlantedSequenceGenerator(syn.SequenceSpec(name, nf, nd, 3, shapes=shapes), spec, plant)
    frames = [tuple(t.clone() for t in g.next_frame()) for _ in range(nf)]
    recs = track_sequence_port(sd, frames, shapes, spec.n_heads, spec.n_levels, spec.n_points, spec.n_layers, spec.nc)
    deaths = 0
    for t, r in enumerate(recs):
        s = r["scores"]
        assert np.minimum(np.abs(s - 0.4), np.abs(s - 0.5)).min() > 0.2, t
        T = r["n_tracks_in"]
        fire = (frames[t][1][:, 0] > 0).numpy()
        assert np.array_equal(r["ids"][T:] >= 0, fire), (t, "births follow the planted objectness")
        planted_cls = frames[t][1][:, 1:1 + spec.nc].argmax(-1).numpy()
        assert np.array_equal(r["labels"][T:], planted_cls), (t, "planted class")
        deaths += int((r["ids"][:T] < 0).sum())
    assert deaths > 0 and max(r["n_tracks_in"] for r in recs) >= 10
    if plant.persistent:   # a persistent class is never dropped
        last = recs[-1]
        T = last["n_tracks_in"]
        keep = np.isin(last["labels"][:T], list(plant.persistent))
        assert np.all(last["ids"][:T][keep] >= 0) and keep.sum() > 0


def test_state_dict_keys_unchanged():
    spec = syn.DecoderSpec(nc=5)
    a, b = syn.make_decoder_state(spec, 3), syn.make_tracking_state(spec, 3)
    assert a.keys() == b.keys() and all(a[k].shape == b[k].shape for k in a)
