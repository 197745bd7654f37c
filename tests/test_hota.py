"""HOTA of the track table (moyolo_b200.hota) against the brute-force oracle port and, where the reference tree is
present, against the reference's own unedited helper functions; plus the demonstration that the reference's edited
`eval_sequence` mutates its input ids (why it cannot serve as a golden)."""
import copy
import importlib
import sys
import types
from pathlib import Path

import numpy as np
import pytest
import torch

from moyolo_b200 import hota as H
from oracle import hota_port as P

REF_UTILS = Path("/root/reference/ultralytics/utils")


def _tables(seed, n_seq=2, n_frames=9, n_obj=5, miss=0.2, fp=0.15, switch=0.1):
    """Seeded ground truth (moving boxes) and a noisy tracker output with misses, false positives and id switches."""
    rng = np.random.default_rng(seed)
    gt, tr = [], []
    for s in range(n_seq):
        pos = rng.random((n_obj, 2)) * 0.6 + 0.2
        wh = rng.random((n_obj, 2)) * 0.1 + 0.05
        born = rng.integers(0, 3, n_obj)
        ids = np.arange(n_obj) + 10 * s
        tid = ids * 3 + 7
        nxt = 1000 + s
        for t in range(n_frames):
            pos = pos + rng.normal(0, 0.01, pos.shape)
            for o in range(n_obj):
                if t < born[o]:
                    continue
                gt.append([s, t, ids[o], pos[o, 0], pos[o, 1], wh[o, 0], wh[o, 1]])
                if rng.random() < miss:
                    continue
                if rng.random() < switch:
                    tid[o] = nxt
                    nxt += 1
                jit = rng.normal(0, 0.008, 4)
                tr.append([s, t, tid[o], pos[o, 0] + jit[0], pos[o, 1] + jit[1], wh[o, 0] * (1 + jit[2]),
                           wh[o, 1] * (1 + jit[3]), 0.9, 0])
            if rng.random() < fp:
                tr.append([s, t, nxt, rng.random(), rng.random(), 0.08, 0.08, 0.6, 0])
                nxt += 1
    return torch.tensor(tr, dtype=torch.float32), torch.tensor(gt, dtype=torch.float32)


def _oracle(track, gt, s):
    tr, g = track[track[:, 0] == s].numpy().astype(np.float64), gt[gt[:, 0] == s].numpy().astype(np.float64)
    n_frames = int(max(tr[:, 1].max(), g[:, 1].max())) + 1
    xyxy = lambda b: np.concatenate([b[:, :2] - b[:, 2:4] / 2, b[:, :2] + b[:, 2:4] / 2], 1)  # noqa: E731
    gi, ki, sims = [], [], []
    for t in range(n_frames):
        gt_t, tr_t = g[g[:, 1] == t], tr[tr[:, 1] == t]
        gi.append([int(x) for x in gt_t[:, 2]])
        ki.append([int(x) for x in tr_t[:, 2]])
        a, b = xyxy(gt_t[:, 3:7]), xyxy(tr_t[:, 3:7])
        sim = [[0.0] * len(b) for _ in range(len(a))]
        for i in range(len(a)):
            for j in range(len(b)):
                iw = max(min(a[i, 2], b[j, 2]) - max(a[i, 0], b[j, 0]), 0.0)
                ih = max(min(a[i, 3], b[j, 3]) - max(a[i, 1], b[j, 1]), 0.0)
                u = (a[i, 2] - a[i, 0]) * (a[i, 3] - a[i, 1]) + (b[j, 2] - b[j, 0]) * (b[j, 3] - b[j, 1]) - iw * ih
                sim[i][j] = iw * ih / u if u > 0 else 0.0
        sims.append(sim)
    return P.hota_sequence(gi, ki, sims)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_hota_matches_bruteforce_oracle(seed):
    track, gt = _tables(seed)
    out = H.hota_from_tables(track, gt)
    assert sorted(out["sequences"]) == [0, 1]
    for s, res in out["sequences"].items():
        ref = _oracle(track, gt, s)
        for f, rf in (("HOTA_TP", "TP"), ("HOTA_FN", "FN"), ("HOTA_FP", "FP")):
            assert np.array_equal(res[f], np.asarray(ref[rf])), (s, f)
        for f in ("HOTA", "DetA", "AssA", "DetRe", "DetPr", "AssRe", "AssPr", "LocA", "OWTA"):
            assert np.allclose(res[f], np.asarray(ref[f]), rtol=1e-9, atol=1e-12), (s, f)
        assert 0.2 < res["HOTA"][0] < 1.0           # the seeded tracker is neither perfect nor useless
    comb = out["combined"]
    assert np.array_equal(comb["HOTA_TP"], sum(r["HOTA_TP"] for r in out["sequences"].values()))
    assert set(out["summary"]) == set(H.FLOAT_ARRAY_FIELDS)


def test_hota_edge_cases():
    track, gt = _tables(5, n_seq=1)
    perfect = torch.cat([gt, torch.ones(gt.shape[0], 1), torch.zeros(gt.shape[0], 1)], 1)
    res = H.hota_from_tables(perfect, gt)["combined"]
    assert np.allclose(res["HOTA"], 1.0) and np.allclose(res["LocA"], 1.0) and res["HOTA_FP"].sum() == 0
    empty = H.eval_sequence(H.sequence_data(track[:0], gt))
    assert np.array_equal(empty["HOTA_FN"], gt.shape[0] * np.ones(19)) and empty["HOTA"].sum() == 0   # hota.py:36-40
    no_gt = H.eval_sequence(H.sequence_data(track, gt[:0]))
    assert np.array_equal(no_gt["HOTA_FP"], track.shape[0] * np.ones(19))                               # hota.py:41-45
    thr = H.hota_from_tables(track, gt, min_score=0.8)["combined"]
    assert thr["HOTA_FP"][0] <= H.hota_from_tables(track, gt)["combined"]["HOTA_FP"][0]


def _ref_hota():
    pkg = types.ModuleType("_ref_ul_utils")
    pkg.__path__ = [str(REF_UTILS)]
    sys.modules["_ref_ul_utils"] = pkg
    return importlib.import_module("_ref_ul_utils.hota").HOTA


@pytest.mark.skipif(not (REF_UTILS / "hota.py").exists(), reason="reference tree not present (GPU box)")
def test_final_fields_and_combination_equal_reference_functions():
    """The unedited parts of the reference metric: _compute_final_fields (hota.py:214-228), combine_sequences
    (:167-177), alpha grid and field names (:16-21)."""
    Ref = _ref_hota()
    ref = Ref()
    assert np.array_equal(ref.array_labels, H.ALPHAS)
    assert ref.float_array_fields == H.FLOAT_ARRAY_FIELDS and ref.integer_array_fields == H.INTEGER_ARRAY_FIELDS
    assert ref.float_fields == H.FLOAT_FIELDS
    track, gt = _tables(7, n_seq=3)
    per = H.hota_from_tables(track, gt)["sequences"]
    base = {s: {k: np.array(v, copy=True) if isinstance(v, np.ndarray) else v for k, v in r.items()} for s, r in per.items()}
    ours = H.combine_sequences(per)
    theirs = ref.combine_sequences(base)
    for f in H.FLOAT_ARRAY_FIELDS + H.INTEGER_ARRAY_FIELDS:
        assert np.allclose(ours[f], theirs[f], rtol=0, atol=1e-12), f
    for f in H.FLOAT_FIELDS:
        assert abs(ours[f] - theirs[f]) < 1e-12, f
    r0 = {k: np.array(v, copy=True) if isinstance(v, np.ndarray) else v for k, v in per[0].items()}
    again = Ref._compute_final_fields(dict(r0))
    for f in ("DetRe", "DetPr", "DetA", "HOTA", "OWTA"):
        assert np.array_equal(again[f], per[0][f]), f


@pytest.mark.skipif(not (REF_UTILS / "hota.py").exists(), reason="reference tree not present (GPU box)")
def test_reference_eval_sequence_mutates_its_input():
    """Why the reference's eval_sequence is not a golden: it shifts the tracker ids in place between its two passes
    (hota.py:84-96), so its match counts land on other ids than its alignment scores. On the same well-formed input
    the published algorithm (ours == brute-force oracle) differs from it."""
    Ref = _ref_hota()
    track, gt = _tables(1, n_seq=1)
    data = H.sequence_data(track, gt)
    ref_in = copy.deepcopy(data)
    ref_in["gt_ids"] = [g.reshape(-1, 1) for g in ref_in["gt_ids"]]       # the validator stores [n, 1] arrays
    before = [t.copy() for t in ref_in["tracker_ids"]]
    theirs = Ref().eval_sequence(ref_in)
    assert any(not np.array_equal(a, b) for a, b in zip(before, ref_in["tracker_ids"])), "ids were not mutated?"
    ours = H.eval_sequence(data)
    assert not np.allclose(ours["AssA"], theirs["AssA"])
    assert all(np.array_equal(a, b) for a, b in zip(data["tracker_ids"], H.sequence_data(track, gt)["tracker_ids"]))
