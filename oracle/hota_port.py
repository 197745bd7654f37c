"""ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product package `moyolo_b200`.

Brute-force restatement of the published HOTA algorithm (Luiten et al., IJCV 2021 = TrackEval
`trackeval/metrics/hota.py`, the file ultralytics/utils/hota.py was copied from) in plain Python loops: no
vectorised indexing, per-pair dictionaries instead of count matrices, an exhaustive-search assignment for tiny frames
and scipy's Hungarian solver otherwise. Used to check moyolo_b200.hota on seeded tables.

Parity status: the reference's own `eval_sequence` is UNPINNABLE -- its edited first pass shifts the tracker ids in
place (hota.py:84-96) so its second pass counts matches under different ids than its first (demonstrated by
tests/test_hota.py::test_reference_eval_sequence_mutates_its_input). The unedited pieces (`_compute_final_fields`,
`combine_sequences`) are compared with the reference's directly in tests/test_hota.py.
"""
from __future__ import annotations

import itertools
from collections import defaultdict

import numpy as np

ALPHAS = [0.05 * (i + 1) for i in range(19)]
EPS = float(np.finfo("float").eps)


def _assign(score):
    """argmax over one-to-one assignments of sum(score): exhaustive for <= 6 x 6, Hungarian above."""
    n, m = len(score), len(score[0]) if len(score) else 0
    if n == 0 or m == 0:
        return []
    if max(n, m) <= 6:
        best, best_pairs = -1.0, []
        small, big, swap = (n, m, False) if n <= m else (m, n, True)
        for perm in itertools.permutations(range(big), small):
            pairs = [(perm[i], i) if swap else (i, perm[i]) for i in range(small)]
            tot = sum(score[r][c] for r, c in pairs)
            if tot > best + 1e-15:
                best, best_pairs = tot, pairs
        return best_pairs
    from scipy.optimize import linear_sum_assignment
    r, c = linear_sum_assignment(-np.asarray(score))
    return list(zip(r.tolist(), c.tolist()))


def hota_sequence(gt_ids, trk_ids, sims):
    """gt_ids[t], trk_ids[t]: lists of hashable ids; sims[t][i][j] = similarity of gt i and tracker j at time t.
    Returns dict field -> list over the 19 alphas (HOTA, DetA, AssA, DetRe, DetPr, AssRe, AssPr, LocA, TP, FN, FP)."""
    pot = defaultdict(float)
    gcount, tcount = defaultdict(int), defaultdict(int)
    for g, k, s in zip(gt_ids, trk_ids, sims):
        for i, gi in enumerate(g):
            gcount[gi] += 1
        for j, kj in enumerate(k):
            tcount[kj] += 1
        for i, gi in enumerate(g):
            for j, kj in enumerate(k):
                denom = sum(s[i][jj] for jj in range(len(k))) + sum(s[ii][j] for ii in range(len(g))) - s[i][j]
                if denom > EPS:
                    pot[(gi, kj)] += s[i][j] / denom
    out = {f: [0.0] * 19 for f in ("TP", "FN", "FP", "LocA", "AssA", "AssRe", "AssPr")}
    matches = [defaultdict(int) for _ in ALPHAS]
    for g, k, s in zip(gt_ids, trk_ids, sims):
        if len(g) == 0:
            for a in range(19):
                out["FP"][a] += len(k)
            continue
        if len(k) == 0:
            for a in range(19):
                out["FN"][a] += len(g)
            continue
        score = [[pot[(gi, kj)] / (gcount[gi] + tcount[kj] - pot[(gi, kj)]) * s[i][j] for j, kj in enumerate(k)]
                 for i, gi in enumerate(g)]
        pairs = _assign(score)
        for a, alpha in enumerate(ALPHAS):
            ok = [(i, j) for i, j in pairs if s[i][j] >= alpha - EPS]
            out["TP"][a] += len(ok)
            out["FN"][a] += len(g) - len(ok)
            out["FP"][a] += len(k) - len(ok)
            for i, j in ok:
                out["LocA"][a] += s[i][j]
                matches[a][(g[i], k[j])] += 1
    for a in range(19):
        tp = max(1.0, out["TP"][a])
        for (gi, kj), c in matches[a].items():
            out["AssA"][a] += c * c / max(1, gcount[gi] + tcount[kj] - c) / tp
            out["AssRe"][a] += c * c / max(1, gcount[gi]) / tp
            out["AssPr"][a] += c * c / max(1, tcount[kj]) / tp
        out["LocA"][a] = max(1e-10, out["LocA"][a]) / max(1e-10, out["TP"][a])
    res = dict(out)
    res["DetRe"] = [out["TP"][a] / max(1, out["TP"][a] + out["FN"][a]) for a in range(19)]
    res["DetPr"] = [out["TP"][a] / max(1, out["TP"][a] + out["FP"][a]) for a in range(19)]
    res["DetA"] = [out["TP"][a] / max(1, out["TP"][a] + out["FN"][a] + out["FP"][a]) for a in range(19)]
    res["HOTA"] = [(res["DetA"][a] * out["AssA"][a]) ** 0.5 for a in range(19)]
    res["OWTA"] = [(res["DetRe"][a] * out["AssA"][a]) ** 0.5 for a in range(19)]
    return res
