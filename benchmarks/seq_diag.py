"""Diagnostic: per-frame GPU-vs-oracle errors on a planted tracking workload (track rows vs detect rows)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from moyolo_b200 import synthetic as syn
from moyolo_b200.tracker import TrackEngine
from oracle.tracker_port import track_sequence_port

name, nf = sys.argv[1], int(sys.argv[2])
nd = 300 if not name.startswith("tiny") else 64
dev = torch.device("cuda:0")
spec, shapes, sd, plant = syn.tracking_workload(name, 7)
g = syn.PlantedSequenceGenerator(syn.SequenceSpec(name, nf, nd, 1, shapes=shapes), spec, plant)
frames = [tuple(t.clone() for t in g.next_frame()) for _ in range(nf)]
torch.set_num_threads(16)
recs = track_sequence_port(sd, frames, shapes, spec.n_heads, spec.n_levels, spec.n_points, spec.n_layers, spec.nc,
                           record_embed=True)
for precision in sys.argv[3:] or ["fp32", "bf16"]:
    eng = TrackEngine(sd, spec, shapes, dev, precision, nd, 1)
    for t in range(nf):
        o = eng.step(*[x[None].to(dev) for x in frames[t]])[0]
        r = recs[t]
        T = r["n_tracks_in"]
        b = o["boxes"].cpu().numpy()
        if b.shape != r["boxes"].shape:
            print(precision, t, "ROW COUNT", b.shape, r["boxes"].shape)
            break
        eb = np.abs(b - r["boxes"])
        es = np.abs(o["scores"].cpu().numpy() - r["scores"])
        ids_ok = np.array_equal(o["ids"].cpu().numpy(), r["ids"])
        lab_ok = np.array_equal(o["labels"].cpu().numpy(), r["labels"])
        hs = eng._last_plan.ws.x[:b.shape[0]].cpu().numpy()
        eh = np.abs(hs - r["hs"])
        print(f"{precision} t={t} T={T} ids_ok={ids_ok} lab_ok={lab_ok} box err track {eb[:T].max() if T else 0:.2e} "
              f"det {eb[T:].max():.2e} | score err track {es[:T].max() if T else 0:.2e} det {es[T:].max():.2e} | "
              f"hs err track {eh[:T].max() if T else 0:.2e} det {eh[T:].max():.2e} rms {np.sqrt((r['hs']**2).mean()):.2f}")
