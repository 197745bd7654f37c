"""Stage timeline of the cluster decoder kernel (csrc/decoder_cluster.cu): %globaltimer stamps of cluster 0 / rank 0
for every layer, from a stand-alone launch on MOT17-shaped inputs. Prints mean microseconds per stage."""
import json
import sys
import torch
sys.path.insert(0, ".")
from moyolo_b200 import executor as ex, synthetic as syn
from moyolo_b200.tracker import DecoderWeights

name = sys.argv[1] if len(sys.argv) > 1 else "MOT17"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 352
dev = torch.device("cuda:0")
spec, shapes, sd, plant = syn.tracking_workload(name, 0)
W = DecoderWeights(sd, spec, dev, "bf16")
cd = ex.ClusterDecoder(W.layers, W.bbox, shapes, W.score_w, W.score_b)
Lv = syn.level_sizes(shapes)
g = torch.Generator().manual_seed(0)
x = torch.randn(R, 256, generator=g).to(dev)
pos = torch.randn(R, 256, generator=g).to(dev) * 0.5
ref = (torch.rand(R, 4, generator=g) * 0.5 + 0.2).to(dev)
values = torch.randn(1, Lv, 6 * 256, generator=g).to(dev, torch.bfloat16)
ro = torch.tensor([0, R], dtype=torch.int32, device=dev)
x_out = torch.empty_like(x)
kv = torch.zeros(2, R, 512, dtype=torch.bfloat16, device=dev)
bar = torch.zeros(1, dtype=torch.int32, device=dev)
refs = [torch.zeros(R, 4, device=dev) for _ in range(6)]
prof = torch.zeros(6, 16, dtype=torch.int64, device=dev)
m = cd.tile_rows(R, 1, R)
print("tile rows", m, "limits", cd.limits(32))
logits, scores, labels = torch.zeros(R, spec.nc, device=dev), torch.zeros(R, device=dev), torch.zeros(R, dtype=torch.int32, device=dev)
acc = None
N = 20
for it in range(N + 3):
    cd.run(x, pos, ref, values, ro, 1, R, m, x_out, kv, bar, refs, logits=logits, scores=scores, labels=labels, profile=prof)
    torch.cuda.synchronize()
    if it >= 3:
        p = prof.cpu().double()
        acc = p if acc is None else acc + p
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(N):
    cd.run(x, pos, ref, values, ro, 1, R, m, x_out, kv, bar, refs, logits=logits, scores=scores, labels=labels)
b.record()
torch.cuda.synchronize()
acc = acc / N
names = ["grid barrier wait", "K/V staging", "attention", "att exchange", "out_proj GEMM + send", "LN1 exchange + LayerNorm1",
         "offsets|logits GEMM", "gather", "gather exchange", "output_proj + exchange + LN2", "FFN1", "FFN2 + send",
         "reduce-scatter wait + sum + send", "LN3 exchange + LayerNorm3", "next qkv + box1 + arrive / outputs"]
d = (acc[:, 1:] - acc[:, :-1]) / 1e3
res = {n: round(float(d[:, i].mean()), 2) for i, n in enumerate(names)}
nxt = (acc[1:, 0] - acc[:-1, 15]) / 1e3
res["box head (to next layer start)"] = round(float(nxt.mean()), 2)
res["layer total us"] = round(float((acc[1:, 0] - acc[:-1, 0]).mean() / 1e3), 2)
res["kernel us (events, incl. launch + memset)"] = round(a.elapsed_time(b) * 1e3 / N, 2)
print(json.dumps(res, indent=1))
