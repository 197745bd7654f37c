import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from moyolo_b200 import ops, synthetic as syn
from oracle import torch_port as tp
from conftest import rel_rms
dev = torch.device('cuda:0')
H, D, L, P, C = 8, 32, 3, 4, 256
shapes = [list(s) for s in syn.PYRAMIDS['MOT17']]
Lv = syn.level_sizes(shapes)
for seed in range(4):
    g = torch.Generator().manual_seed(seed)
    R = 382
    value32 = torch.randn(1, Lv, C, generator=g)
    value = value32.to(dev, torch.bfloat16)
    xq = torch.randn(R, C, generator=g).to(dev, torch.bfloat16)
    w = (torch.randn(H * L * P * 3, C, generator=g) * 0.05).to(dev, torch.bfloat16)
    bias = (torch.randn(H * L * P * 3, generator=g) * 0.5).to(dev)
    refer = torch.cat([torch.rand(R, 1, 2, generator=g), torch.rand(R, 1, 2, generator=g) * 0.4 + 0.02], -1).to(dev)
    fused = ops.msda_proj_fused(value, shapes, xq, w, bias, refer, H, P, 1)
    ol = ops.linear(xq, w, bias, out_dtype=torch.float32)
    n_off = H * L * P * 2
    two = ops.msda_fused(value, shapes, ol[:, :n_off], ol[:, n_off:], refer, H, P, 1)
    olr = xq.double().cpu() @ w.double().cpu().T + bias.double().cpu()
    off = olr[:, :n_off].view(R, H, L, P, 2)
    aw = torch.softmax(olr[:, n_off:].view(R, H, L * P), -1).view(R, H, L, P)
    rf = refer.double().cpu()
    loc = rf[:, :, None, None, :2] + off / P * rf[:, :, None, None, 2:] * 0.5
    ref = tp.msda_core_gather(value.double().cpu().view(1, Lv, H, D), shapes, loc[None], aw[None])[0]
    olf = ol.double().cpu()
    print(seed, 'fused', rel_rms(fused.double().cpu().numpy(), ref.numpy()), 'two', rel_rms(two.double().cpu().numpy(), ref.numpy()),
          'ol two-launch err', float((olf - olr).abs().max()))
