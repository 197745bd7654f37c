"""Host->device copy rate of one frame's inputs (pinned memory) when 1..N ranks copy at the same time: the ceiling
of the `e2e` leg's input stream on this box. Run under torchrun."""
import os
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = 13566 * 256 * 2 + 300 * 256 * 4 + 300 * 4 * 4      # MOT17 frame: feats bf16 + det_embed + det_refer
hosts = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(20)]
dst = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(4)]
for active in sorted({1, 2, 4, world} & set(range(1, world + 1))):
    for _ in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if rank < active:
            for i in range(200):
                dst[i % 4].copy_(hosts[i % 20], non_blocking=True)
        b.record()
        torch.cuda.synchronize()
    gbs = torch.tensor([nbytes * 200 / (a.elapsed_time(b) * 1e-3) / 1e9 if rank < active else 0.0], device=dev)
    if world > 1:
        out = torch.zeros(world, device=dev)
        dist.all_gather_into_tensor(out, gbs)
        gbs = out
    if rank == 0:
        v = [round(float(x), 1) for x in gbs.cpu().tolist()]
        print(f"{active} rank(s) copying 7.26 MB frames: GB/s per rank {v[:active]}  sum {sum(v):.1f}")
if world > 1:
    dist.destroy_process_group()
