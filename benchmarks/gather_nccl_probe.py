"""Device time of the final track-row gather over NCCL (sharding.gather_track_rows) and of the bare collective."""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from moyolo_b200 import sharding

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
cap = 5120
rows = torch.rand(800 + 10 * rank, 9, device=dev)
send = torch.zeros(cap + 1, 9, device=dev)
recv = torch.empty(world * (cap + 1), 9, device=dev)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0], ts[-1]


r1 = timed(lambda: dist.all_gather_into_tensor(recv, send))
r2 = timed(lambda: sharding.gather_track_rows(rows, capacity=cap))
r3 = timed(lambda: sharding.gather_track_rows(rows, capacity=1024))
small_s, small_r = torch.zeros(1025, 9, device=dev), torch.empty(world * 1025, 9, device=dev)
r4 = timed(lambda: dist.all_gather_into_tensor(small_r, small_s))
if rank == 0:
    print(f"world {world}: all_gather_into_tensor {send.numel() * 4 / 1e3:.0f} KB/rank: median {r1[0]:.1f} us (min {r1[1]:.1f}, max {r1[2]:.1f})")
    print(f"           all_gather_into_tensor 37 KB/rank: median {r4[0]:.1f} us")
    print(f"           gather_track_rows cap 5120: median {r2[0]:.1f} us (min {r2[1]:.1f}); cap 1024: median {r3[0]:.1f} us")
dist.destroy_process_group()
