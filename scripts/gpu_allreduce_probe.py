"""Stand-alone NCCL all-reduce of the step's gradient volume (39 MB fp32): what does the collective itself cost?"""
import os
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
for mb in (39.2, 26.0, 13.0, 4.0):
    n = int(mb * 1e6 / 4)
    x = torch.ones(n, device="cuda")
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20
    if rank == 0:
        print(f"world {world}: all_reduce {mb:5.1f} MB fp32: {t * 1e3:7.1f} us  algbw {mb / t:6.1f} GB/s  busbw {mb / t * 2 * (world - 1) / world:6.1f} GB/s", flush=True)
dist.destroy_process_group()
