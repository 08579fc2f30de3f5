"""multimem (NVLS, in-switch reduction) all-reduce through torch symmetric memory vs NCCL, on the step's 39 MB."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
group = dist.group.WORLD
n = int(39.2e6 / 4) // 1024 * 1024
try:
    t = symm_mem.empty(n, dtype=torch.float32, device=f"cuda:{local}")
    hdl = symm_mem.rendezvous(t, group)
    print(rank, "rendezvous ok; multicast:", getattr(hdl, "multicast_ptr", None) not in (None, 0), flush=True)
except Exception as e:  # noqa: BLE001
    print(rank, "symm_mem unavailable:", repr(e), flush=True)
    raise


def timeit(name, fn, reps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    if rank == 0:
        print(f"world {world}: {name:34s} {us:8.1f} us  busbw {n * 4 / us / 1e3 * 2 * (world - 1) / world:6.1f} GB/s", flush=True)


ref = torch.full((n,), float(rank + 1), device=f"cuda:{local}")
timeit("nccl all_reduce", lambda: dist.all_reduce(ref))
for name in ("multimem_all_reduce_", "two_shot_all_reduce_"):
    op = getattr(torch.ops.symm_mem, name)
    try:
        t.fill_(float(rank + 1))
        torch.cuda.synchronize(); dist.barrier()
        op(t, "sum", group.group_name)
        torch.cuda.synchronize()
        want = world * (world + 1) / 2
        ok = bool((t == want).all().item())
        if rank == 0:
            print(f"{name}: correct = {ok}", flush=True)
        timeit(name, lambda: op(t, "sum", group.group_name))
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print(f"{name}: failed: {e!r}", flush=True)
dist.destroy_process_group()
