"""torchrun --nproc-per-node N scripts/gpu_multicast_check.py: the in-switch all-reduce (vqa_multimem_all_reduce)
against NCCL on random gradients of the step's size, ragged tail included; then its timing."""
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_transfer_externaldata_b200 import lib as L  # noqa: E402

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
lib = L.load()
n = 9_790_272 + 64 * 3          # not a multiple of world * 4 * threads
cap = (n + 1023) // 1024 * 1024
buf = symm_mem.empty(cap, dtype=torch.float32, device=f"cuda:{local}")
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
assert hdl.has_multicast_support and hdl.multicast_ptr
g = torch.Generator(device="cuda").manual_seed(100 + rank)
x = torch.randn(n, device="cuda", generator=g)
ref = x.clone()
dist.all_reduce(ref)
buf[:n].copy_(x)
s = torch.cuda.current_stream()
hdl.barrier(channel=0)
L.check(lib.vqa_multimem_all_reduce(hdl.multicast_ptr, n, rank, world, 0, s.cuda_stream))
hdl.barrier(channel=1)
torch.cuda.synchronize()
err = (buf[:n] - ref).abs().max().item() / ref.abs().max().item()
# every rank must hold the SAME bits
mine = buf[:n].clone()
other = mine.clone()
dist.broadcast(other, src=0)
same = bool(torch.equal(mine, other))
print(f"rank {rank}: max rel err vs NCCL {err:.2e}; bit-identical to rank 0: {same}", flush=True)
assert err < 1e-6 and same

for ctas in (32, 64, 128):
    for _ in range(5):
        hdl.barrier(channel=0)
        L.check(lib.vqa_multimem_all_reduce(hdl.multicast_ptr, n, rank, world, ctas, s.cuda_stream))
        hdl.barrier(channel=1)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        hdl.barrier(channel=0)
        L.check(lib.vqa_multimem_all_reduce(hdl.multicast_ptr, n, rank, world, ctas, s.cuda_stream))
        hdl.barrier(channel=1)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"world {world}: multimem all-reduce {n * 4 / 1e6:.1f} MB, {ctas} CTAs: {us:.1f} us (incl. 2 barriers)", flush=True)
dist.destroy_process_group()
