#!/usr/bin/env python
"""Developer aid (multi-GPU box, under torchrun): time ways of getting rank 0's baseband to every rank."""
import os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3 << 24
x = torch.zeros(n, dtype=torch.int32, device=dev)
if rank == 0:
    x.random_()
per = n // world
chunks = list(x.split(per))

def bcast():
    dist.broadcast(x, src=0)

def scatter_gather():
    dist.scatter(chunks[rank], chunks if rank == 0 else None, src=0)
    dist.all_gather_into_tensor(x, chunks[rank])

def timeit(fn, name):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-16s world=%d %.3f ms  %.1f GB/s of payload" % (name, world, t.item(), n * 4 / t.item() / 1e6), flush=True)

timeit(bcast, "broadcast")
timeit(scatter_gather, "scatter+allgather")
ref = x.clone(); dist.broadcast(ref, src=0)
scatter_gather()
assert torch.equal(ref, x)
dist.destroy_process_group()
