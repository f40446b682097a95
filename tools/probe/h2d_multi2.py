"""Concurrent H2D of LARGE pinned buffers under torchrun: torch's pinned allocator vs
transparent-huge-page backed memory registered with cudaHostRegister."""
import mmap
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.dup2(2, 1)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 5 << 30
dev = torch.empty(n, dtype=torch.uint8, device="cuda")


def run(host, label):
    def h2d(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            dev.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        return reps * n / (time.perf_counter() - t0) / 1e9

    h2d(1)
    alone = None
    for r in range(world):
        dist.barrier()
        if r == rank:
            alone = h2d(2)
        dist.barrier()
    dist.barrier()
    together = h2d(3)
    dist.barrier()
    print(f"rank {rank} {label}: pinned={host.is_pinned()} alone {alone:.1f} GB/s together {together:.1f} GB/s", file=sys.stderr, flush=True)


a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
a.fill_(1)
run(a, "torch pin_memory 5 GiB")
del a
torch._C._host_emptyCache() if hasattr(torch._C, "_host_emptyCache") else None
m = mmap.mmap(-1, n, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
try:
    m.madvise(mmap.MADV_HUGEPAGE)
except Exception as e:
    print("madvise failed", e, file=sys.stderr)
arr = np.frombuffer(m, np.uint8)
arr[:] = 1
rc = torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, n, 0)
b = torch.from_numpy(arr)
run(b, f"THP + cudaHostRegister (rc={rc}) 5 GiB")
if rank == 0:
    print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), file=sys.stderr)
    print([l for l in open("/proc/meminfo") if "Huge" in l], file=sys.stderr)
dist.destroy_process_group()
