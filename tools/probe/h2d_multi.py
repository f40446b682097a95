"""Per-rank and concurrent H2D bandwidth under torchrun, with and without binding each rank to
its GPU's CPUs before the pinned buffer is allocated (is the aggregate host->GPU rate a NUMA
placement problem or a platform limit?)."""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.dup2(2, 1)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
bind = len(sys.argv) > 1 and sys.argv[1] == "bind"
aff0 = len(os.sched_getaffinity(0))
if bind:
    import pynvml

    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
aff1 = sorted(os.sched_getaffinity(0))
n = 2 << 30
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
host.fill_(1)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")


def h2d(reps=3):
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
        alone = h2d()
    dist.barrier()
dist.barrier()
together = h2d(4)
dist.barrier()
print(f"rank {rank} bind={bind} cpus {aff0}->{len(aff1)} [{aff1[0]}..{aff1[-1]}] alone {alone:.1f} GB/s together {together:.1f} GB/s",
      file=sys.stderr, flush=True)
dist.destroy_process_group()
