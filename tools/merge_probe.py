"""Where the time of the one-sided merge goes (2+ GPUs): allocation, IPC mapping, the pushes."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from crispr_hawk_b200 import _cabi, synth  # noqa: E402
from crispr_hawk_b200.workload import Workload  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = f"cuda:{local}"
k = synth.CONFIGS["c2"]
ctx = _cabi.Context.default(local)
wl = Workload(synth.config_cohort("c2", 1.0, hap_block=rank), k["pam"], k["guidelen"], k["right"], ctx, local)
res = wl.step_resident()
lib = ctx.lib
n, ts = res.n_guides, res.text_stride
ns = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
dist.all_gather(ns, torch.tensor([n], dtype=torch.int64, device=dev))
ns = [int(x.item()) for x in ns]
total, at = sum(ns), sum(ns[:rank])
off = (C.c_int64 * 7)()
nbytes = C.c_int64()
lib.hawk_merge_layout(total, ts, 1, off, C.byref(nbytes))


def tick(what, t0):
    torch.cuda.synchronize(dev)
    print(f"[rank {rank}] {what:28s} {1e3 * (time.perf_counter() - t0):8.2f} ms", file=sys.stderr)
    return time.perf_counter()


for rep in range(2):
    dist.barrier()
    t0 = time.perf_counter()
    base = C.c_void_p()
    handle = torch.zeros(64, dtype=torch.uint8)
    if rank == 0:
        hbuf = (C.c_uint8 * 64)()
        _cabi.check(lib.hawk_peer_alloc(ctx.handle, nbytes.value, C.byref(base), hbuf))
        handle = torch.tensor(list(hbuf), dtype=torch.uint8)
    t0 = tick("alloc + export", t0)
    hdev = handle.to(dev)
    dist.broadcast(hdev, src=0)
    handle = hdev.cpu()
    t0 = tick("broadcast handle", t0)
    if rank > 0:
        hbuf = (C.c_uint8 * 64)(*handle.tolist())
        _cabi.check(lib.hawk_peer_open(ctx.handle, hbuf, C.byref(base)))
    t0 = tick("ipc open", t0)
    pushed = C.c_int64()
    for with_text in (1, 0, 1):
        _cabi.check(lib.hawk_result_push(res.handle, 0, 0, -1, base, total, at, with_text, C.byref(pushed)))
        t0 = tick(f"push with_text={with_text} {pushed.value / 1e6:.0f} MB", t0)
    if rank > 0:
        _cabi.check(lib.hawk_peer_close(ctx.handle, base))
    t0 = tick("ipc close", t0)
    dist.barrier()
    if rank == 0:
        _cabi.check(lib.hawk_peer_free(ctx.handle, base))
    t0 = tick("free", t0)
dist.destroy_process_group()
