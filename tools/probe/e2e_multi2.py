"""Same box, same ranks: raw copies of the bench's own pinned buffers vs the end-to-end calls."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.dup2(2, 1)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from crispr_hawk_b200 import _cabi, synth  # noqa: E402
from crispr_hawk_b200.workload import Workload  # noqa: E402

k = synth.CONFIGS["c2"]
c = synth.config_cohort("c2", 1.0, n_alt_hap=k["n_alt_hap"], hap_block=rank)
wl = Workload(c, k["pam"], k["guidelen"], k["right"], _cabi.Context.default(local), local)
hb = wl.host_buffers()
host = hb["ascii"]
dev = torch.empty_like(wl.ascii_dev)
d2h_src = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
d2h_dst = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
s2 = torch.cuda.Stream()


def together(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps


def both():
    dev.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2):
        d2h_dst.copy_(d2h_src, non_blocking=True)


out = {
    "h2d_5GB_ms": together(lambda: dev.copy_(host, non_blocking=True)),
    "d2h_1GiB_ms": together(lambda: d2h_dst.copy_(d2h_src, non_blocking=True)),
    "h2d_and_d2h_ms": together(both),
    "pack_only_ms": together(lambda: wl.batch.repack(wl.ascii_dev.data_ptr()) if wl.batch else wl.prepare_resident()),
    "step_resident_ms": together(lambda: wl.step_resident().close()),
    "two_call_ms": together(wl.step_host_twocall),
    "stream_ms": together(wl.step_host),
}
print(f"rank {rank} " + json.dumps({k2: round(v, 1) for k2, v in out.items()}), file=sys.stderr, flush=True)
dist.destroy_process_group()
