"""Under torchrun: ms per end-to-end call (config 2, host texts -> host table) per rank, ranks one
at a time vs all together, streamed (several group counts) vs two calls."""
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
wl.host_buffers()


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps


out = {}
for name, fn in (("two_call", wl.step_host_twocall), ("stream_g8", lambda: wl.step_host(n_groups=8)),
                 ("stream_g26", lambda: wl.step_host(n_groups=26)), ("stream_g104", lambda: wl.step_host(n_groups=104))):
    out[name + "_together"] = timed(fn)
    alone = None
    for r in range(world):
        dist.barrier()
        if r == rank:
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            alone = 1e3 * (time.perf_counter() - t0)
        dist.barrier()
    out[name + "_alone"] = alone
print(f"rank {rank} " + json.dumps({k2: round(v, 1) for k2, v in out.items()}), file=sys.stderr, flush=True)
dist.destroy_process_group()
