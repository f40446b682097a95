"""Group-count sweep of the streamed search on the bench workload (c2 by default):
ms per call for hawk_search_stream / hawk_search_stream_edits vs the two-call path."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from crispr_hawk_b200 import _cabi, synth  # noqa: E402
from crispr_hawk_b200.workload import Workload  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
k = synth.CONFIGS[name]
c = synth.config_cohort(name, 1.0, n_alt_hap=k["n_alt_hap"])
wl = Workload(c, k["pam"], k["guidelen"], k["right"])
wl.host_buffers()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps


out = {"two_call": timed(wl.step_host_twocall), "edits_two_call": timed(wl.step_edits_twocall)}
for g in (1, 4, 8, 13, 26, 52, 104, 0):
    out[f"stream_g{g}"] = timed(lambda: wl.step_host(n_groups=g))
for g in (1, 4, 8, 13, 26, 52, 0):
    out[f"edits_g{g}"] = timed(lambda: wl.step_edits(n_groups=g))
print(json.dumps({k: round(v, 2) for k, v in out.items()}))
