import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from crispr_hawk_b200 import synth, _cabi
from crispr_hawk_b200.workload import Workload
k = synth.CONFIGS["c2"]
c = synth.config_cohort("c2")
wl = Workload(c, k["pam"], k["guidelen"], k["right"])
wl.step_edits()
e = wl._edits
for i in range(3):
    t0 = time.perf_counter()
    batch = _cabi.Batch.from_edits(wl.ctx, e["ref"], c.region_start, e["off"], e["pos"], e["rl"], e["al"], e["ao"], e["pool"])
    t1 = time.perf_counter()
    res = _cabi.search(wl.ctx, batch, wl.params, wl.a, wl.b, wl.d.is_ref)
    t2 = time.perf_counter()
    table = res.table(wl._edits_out)
    t3 = time.perf_counter()
    res.close(); batch.close()
    print(f"from_edits {1e3*(t1-t0):.2f} ms  search {1e3*(t2-t1):.2f} ms  fetch {1e3*(t3-t2):.2f} ms", file=sys.stderr)
