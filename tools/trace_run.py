import sys, time
sys.path.insert(0, '/root/repo')
import torch
from crispr_hawk_b200 import synth, _cabi
from crispr_hawk_b200.workload import Workload
k = synth.CONFIGS["c2"]
c = synth.config_cohort("c2")
wl = Workload(c, k["pam"], k["guidelen"], k["right"])
wl.prepare_resident()
for i in range(4):
    t=time.perf_counter(); r = wl.step_resident(); r.close(); print("step ms", 1e3*(time.perf_counter()-t), file=sys.stderr)
