import os, sys
sys.path.insert(0, "/root/repo")
import torch
from crispr_hawk_b200 import synth
from crispr_hawk_b200.workload import UnphasedWorkload
k = synth.CONFIGS["c4"]
c = synth.config_cohort("c4", 1.0)
wl = UnphasedWorkload(c, k["pam"], k["guidelen"], k["right"])
wl.prepare_resident()
for i in range(2):
    wl.step_resident().close()
