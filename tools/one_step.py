"""One workload, a few resident steps (for ncu captures): one_step.py <config> [scale] [auto|staged|fused]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: F401,E402

from crispr_hawk_b200 import synth  # noqa: E402
from crispr_hawk_b200.workload import UnphasedWorkload, Workload  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
mode = sys.argv[3] if len(sys.argv) > 3 else "auto"
fused = {"auto": None, "staged": False, "fused": True}[mode]
k = synth.CONFIGS[name]
W = UnphasedWorkload if k.get("unphased") else Workload
wl = W(synth.config_cohort(name, scale), k["pam"], k["guidelen"], k["right"])
wl.prepare_resident()
for _ in range(2):
    wl.step_resident(fused=fused).close()
