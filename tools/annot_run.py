"""One hawk_result_annotate over a config-2-shaped table (for ncu captures of the N2 kernels)."""
import sys

sys.path.insert(0, ".")
from crispr_hawk_b200 import synth  # noqa: E402
from crispr_hawk_b200.workload import Workload  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
k = synth.CONFIGS["c2"]
wl = Workload(synth.config_cohort("c2", scale, n_alt_hap=int(k["n_alt_hap"] * scale)), k["pam"], k["guidelen"], k["right"])
print(wl.annotate_measure(reps=2))
