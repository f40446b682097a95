"""HAWK_TRACE of one streamed call (texts or edits) on the bench workload."""
import os
import sys

import torch

sys.path.insert(0, ".")
from crispr_hawk_b200 import synth  # noqa: E402
from crispr_hawk_b200.workload import Workload  # noqa: E402

mode, groups = sys.argv[1], int(sys.argv[2])
k = synth.CONFIGS["c2"]
c = synth.config_cohort("c2", 1.0, n_alt_hap=k["n_alt_hap"])
wl = Workload(c, k["pam"], k["guidelen"], k["right"])
wl.host_buffers()
fn = (lambda: wl.step_host(n_groups=groups)) if mode == "texts" else (lambda: wl.step_edits(n_groups=groups))
fn()
fn()
torch.cuda.synchronize()
sys.stderr.write("==== traced call ====\n")
sys.stderr.flush()
import time
t0 = time.perf_counter()
fn()
sys.stderr.write(f"==== end: {1e3 * (time.perf_counter() - t0):.2f} ms ====\n")
