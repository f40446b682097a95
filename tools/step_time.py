"""Timing aid: resident steps of a workload, fused and staged, with per-kind kernel times."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from crispr_hawk_b200 import synth  # noqa: E402
from crispr_hawk_b200.workload import UnphasedWorkload, Workload  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
k = synth.CONFIGS[name]
c = synth.config_cohort(name, scale)
W = UnphasedWorkload if k.get("unphased") else Workload
wl = W(c, k["pam"], k["guidelen"], k["right"])
wl.prepare_resident()
for fused in (False, True, None, False, None):
    for _ in range(3):
        wl.step_resident(fused=fused).close()
    torch.cuda.synchronize()
    t = time.perf_counter()
    n = 10
    for _ in range(n):
        r = wl.step_resident(fused=fused)
        r.close()
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t) / n
    wl.ctx.set_profiling(True)
    wl.ctx.profile()
    for _ in range(3):
        wl.step_resident(fused=fused).close()
    prof = {k2: round(v[0] / 3, 3) for k2, v in wl.ctx.profile().items()}
    wl.ctx.set_profiling(False)
    print(f"{name} x{scale} fused={fused}: {ms:.3f} ms/step  rows {r.n_guides:,}  {prof}", file=sys.stderr)
