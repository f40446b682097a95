"""Timing aid: BASELINE config 4 (unphased) resident steps with the host-side trace on."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: F401,E402

from crispr_hawk_b200 import synth  # noqa: E402
from crispr_hawk_b200.workload import UnphasedWorkload  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
k = synth.CONFIGS["c4"]
t = time.perf_counter()
c = synth.config_cohort("c4", scale)
wl = UnphasedWorkload(c, k["pam"], k["guidelen"], k["right"])
print(f"workload built in {time.perf_counter() - t:.1f}s: {wl.d.n_hap} haplotypes, {wl.scanned_bp:,} hap-bp", file=sys.stderr)
wl.prepare_resident()
for i in range(4):
    t = time.perf_counter()
    r = wl.step_resident()
    ms = 1e3 * (time.perf_counter() - t)
    print(f"step {i}: {ms:.2f} ms, {r.n_guides:,} rows, hits {r.n_hits}", file=sys.stderr)
    r.close()
wl.ctx.set_profiling(True)
wl.ctx.profile()
for i in range(3):
    wl.step_resident().close()
print({k2: (round(v[0] / 3, 3), v[1] // 3) for k2, v in wl.ctx.profile().items()}, file=sys.stderr)
t = time.perf_counter()
tab, h2d, d2h = wl.step_host()
tab, h2d, d2h = wl.step_host()
t = time.perf_counter()
tab, h2d, d2h = wl.step_host()
print(f"host step: {1e3 * (time.perf_counter() - t):.1f} ms, h2d {h2d:,} d2h {d2h:,}", file=sys.stderr)
