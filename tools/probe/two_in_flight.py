"""Throughput with two batches in flight on one GPU: two contexts (two streams), two host
threads, each running resident steps of its own copy of the workload. K1 is bound by its
arithmetic and the K2 / table kernels by memory latency, so the two overlap."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from crispr_hawk_b200 import _cabi, synth  # noqa: E402
from crispr_hawk_b200.workload import Workload  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
k = synth.CONFIGS[name]
cohort = synth.config_cohort(name, 1.0)
ctxs = [_cabi.Context.default(), _cabi.Context(0)]
wls = [Workload(cohort, k["pam"], k["guidelen"], k["right"], ctx=c) for c in ctxs]
for w in wls:
    w.prepare_resident()
    for _ in range(3):
        w.step_resident().close()
torch.cuda.synchronize()


def run(w, n, out):
    rows = 0
    for _ in range(n):
        r = w.step_resident()
        rows += r.n_guides
        r.close()
    out.append(rows)


N = 20
t0 = time.perf_counter()
out = []
run(wls[0], N, out)
torch.cuda.synchronize()
one = (time.perf_counter() - t0) / N
for rep in range(2):
    out = []
    th = [threading.Thread(target=run, args=(w, N // 2, out)) for w in wls]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    two = (time.perf_counter() - t0) / N
    print(f"{name}: one batch in flight {1e3 * one:.3f} ms/step; two in flight {1e3 * two:.3f} ms/step "
          f"({one / two:.2f}x), rows {sum(out):,}", file=sys.stderr)
