"""Edit-list path, where the time goes: hawk_search_stream_edits by group count (with / without the
window text), next to its pieces run alone (batch from edits, resident search, table fetch)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from crispr_hawk_b200 import _cabi, synth  # noqa: E402
from crispr_hawk_b200.workload import Workload  # noqa: E402

k = synth.CONFIGS["c2"]
wl = Workload(synth.config_cohort("c2", 1.0), k["pam"], k["guidelen"], k["right"])


def timed(fn, n=4):
    fn()
    best = 1e9
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, 1e3 * (time.perf_counter() - t0))
    return round(best, 2)


out = {}
for text in (True, False):
    for g in (1, 2, 4, 6, 8, 12, 16, 24, 32):
        out[f"stream text={int(text)} groups={g}"] = timed(lambda: wl.step_edits(n_groups=g, want_text=text))
    out[f"stream text={int(text)} groups=auto"] = timed(lambda: wl.step_edits(want_text=text))
e = wl._edit_buffers()
c = wl.cohort
mk = lambda: _cabi.Batch.from_edits(wl.ctx, e["ref"], c.region_start, e["off"], e["pos"], e["rl"], e["al"], e["ao"], e["pool"])  # noqa: E731
out["batch_from_edits (whole cohort)"] = timed(lambda: mk().close())
b = mk()
out["search on it"] = timed(lambda: _cabi.search(wl.ctx, b, wl.params, wl.a, wl.b, wl.d.is_ref).close())
r = _cabi.search(wl.ctx, b, wl.params, wl.a, wl.b, wl.d.is_ref)
bufs = _cabi.alloc_table(r.n_guides + 16, r.text_stride, pinned=True)
out["fetch table"] = timed(lambda: r.table(bufs))
out["fetch rows only"] = timed(lambda: r.table(bufs, want_text=False))
out["rows"] = int(r.n_guides)
print(json.dumps(out, indent=1))
