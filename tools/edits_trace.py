"""HAWK_TRACE of the edit-list path (host-side phase times of one warm call each)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from crispr_hawk_b200 import _cabi, synth  # noqa: E402
from crispr_hawk_b200.workload import Workload  # noqa: E402

k = synth.CONFIGS["c2"]
wl = Workload(synth.config_cohort("c2", 1.0), k["pam"], k["guidelen"], k["right"])
e = wl._edit_buffers()
c = wl.cohort
mk = lambda: _cabi.Batch.from_edits(wl.ctx, e["ref"], c.region_start, e["off"], e["pos"], e["rl"], e["al"], e["ao"], e["pool"])  # noqa: E731
for _ in range(2):
    mk().close()
    wl.step_edits(n_groups=2, want_text=False)
torch.cuda.synchronize()
os.environ["HAWK_TRACE"] = "1"
print("== batch_from_edits, whole cohort", file=sys.stderr)
b = mk()
torch.cuda.synchronize()
print("== search on it", file=sys.stderr)
_cabi.search(wl.ctx, b, wl.params, wl.a, wl.b, wl.d.is_ref).close()
b.close()
print("== stream edits, 2 groups, rows only", file=sys.stderr)
wl.step_edits(n_groups=2, want_text=False)
