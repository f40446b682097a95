"""GPU parity of the N2 row collapse (hawk_result_collapse + crispr_hawk_b200/report_rows.py)
against the collapsed report rows of the unmodified reference (tests/golden/report.json.gz),
and of the device grouping against a plain grouping of the same table at workload scale."""

import numpy as np
import pytest

from crispr_hawk_b200 import synth
from crispr_hawk_b200.workload import Workload
from tests.helpers import load_golden
from tests.test_report_rows import collapsed_rows

pytestmark = pytest.mark.gpu

REPORT = load_golden("report")


@pytest.mark.parametrize("case", REPORT, ids=[c["name"] for c in REPORT])
def test_collapsed_rows_match_reference_report(case):
    n, got = collapsed_rows(case)
    assert n == case["n_guides"] and len(got) == len(case["rows"])
    for k, (g, w) in enumerate(zip(got, case["rows"])):
        assert g == w, f"row {k}: {g} != {w}"


def test_device_groups_equal_a_host_grouping_at_scale():
    """c2-shaped cohort (NGG, 20 nt), some 0.5 M rows: every group of the device order holds exactly
    the rows sharing (start, stop, strand, origin, core text); groups ascend by (start, stop);
    rows inside a group ascend (emission order)."""
    k = synth.CONFIGS["c2"]
    c = synth.config_cohort("c2", 0.1, n_alt_hap=2000)
    wl = Workload(c, k["pam"], k["guidelen"], k["right"])
    res = wl.step_resident()
    table = res.table()
    is_ref = np.zeros(c.n_hap, np.uint8)
    is_ref[0] = 1
    perm, head, collision = res.collapse(is_ref)
    res.close()
    n = len(table["hap"])
    assert n > 300_000 and not collision
    assert np.array_equal(np.sort(perm), np.arange(n))
    core = table["text"][:, 10 : 10 + wl.guidelen + len(wl.fwd)]
    key = np.concatenate([table["start"].astype(">u4").view(np.uint8).reshape(n, 4), table["stop"].astype(">u4").view(np.uint8).reshape(n, 4),
                          table["strand"].reshape(n, 1), is_ref[table["hap"]].reshape(n, 1), core], axis=1)  # fmt: skip
    _, want_id = np.unique(key, axis=0, return_inverse=True)
    want_id = want_id.ravel()
    gid = np.cumsum(head) - 1
    got_id = np.empty(n, np.int64)
    got_id[perm] = gid
    # same partition: the (got, want) pairs are a bijection
    pairs = np.unique(np.stack([got_id, want_id], axis=1), axis=0)
    assert len(pairs) == gid[-1] + 1 == want_id.max() + 1
    p = perm.astype(np.int64)
    ss = (table["start"][p].astype(np.int64) << 32) | table["stop"][p].astype(np.int64)
    assert np.all(np.diff(ss) >= 0)
    assert np.all(np.diff(p)[head[1:] == 0] > 0)
    assert gid[-1] + 1 < n  # there was something to collapse
