"""GPU parity of N4 (hawk_result_cfdon) against the reference's CFDon scores
(tests/golden/cfdon.json.gz), floats bit for bit; error path; a workload-scale run."""

import numpy as np
import pytest

from crispr_hawk_b200 import _cabi, scoring, synth
from crispr_hawk_b200.workload import Workload
from tests.test_cfdon import CFD, check, synthetic_cfd_dicts

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CFD["cases"], ids=[c["name"] for c in CFD["cases"]])
def test_cfdon_matches_reference_scores(case):
    check(case)


def test_cfdon_at_scale_and_missing_keys():
    k = synth.CONFIGS["c2"]
    c = synth.config_cohort("c2", 0.1, n_alt_hap=1500)
    wl = Workload(c, k["pam"], k["guidelen"], k["right"])
    res = wl.step_resident()
    table = res.table()
    is_ref = np.zeros(c.n_hap, np.uint8)
    is_ref[0] = 1
    mmd, pamd = synthetic_cfd_dicts(11)
    mm, pam2 = scoring.cfd_tables(mmd, pamd)
    col = res.cfdon(is_ref, mm, pam2)
    n = len(col)
    assert n == len(table["hap"]) > 200_000
    ref_rows = table["hap"] == 0
    # every key that has a REF guide is scored, the others are NaN
    first = table["bucket"].astype(np.int64)
    has_ref = ref_rows[first]
    assert np.isnan(col[~has_ref]).all() and not np.isnan(col[has_ref]).any()
    # spot-check 300 rows against the sequential Python product
    rng = np.random.default_rng(0)
    G, P = wl.guidelen, len(wl.fwd)
    comp = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
    for i in rng.choice(np.flatnonzero(has_ref), 300, replace=False):
        core = lambda r: table["text"][r, 10 : 10 + G + P].tobytes()  # noqa: E731
        a, b = core(first[i]), core(i)
        if table["strand"][i]:
            a, b = a.translate(comp)[::-1], b.translate(comp)[::-1]
        a, b = a.upper().decode(), b.upper().decode()
        s = 1.0
        for j in range(G):
            if a[j] != b[j]:
                s *= mm[j, "ACGT".index(a[j]), "ACGT".index(b[j])]
        s *= pam2["ACGT".index(b[G + P - 2]), "ACGT".index(b[G + P - 1])]
        assert s == col[i]
    # a missing key is an error, and says which row
    mmd.pop("rA:dA,3")
    mm2, _ = scoring.cfd_tables(mmd, pamd)
    with pytest.raises(_cabi.HawkLibraryError) as ei:
        res.cfdon(is_ref, mm2, pam2)
    assert ei.value.code == _cabi.HAWK_ECFD and 0 <= ei.value.bad_row < n
    res.close()
