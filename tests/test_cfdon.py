"""N4 (CFDon): crispr_hawk_b200.scoring + hawk_result_cfdon against the scores the unmodified
reference computed (tests/golden/cfdon.json.gz, generator: tests/golden/make_golden_cfdon.py) --
score strings as the Guide holds them and the raw floats bit for bit. CPU: the kernels' own
cfdon_row compiled for the host (tests/fake_backend.py); tests/test_gpu_cfdon.py: the GPU."""

import numpy as np
import pytest

import crispr_hawk_b200 as hawk
from crispr_hawk_b200 import scoring
from tests import fake_backend
from tests.helpers import fixture_objects, load_golden

CFD = load_golden("cfdon")


def synthetic_cfd_dicts(seed):
    """The generator's tables (tests/test_real_driver.synthetic_cfd_dicts, restated: that module
    needs the reference)."""
    import random

    rnd = random.Random(seed)
    rc = {"A": "T", "C": "G", "G": "C", "U": "A"}
    mm = {f"r{w}:d{rc[g]},{i + 1}": rnd.random() for i in range(20) for w in "ACGU" for g in "ACGU" if w != g}
    pam = {a + b: rnd.random() for a in "ACGT" for b in "ACGT"}
    pam["GG"] = 1.0
    return mm, pam


def cfdon_rows(case):
    region, haps = fixture_objects(case)
    packed = hawk.encode_region(haps, 0, True)
    pam = hawk.PAM(case["pam"], case["right"], True)
    pam.encode(0)
    table, res = hawk.search_table(pam, region, haps, packed, case["guidelen"], case["right"], True, True, 0, True)
    is_ref = np.array([h.samples == "REF" for h in haps], np.uint8)
    mm, pam2 = scoring.cfd_tables(*synthetic_cfd_dicts(CFD["table_seed"]))
    col = res.cfdon(is_ref, mm, pam2)
    res.close()
    order = np.argsort(table["bucket"], kind="stable")
    return [[int(table["start"][i]), int(table["strand"][i]), haps[int(table["hap"][i])].id,
             "NA" if np.isnan(col[i]) else str(round(float(col[i]), 4)), float(col[i]).hex()] for i in order]  # fmt: skip


def check(case):
    got = cfdon_rows(case)
    want = case["rows"]
    assert len(got) == len(want)
    for k, (g, w) in enumerate(zip(got, want)):
        assert g == w, f"guide {k}: {g} != {w}"
    assert any(r[3] not in ("NA", "1.0") for r in got)


@pytest.mark.parametrize("case", CFD["cases"], ids=[c["name"] for c in CFD["cases"]])
def test_cfdon_matches_reference_scores(case, monkeypatch):
    fake_backend.activate(monkeypatch)
    check(case)


def test_tables_follow_the_reference_keys():
    mmd, pamd = synthetic_cfd_dicts(3)
    mm, pam2 = scoring.cfd_tables(mmd, pamd)
    assert mm.shape == (20, 4, 4) and pam2.shape == (4, 4)
    assert mm[4, 0, 3] == mmd["rA:dA,5"] and mm[0, 3, 1] == mmd["rU:dG,1"] and mm[19, 2, 0] == mmd["rG:dT,20"]
    assert np.isnan(mm[:, range(4), range(4)]).all() and not np.isnan(mm[:, ~np.eye(4, dtype=bool)]).any()
    assert pam2[2, 2] == pamd["GG"] and pam2[0, 3] == pamd["AT"]
    del mmd["rC:dC,7"]
    assert np.isnan(scoring.cfd_tables(mmd, pamd)[0][6, 1, 2])
