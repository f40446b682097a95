"""N2 row collapse, host side (crispr_hawk_b200/report_rows.py) against the rows the unmodified
reference's `_construct_report` + `_collapse_report_entries` produced (tests/golden/report.json.gz,
generator: tests/golden/make_golden_report.py). On this box the device layer is the kernels' own
core compiled for the CPU (tests/fake_backend.py); tests/test_gpu_report.py runs the same cases
through hawk_result_collapse on the GPU."""

import numpy as np
import pytest

import crispr_hawk_b200 as hawk
from crispr_hawk_b200 import annotation, report_rows
from tests import fake_backend
from tests.helpers import fixture_objects, load_golden

REPORT = load_golden("report")


def collapsed_rows(case):
    region, haps = fixture_objects(case)
    packed = hawk.encode_region(haps, 0, True)
    pam = hawk.PAM(case["pam"], case["right"], True)
    pam.encode(0)
    table, res = hawk.search_table(pam, region, haps, packed, case["guidelen"], case["right"], True, True, 0, True)
    cols = hawk.annotate_table(table, res, packed.batch, haps, case["right"])
    groups = annotation.report_groups(table, res, haps)
    res.close()
    out = report_rows.collapse_table(table, groups, cols, haps, case["contig"], case["target"], case["pam"], case["guidelen"])
    return len(table["hap"]), [[out[c][k] for c in case["columns"]] for k in range(len(out["chr"]))]


@pytest.mark.parametrize("case", REPORT, ids=[c["name"] for c in REPORT])
def test_collapsed_rows_match_reference_report(case, monkeypatch):
    fake_backend.activate(monkeypatch)
    n, got = collapsed_rows(case)
    assert n == case["n_guides"] and len(got) == len(case["rows"])
    for k, (g, w) in enumerate(zip(got, case["rows"])):
        assert g == w, f"row {k}: {g} != {w}"


def test_sample_polish_and_joins():
    assert report_rows.polish_samples_phased("s1:0/1,s2:1/1") == "s1:0/1,s2:1/1"
    assert report_rows.polish_samples_phased("s1:0|1,s1:1|0,s2:1|0") == "s1:1|1,s2:1|0"
    assert report_rows.collapse_samples(["s2:1|0,s1:0|1", "s1:1|0"]) == "s1:1|1,s2:1|0"
    assert report_rows.collapse_samples(["REF"]) == "REF" and report_rows.collapse_samples([]) == ""
    assert report_rows.collapse_haplotype_ids(["h3,h1", "h2", "h1"]) == "h1,h2,h3"
    assert report_rows.check_variant_ids(["b,a", "a,b"]) == "a,b" and report_rows.check_variant_ids(["NA"]) == "NA"
    assert report_rows.pam_class("NNGRRT") == "[ACGT][ACGT]G[AG][AG]T" and report_rows.pam_class("TTTV") == "TTT[ACG]"
    assert report_rows.split_core("a" * 10 + "GGGGG" + "TTT" + "c" * 10, False, 5, 3) == ("GGGGG", "TTT")
    assert report_rows.split_core("a" * 10 + "TTT" + "GGGGG" + "c" * 10, True, 5, 3) == ("GGGGG", "TTT")


def test_groups_after_a_hash_collision_are_joined_on_the_exact_key():
    """A collision leaves equal keys in two runs (A B A): the host joins them on the bytes."""
    perm = np.array([4, 0, 2, 1, 3], np.uint32)
    head = np.array([1, 1, 1, 0, 1], np.uint8)
    key = {4: "x", 0: "A", 2: "B", 1: "B", 3: "A"}
    plain = report_rows.groups_of(perm, head, False)
    assert [g.tolist() for g in plain] == [[4], [0], [2, 1], [3]]
    joined = report_rows.groups_of(perm, head, True, lambda i: key[i])
    assert sorted(g.tolist() for g in joined) == [[0, 3], [1, 2], [4]]
    assert report_rows.groups_of(np.empty(0, np.uint32), np.empty(0, np.uint8), False) == []
