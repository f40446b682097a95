"""Pin the CPU oracle against reference-generated golden vectors and the
reference's own known-answer tests (tests/test_encoder.py:6-46,
tests/test_pam.py:19-24, tests/test_utils.py:9-17 in /root/reference)."""

import pytest

from oracle import hawk_oracle as O
from tests.helpers import all_golden_cases, fixture_objects, golden_guides, load_golden

CASES = all_golden_cases()


def test_reference_known_answers():
    # reference tests/test_encoder.py:6-13,36-55
    assert O.encode("ACGTN") == [1, 2, 4, 8, 15]
    assert O.encode("acgtn") == [1, 2, 4, 8, 15]
    assert O.encode("") == []
    assert len(O.IUPAC_BITS) == 15
    with pytest.raises(O.OracleError):
        O.encode("ACXT")
    # reference tests/test_pam.py:19-24, SURVEY.md A2 probes
    p = O.OraclePam("NGG")
    assert (p.pam, p.pamrc, p.bits, p.bitsrc) == ("NGG", "CCN", 0xF44, 0x22F)
    q = O.OraclePam("TTTV")
    assert (q.pamrc, q.bits, q.bitsrc) == ("BAAA", 0x8887, 0xE111)
    # reference tests/test_utils.py:9-17
    assert O.reverse_complement("ACGT") == "ACGT"
    assert O.reverse_complement("AAGC") == "GCTT"
    assert O.reverse_complement("NNGRRT") == "AYYCNN"


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_golden(case):
    region, haps = fixture_objects(case)
    ohaps = [O.OracleHap.from_object(h) for h in haps]
    pam = O.OraclePam(case["pam"])
    assert pam.bits == case["pam_bits"] and pam.bitsrc == case["pam_bitsrc"]
    assert pam.pamrc == case["pam_rc"]
    bounds = [
        list(O.compute_scan_start_stop(h, region.start, region.stop, len(pam))) for h in ohaps
    ]
    assert bounds == case["scan_bounds"]
    bits = [O.encode(h.seq) for h in ohaps]
    hits = O.pam_search(pam, region.start, region.stop, ohaps, bits)
    assert [[f, r] for f, r in hits] == case["pam_hits"]
    got = O.search(
        case["pam"], region.start, region.stop, ohaps, case["guidelen"], case["right"],
        case["variants_present"], case["phased"],
    )  # fmt: skip
    assert got == golden_guides(case)


def test_kat1_literal():
    """scan bounds, raw PAM hits and the first/last guide."""
    case = load_golden("kat")[0]
    assert case["scan_bounds"] == [[100, 177]]
    assert case["pam_hits"] == [
        [[108, 134, 139], [105, 115, 122, 127, 147, 148, 157, 165, 166, 167, 168]]
    ]
    g = case["guides"]
    assert len(g) == 14
    assert g[0][:5] == [989, 1012, 0, "TTAGGTATGTCTTAGTGACTCTAAATACCAAGGCAGTCCTCGA", False]
    assert g[-1][:5] == [1069, 1092, 1, "CAATCTACCCCCTGTTATGCGCGTTTGTCGTTAGACCAATGTC", True]
