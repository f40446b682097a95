"""Live A/B of the CPU oracle against the unmodified reference (build container
only; skipped where /root/reference is absent)."""

import pytest

from oracle import hawk_oracle as O
from oracle import refshim
from tests.synth_cases import config1_cases, kat_cases, make_case, random_cases

pytestmark = pytest.mark.ref


def _ab(case):
    region, haps = refshim.build_case(
        case.ref_text, case.bed_start, case.bed_stop, case.vcf_lines, case.samples, case.phased
    )
    pam, bits, guides = refshim.run_search(
        region, haps, case.pam, case.guidelen, case.right, case.variants_present, case.phased
    )
    ohaps = [O.OracleHap.from_object(h) for h in haps]
    for h, b in zip(ohaps, bits):
        assert O.encode(h.seq) == b
    mine = O.search(
        case.pam, region.start, region.stop, ohaps, case.guidelen, case.right,
        case.variants_present, case.phased,
    )  # fmt: skip
    assert [g.guide_id for g in guides] == [g.guide_id for g in mine]
    assert [O.guide_tuple_from_object(g) for g in guides] == mine
    return len(mine)


def test_kats():
    assert [_ab(c) for c in kat_cases()] == [14, 9, 29, 37]


def test_config1():
    for c in config1_cases():
        assert _ab(c) > 500


@pytest.mark.parametrize("seed0", [1000, 1010, 1020])
def test_random_sweep(seed0):
    n = 0
    for c in random_cases(4, seed0=seed0):
        n += _ab(c)
    assert n > 1000


def test_pam_objects():
    ref = refshim.load()
    for p in ["NGG", "TTTV", "NNGRRT", "TTN", "NGK", "YTTV", "NNNNGATT", "TTCN", "NRG"]:
        r = ref.pam.PAM(p, False, True)
        r.encode(0)
        o = O.OraclePam(p)
        assert (r.pam, r.pamrc, r.bits, r.bitsrc, r.bits_list) == (
            o.pam, o.pamrc, o.bits, o.bitsrc, o.bits_list,
        )  # fmt: skip
