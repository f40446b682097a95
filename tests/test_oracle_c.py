"""Pin the C restatement (oracle/scan_oracle.c) against the reference-generated golden
vectors: phased, variant-free and unphased (resolve_guide) cases."""

import numpy as np
import pytest

from crispr_hawk_b200 import marshal
from crispr_hawk_b200.pam import pam_patterns
from oracle import c_oracle
from tests.helpers import all_golden_cases, fixture_objects, golden_guides, split_hits

CASES = all_golden_cases()


def run_c_oracle(case, threads=2):
    region, haps = fixture_objects(case)
    texts = [marshal.hap_text(h) for h in haps]
    buf, off, lens = marshal.stage_ascii(texts)
    fwd, rc = pam_patterns(case["pam"])
    bounds = [marshal.scan_bounds(h, region.start, region.stop, len(fwd)) for h in haps]
    is_ref = [h.samples == "REF" for h in haps]
    seg = marshal.segment_table(haps)
    unphased = bool(case["variants_present"] and not case["phased"])
    return haps, c_oracle.search(buf, off, lens, [b[0] for b in bounds], [b[1] for b in bounds], is_ref,
                                 seg, fwd, rc, case["guidelen"], case["right"], threads=threads,
                                 unphased=unphased, alleles=marshal.allele_table(haps) if unphased else None)  # fmt: skip


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_c_oracle_matches_golden(case):
    haps, out = run_c_oracle(case)
    hits = [[f, r] for f, r in zip(split_hits(out["hits"][0], len(haps)), split_hits(out["hits"][1], len(haps)))]
    assert hits == case["pam_hits"]
    want = golden_guides(case)
    got = [
        (int(out["start"][i]), int(out["stop"][i]), int(out["strand"][i]),
         out["text"][i].tobytes().decode(), haps[int(out["hap"][i])].id)
        for i in range(len(out["hap"]))
    ]  # fmt: skip
    assert got == [(g.start, g.stop, g.strand, g.sequence, g.hapid) for g in want]


def test_c_encode():
    assert c_oracle.encode(b"ACGTNacgtn").tolist() == [1, 2, 4, 8, 15] * 2
    with pytest.raises(ValueError):
        c_oracle.encode(b"ACGU")
