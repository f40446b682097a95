"""GPU parity of N2 (hawk_result_annotate + crispr_hawk_b200.annotation): variants visible in
each guide, their allele-frequency strings, reverse-complemented sequences and GC fractions
against what the unmodified reference's annotation.py left in its Guide objects
(tests/golden/annot.json.gz), and against the N2 oracle on device-built haplotypes."""

import gzip
import json
import os

import numpy as np
import pytest

import crispr_hawk_b200 as hawk
from crispr_hawk_b200 import _cabi, marshal
from oracle import annot_oracle as A
from tests.helpers import GOLDEN_DIR, all_golden_cases, fixture_objects

pytestmark = pytest.mark.gpu

with gzip.open(os.path.join(GOLDEN_DIR, "annot.json.gz"), "rb") as fh:
    ANNOT = json.loads(fh.read().decode())
CASES = [c for c in all_golden_cases() if c["name"] in ANNOT]


def annotated_rows(case, haps, region, packed):
    pam = hawk.PAM(case["pam"], case["right"], True)
    pam.encode(0)
    table, res = hawk.search_table(pam, region, haps, packed, case["guidelen"], case["right"],
                                   case["variants_present"], case["phased"], 0, True)  # fmt: skip
    cols = hawk.annotate_table(table, res, packed.batch, haps, case["right"])
    res.close()
    order = np.argsort(table["bucket"], kind="stable").tolist()
    return [[cols["variants"][i], cols["afs_str"][i], cols["sequence"][i], cols["right"][i], cols["gc"][i]] for i in order], table, order


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_annotation_matches_reference_golden(case):
    region, haps = fixture_objects(case)
    if not haps:
        pytest.skip("no haplotypes")
    packed = hawk.encode_region(haps, 0, True)
    got, _, _ = annotated_rows(case, haps, region, packed)
    want = ANNOT[case["name"]]
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == w, f"guide {i}: {g} != {w}"


def test_annotation_on_edit_list_haplotypes_matches_oracle():
    """Batches built from edit lists keep them as the variant table (no hawk_batch_set_variants)."""
    rng = np.random.default_rng(21)
    L, g0 = 3000, 5000
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, L))
    hap_edits = [[]]
    for _ in range(6):
        edits, p = [], 120
        while p < L - 150:
            p += int(rng.integers(8, 60))
            kind = rng.random()
            if kind < 0.6:
                alt = rng.choice([b for b in "ACGT" if b != ref[p]])
                edits.append(hawk.Edit(g0 + p, ref[p], str(alt)))
            elif kind < 0.8:
                k = int(rng.integers(1, 6))
                edits.append(hawk.Edit(g0 + p, ref[p], ref[p] + "".join("ACGT"[i] for i in rng.integers(0, 4, k))))
            else:
                k = int(rng.integers(1, 6))
                edits.append(hawk.Edit(g0 + p, ref[p : p + k + 1], ref[p]))
                p += k
        hap_edits.append(edits)
    afs = [{f"chr1-{e.pos}-{e.ref}/{e.alt}": round(float(rng.random()), 4) for e in edits} for edits in hap_edits]
    haps, packed = hawk.build_phased(ref, g0, hap_edits, afs=afs)

    class Region:
        contig, start, stop = "chr1", g0, g0 + L - 1
        coordinates = "chr1:x-y"

    for pamseq, G, right in (("NGG", 20, False), ("TTTV", 23, True)):
        case = {"pam": pamseq, "right": right, "guidelen": G, "variants_present": True, "phased": True}
        got, table, order = annotated_rows(case, haps, Region, packed)
        assert packed.batch.has_variants  # kept from the edit lists
        P = len(pamseq)
        n_alt = 0
        for row, i in zip(got, order):
            h = haps[int(table["hap"][i])]
            s = int(table["strand"][i])
            rp = (not right) if s == 1 else right
            pivot = int(table["pos"][i]) - (0 if rp else G)
            pm = [h.posmap[pivot + j] for j in range(G + P)]
            seq = table["text"][i].tobytes().decode("ascii")
            want = A.annotate_guide(seq, G, P, s, rp, int(table["stop"][i]), pm, h.variants, h.afs)
            assert row == list(want)
            n_alt += h.variants != "NA"
        assert n_alt > 50


def test_variant_table_normalisation():
    class H:
        def __init__(self, v):
            self.variants = v

    vt = marshal.variant_table([H("NA"), H("chr1-20-AC/ACGG,chr1-10-ACGT/AC,chr1-15-A/G")])
    assert vt.var_off.tolist() == [0, 0, 3]
    assert vt.var_pos.tolist() == [11, 15, 21] and vt.var_reflen.tolist() == [3, 1, 1] and vt.var_altlen.tolist() == [1, 1, 3]
    assert vt.alt_pool.tobytes() == b"CGCGG" and vt.ids[1] == ["chr1-10-ACGT/AC", "chr1-15-A/G", "chr1-20-AC/ACGG"]


@pytest.mark.parametrize("case", CASES[:12], ids=[c["name"] for c in CASES[:12]])
def test_annotation_seam_on_guide_objects(case):
    """The drop-in form: search() returns a GuideList, the four mirrors of annotation.py's loops
    leave in the Guide objects what the reference leaves in its own."""
    from crispr_hawk_b200 import annotation as ann

    region, haps = fixture_objects(case)
    if not haps:
        pytest.skip("no haplotypes")
    pam = hawk.PAM(case["pam"], case["right"], True)
    pam.encode(0)
    guides = hawk.search(pam, region, haps, None, case["guidelen"], case["right"], case["variants_present"],
                         case["phased"], 0, True)  # fmt: skip
    guides = ann._annotate_variants(guides, 0, True)
    guides = ann.annotate_variants_afs(guides, 0)
    guides = ann.reverse_guides(guides, 0)
    guides = ann.gc_content(guides, 0, True)
    want = ANNOT[case["name"]]
    assert len(guides) == len(want)
    for g, w in zip(guides, want):
        assert [g.variants, g.afs_str, g.sequence, bool(g.right), g.gc] == w


@pytest.mark.parametrize("case", CASES[:10], ids=[c["name"] for c in CASES[:10]])
def test_guide_table_wire_format(case):
    """N3: the SoA table as the wire format -- lazy Guide views equal the reference's annotated
    guides, and the batched scorer input equals scoring.py:49-84 applied to them."""
    region, haps = fixture_objects(case)
    if not haps:
        pytest.skip("no haplotypes")
    pam = hawk.PAM(case["pam"], case["right"], True)
    pam.encode(0)
    packed = hawk.encode_region(haps, 0, True)
    table, res = hawk.search_table(pam, region, haps, packed, case["guidelen"], case["right"],
                                   case["variants_present"], case["phased"], 0, True)  # fmt: skip
    cols = hawk.annotate_table(table, res, packed.batch, haps, case["right"])
    res.close()
    gt = hawk.GuideTable(table, haps, pam, case["guidelen"], case["right"], annotation=cols, debug=True)
    want = ANNOT[case["name"]]
    assert len(gt) == len(want)
    got = [[g.variants, g.afs_str, g.sequence, bool(g.right), g.gc] for g in gt]
    assert got == want
    assert gt[len(gt) - 1] is gt[-1]  # views are cached: mutations by downstream code persist
    assert gt.scorer_sequences() == [w[2][6:-7].upper() for w in want]
    assert gt.scorer_sequences(sgdesigner=True) == [w[2][10:-7].upper() for w in want]
    plain = hawk.GuideTable(table, haps, pam, case["guidelen"], case["right"])
    assert [g.sequence for g in plain] == [bytes(r).decode() for r in plain.sequences()]


def test_two_variants_at_one_position_are_refused():
    """hawk_batch_set_variants: two variants at one normalised position of a haplotype have no
    well-defined annotation in the reference (set order); the library says so instead of picking."""
    case = next(c for c in CASES if c["phased"] and any(h["variants"] != "NA" for h in c["haps"]))
    region, haps = fixture_objects(case)
    packed = hawk.encode_region(haps, 0, True)
    vt = marshal.variant_table(haps)
    assert not vt.ambiguous
    packed.batch.set_variants(vt)
    h = next(i for i in range(len(haps)) if vt.var_off[i + 1] > vt.var_off[i])
    j = int(vt.var_off[h])
    dup = marshal.VariantTable(vt.var_off.copy(), vt.var_pos.copy(), vt.var_reflen, vt.var_altlen, vt.var_altoff, vt.alt_pool, vt.ids)
    dup.var_off[h + 1 :] += 1
    for name in ("var_pos", "var_reflen", "var_altlen", "var_altoff"):
        a = getattr(vt, name)
        setattr(dup, name, np.concatenate((a[: j + 1], a[j : j + 1], a[j + 1 :])))
    with pytest.raises(_cabi.HawkLibraryError) as ei:
        packed.batch.set_variants(dup)
    assert ei.value.code == _cabi.HAWK_EINVAL and "two variants" in str(ei.value)
