"""install() against the GENUINE reference driver: `crisprhawk.crisprhawk.encode_haplotypes`,
`guides_search` (crisprhawk.py:64-115) and `annotation.annotate_guides` (annotation.py:545-600)
run unmodified, once as they are and once with crispr_hawk_b200 installed, and must leave the
same Guide objects behind -- every field. The build container has the reference but no GPU, so
the device layer is the kernels' own core compiled for the CPU (tests/fake_backend.py); the
kernels themselves are compared with the same reference outputs on the GPU box
(tests/test_gpu_parity.py, tests/test_gpu_annot.py)."""

import sys
import types

import pytest

import crispr_hawk_b200 as hawk
from crispr_hawk_b200 import search_guides
from oracle import refshim
from tests import fake_backend
from tests.synth_cases import config1_cases, kat_cases, make_case, random_cases

pytestmark = pytest.mark.ref


def load_driver():
    """`import crisprhawk.crisprhawk` with the out-of-scope first-party modules it pulls in
    (scorers, graphics, off-target search, converter: SURVEY.md 8c) replaced by stubs."""
    refshim.load()
    for name, attrs in {
        "crisprhawk.scoring": dict(scoring_guides=lambda guides, *a: guides),
        "crisprhawk.scoring_envs": dict(ScoringEnvs=object),
        "crisprhawk.graphical_reports": dict(compute_graphical_reports=lambda *a: None),
        "crisprhawk.candidate_guides": dict(candidate_guides_analysis=lambda *a: None),
        "crisprhawk.search_offtargets": dict(offtargets_search=lambda guides, *a: guides),
        "crisprhawk.converter": dict(convert_gnomad_vcf=lambda *a: None),
        "crisprhawk.crisprme_data": dict(prepare_data_crisprme=lambda *a: None),
    }.items():
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
    import crisprhawk.crisprhawk as drv

    return drv


def guide_fields(g):
    pm = g.posmap
    # collapsed sample strings are joined from a Python set in the reference (haplotypes.py:258): compare as sets
    return (g.start, g.stop, g.strand, g.sequence, g.guidelen, g.pamlen, bool(g.right), frozenset(g.samples.split(",")), g.variants, g.afs_str,
            g.gc, g.hapid, g.guide_id, g.pam, g.guide, tuple(pm[j] for j in range(len(pm))), g.afs is not None)  # fmt: skip


def run_driver(drv, case):
    region, haps = refshim.build_case(case.ref_text, case.bed_start, case.bed_stop, case.vcf_lines, case.samples, case.phased)
    args = types.SimpleNamespace(guidelen=case.guidelen, right=case.right, verbosity=0, debug=True, annotations=[],
                                 gene_annotations=[], annotation_colnames=[], gene_annotation_colnames=[])  # fmt: skip
    pam = drv.encode_pam(case.pam, case.right, 0, True)
    bits = drv.encode_haplotypes({region: haps}, args)
    guides = drv.guides_search(pam, {region: haps}, bits, case.variants_present, case.phased, args)
    returned = guides[region]
    guides = drv.annotate_guides(guides, args)
    return region, returned, guides[region]


CASES = [c for c in config1_cases() if c.phased] + [kat_cases()[0], kat_cases()[2]] + [
    make_case(100 + k, phased=True, pam=p, guidelen=g, right=r) for k, (p, g, r) in enumerate([("NGG", 20, False), ("TTTV", 23, True), ("NNGRRT", 21, False)])
]  # fmt: skip


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_installed_driver_leaves_the_same_guides(case, monkeypatch):
    drv = load_driver()
    _, _, want = run_driver(drv, case)
    want = [guide_fields(g) for g in want]
    fake_backend.activate(monkeypatch)
    hawk.install()
    try:
        assert drv.search is search_guides.search
        _, returned, got = run_driver(drv, case)
        assert isinstance(returned, search_guides.GuideList)
        assert got is returned  # the seam returns the list it was given (the reference builds no new one either)
        assert returned.built == 0 or len(returned) == 0  # nothing touched an object so far: all stages are pending
        assert [guide_fields(g) for g in got] == want
        assert got.built == len(got)
    finally:
        hawk.uninstall()
    assert drv.search is not search_guides.search


def test_evicted_tables_fall_back_to_the_reference_functions(monkeypatch):
    """LIVE_TABLES forced small: the first region's table is released before it is annotated and
    its list takes the reference's own annotation functions -- same guides all the same."""
    drv = load_driver()
    case = config1_cases()[0]
    case2 = make_case(7, bed_len=900, n_sites=20, n_samples=5, phased=True, bed_start=50001)  # another region
    _, _, want = run_driver(drv, case)
    want = [guide_fields(g) for g in want]
    _, _, want2 = run_driver(drv, case2)
    want2 = [guide_fields(g) for g in want2]
    fake_backend.activate(monkeypatch)
    monkeypatch.setattr(search_guides, "LIVE_TABLES", search_guides._LiveTables(cap_bytes=1))
    hawk.install()
    try:
        region, haps = refshim.build_case(case.ref_text, case.bed_start, case.bed_stop, case.vcf_lines, case.samples, True)
        region2, haps2 = refshim.build_case(case2.ref_text, case2.bed_start, case2.bed_stop, case2.vcf_lines, case2.samples, True)
        args = types.SimpleNamespace(guidelen=20, right=False, verbosity=0, debug=True, annotations=[], gene_annotations=[])
        pam = drv.encode_pam("NGG", False, 0, True)
        hh = {region: haps, region2: haps2}
        guides = drv.guides_search(pam, hh, drv.encode_haplotypes(hh, args), True, True, args)
        first, second = guides[region], guides[region2]
        assert first.hawk["res"] is None and second.hawk["res"] is not None  # the older table was evicted
        guides = drv.annotate_guides(guides, args)
        assert [guide_fields(g) for g in guides[region]] == want
        assert [guide_fields(g) for g in guides[region2]] == want2
    finally:
        hawk.uninstall()


def test_lazy_list_behaves_like_the_list_it_replaces(monkeypatch):
    drv = load_driver()
    case = config1_cases()[0]
    fake_backend.activate(monkeypatch)
    hawk.install()
    try:
        _, guides, _ = run_driver(drv, case)
    finally:
        hawk.uninstall()
    n = len(guides)
    assert n > 100 and guides.built == 0 and "None" not in repr(guides)[:200] and guides.built == n  # repr builds
    fake_backend.activate(monkeypatch)
    region, haps = refshim.build_case(case.ref_text, case.bed_start, case.bed_stop, case.vcf_lines, case.samples, True)
    pam = hawk.PAM("NGG", False, True)
    pam.encode(0)
    lazy = hawk.search(pam, region, haps, hawk.encode_region(haps, 0, True), 20, False, True, True, 0, True)
    assert lazy.built == 0
    g5 = lazy[5]
    assert lazy.built == 1 and lazy[5] is g5 and lazy[-1] is lazy[n - 1] and lazy.built == 2
    assert [g.guide_id for g in lazy[2:4]] == [guides[2].guide_id, guides[3].guide_id] and lazy.built == 4
    both = lazy + []  # any whole-list operation builds the rest first
    assert lazy.built == n and type(both) is list and all(g is not None for g in both)
    assert [g.guide_id for g in lazy] == [g.guide_id for g in guides]
    assert sorted(lazy, key=lambda g: g.start)[0].start == min(g.start for g in guides)


# --------------------------------------------------------------------------- N1 through install()
N1_CASES = config1_cases() + [kat_cases()[2]] + random_cases(5)


def hap_fields(h):
    pm = h.posmap
    return (h.sequence.sequence, tuple(pm[i] for i in range(len(h))), h.start, h.stop, frozenset(h.samples.split(",")),
            h.variants, dict(h.afs))  # fmt: skip


@pytest.mark.parametrize("case", N1_CASES, ids=[c.name for c in N1_CASES])
def test_haplotype_seam_equals_reference_builder(case, monkeypatch):
    """install() rebinds crisprhawk.haplotypes.add_variants_phased: the reference's VariantRecord
    lists become device-built EditHaplotypes. Same haplotypes (text, position map, bounds, sample
    sets, variant ids, allele frequencies, order) and the same search() output as the reference's
    own builder; unphased regions fall through to it untouched."""
    ref = refshim.load()
    load_driver()
    region, want_haps = refshim.build_case(case.ref_text, case.bed_start, case.bed_stop, case.vcf_lines, case.samples, case.phased)
    _, _, want = refshim.run_search(region, want_haps, case.pam, case.guidelen, case.right, case.variants_present, case.phased)
    fake_backend.activate(monkeypatch)
    hawk.install()
    try:
        from crispr_hawk_b200 import haplotypes as HN

        assert ref.haplotypes.add_variants_phased is HN.add_variants_phased
        region2, haps = refshim.build_case(case.ref_text, case.bed_start, case.bed_stop, case.vcf_lines, case.samples, case.phased)
        if case.phased and case.vcf_lines:
            assert all(isinstance(h, HN.EditHaplotype) for h in haps)
        assert [hap_fields(h) for h in haps] == [hap_fields(h) for h in want_haps]
        for h, w in zip(haps, want_haps):
            assert marshal_bounds(h, region2, 3) == tuple(ref.search_guides.compute_scan_start_stop(w, region.start, region.stop, 3))
        pam = hawk.PAM(case.pam, case.right, True)
        pam.encode(0)
        bits = hawk.encode_region(haps, 0, True)
        if case.phased and case.vcf_lines:
            assert bits is haps[0]._region_pack  # no text ever goes back up: the batch built from the edits is searched
        got = hawk.search(pam, region2, haps, bits, case.guidelen, case.right, case.variants_present, case.phased, 0, True)
    finally:
        hawk.uninstall()
    assert ref.haplotypes.add_variants_phased is not HN.add_variants_phased

    def fields(g):
        pm = g.posmap
        return (g.start, g.stop, g.strand, g.sequence, bool(g.right), frozenset(g.samples.split(",")), g.variants, g.hapid,
                tuple(pm[j] for j in range(len(pm))), g.guide_id)  # fmt: skip

    assert [fields(g) for g in got] == [fields(g) for g in want]


def marshal_bounds(h, region, pamlen):
    from crispr_hawk_b200 import marshal

    return tuple(marshal.scan_bounds(h, region.start, region.stop, pamlen))


def test_unsupported_shapes_go_to_the_reference_builder(monkeypatch):
    """Two records overlapping on one chromosome copy: not an edit list the device builder
    takes -> the reference's own add_variants_phased builds the region."""
    ref = refshim.load()
    load_driver()
    text = kat_cases()[0].ref_text
    lines = ["chr1\t1020\t.\t" + text[119:123] + "\t" + text[119] + "\t.\tPASS\tAF=0.2\tGT\t1|0\t0|1",
             "chr1\t1021\t.\t" + text[120] + "\t" + ("A" if text[120] != "A" else "C") + "\t.\tPASS\tAF=0.2\tGT\t1|0\t0|0"]  # fmt: skip
    fake_backend.activate(monkeypatch)
    hawk.install()
    try:
        from crispr_hawk_b200 import haplotypes as HN

        try:
            _, haps = refshim.build_case(text, 1001, 1080, lines, ["S1", "S2"], True)
        except Exception as e:  # whatever the reference itself does with such records is what happens
            haps = e
        hawk.uninstall()
        try:
            _, want = refshim.build_case(text, 1001, 1080, lines, ["S1", "S2"], True)
        except Exception as e:
            want = e
        if isinstance(want, Exception):
            assert type(haps) is type(want)
        else:
            assert not any(isinstance(h, HN.EditHaplotype) for h in haps)
            assert [hap_fields(h) for h in haps] == [hap_fields(h) for h in want]
    finally:
        hawk.uninstall()


# --------------------------------------------------------------------------- N2: report rows
def run_report(drv, case):
    """reports.report_guides' two steps (reports.py:1029-1056) on the driver's annotated guides."""
    import crisprhawk.reports as R

    region, _, guides = run_driver(drv, case)
    pam = drv.encode_pam(case.pam, case.right, 0, True)
    rep = R._construct_report({region: guides}, pam, [], [], [], [], False, False)[region]
    return rep, (R._collapse_report_entries(rep, pam, [], [], False) if not rep.empty else rep)


REPORT_CASES = CASES[:4] + CASES[-3:] + [
    make_case(120, phased=True, pam="NNGRRT", guidelen=21, right=False, bed_len=1500, n_sites=60, n_samples=8, indel_frac=0.4)
]  # fmt: skip


@pytest.mark.parametrize("case", REPORT_CASES, ids=[c.name for c in REPORT_CASES])
def test_installed_report_collapse_equals_the_reference(case, monkeypatch):
    """The genuine `_construct_report` + `_collapse_report_entries`, plain and with the package
    installed (the collapse then uses the device-computed groups): identical DataFrames."""
    drv = load_driver()
    import crisprhawk.reports as R

    _, want = run_report(drv, case)
    fake_backend.activate(monkeypatch)
    hawk.install()
    try:
        assert R._collapse_report_entries.__module__ == "crispr_hawk_b200.report_rows"
        rep, got = run_report(drv, case)
        assert rep.empty or "hawk_groups" in rep.attrs  # the device groups were used, not the pandas groupby
        assert list(got.columns) == list(want.columns)
        assert len(got) == len(want) <= len(rep)
        # sample strings of N1-built haplotypes are joined in a different (set) order upstream: normalise
        for df in (got, want):
            if "samples" in df:
                df["samples"] = [",".join(sorted(s.split(","))) for s in df["samples"]]
        assert got.equals(want), got.compare(want)
    finally:
        hawk.uninstall()
    assert R._collapse_report_entries.__module__ == "crisprhawk.reports"


# --------------------------------------------------------------------------- N4: CFDon
def load_scoring():
    """The genuine `crisprhawk.scoring` (cfdon_score, group_guides_position) and
    `crisprhawk.scores.crisprhawk_scores.cfdon` / `cfdscore.compute_cfd`, with the learned scorers'
    packages (Azimuth, RS3, DeepCpf1, Elevation, PLM-CRISPR, CRISPRon, sgDesigner: matplotlib,
    rs3, model files -- out of scope) replaced by empty stubs."""
    import importlib
    import os

    refshim.load()
    import crisprhawk

    base = os.path.join(os.path.dirname(crisprhawk.__file__), "scores")

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    cur = sys.modules.get("crisprhawk.scoring")
    if cur is not None and getattr(cur, "__file__", None):
        return cur, sys.modules["crisprhawk.scores.crisprhawk_scores"]
    pkg = stub("crisprhawk.scores")
    pkg.__path__ = [base]
    for sub in ("azimuth", "deepCpf1", "elevation", "elevation.cmds", "plm_crispr", "crispron", "sgdesigner"):
        stub("crisprhawk.scores." + sub).__path__ = [os.path.join(base, *sub.split("."))]
    stub("crisprhawk.scores.azimuth.model_comparison", predict=None)
    stub("rs3")
    stub("rs3.seq", predict_seq=None)
    stub("crisprhawk.scores.deepCpf1.seqdeepcpf1", SeqDeepCpf1=None, preprocess=None, load_deepcpf1_weights=None,
         compute_deepcpf1=None)  # fmt: skip
    stub("crisprhawk.scores.elevation.cmds.predict", Predict=None)
    stub("crisprhawk.scores.plm_crispr.plm_crispr", compute_plm_crispr_score=None)
    stub("crisprhawk.scores.crispron.crispron", compute_crispron_score=None)
    stub("crisprhawk.scores.sgdesigner.sgdesigner", compute_sgdesigner_score=None)
    cs = importlib.import_module("crisprhawk.scores.crisprhawk_scores")
    for n in dir(cs):
        if not n.startswith("_"):
            setattr(pkg, n, getattr(cs, n))
    sys.modules.pop("crisprhawk.scoring", None)  # a stub of it may be there (load_driver)
    scoring = importlib.import_module("crisprhawk.scoring")
    return scoring, cs


def synthetic_cfd_dicts(seed, drop=None):
    """Factor tables with the reference's keys (cfdscore.py:89-94) and seeded values in (0, 1]."""
    import random

    rnd = random.Random(seed)
    rc = {"A": "T", "C": "G", "G": "C", "U": "A"}
    mm = {f"r{w}:d{rc[g]},{i + 1}": rnd.random() for i in range(20) for w in "ACGU" for g in "ACGU" if w != g}
    pam = {a + b: rnd.random() for a in "ACGT" for b in "ACGT"}
    pam["GG"] = 1.0
    for k in drop or ():
        mm.pop(k, None), pam.pop(k, None)
    return mm, pam


CFD_CASES = [CASES[0]] + [make_case(130 + k, phased=True, pam=p, guidelen=g, right=False, bed_len=1200, n_sites=50, n_samples=6,
                                    indel_frac=0.3) for k, (p, g) in enumerate([("NGG", 20), ("NRG", 20), ("NGG", 23), ("NGG", 18)])]  # fmt: skip


def run_cfdon(drv, scoring, case):
    region, _, guides = run_driver(drv, case)
    out = scoring.cfdon_score(guides, 0, True)
    return [(g.start, g.strand, g.guide, g.pam, g.hapid, g.cfdon_score) for g in out], guides, out


@pytest.mark.parametrize("case", CFD_CASES, ids=[c.name for c in CFD_CASES])
def test_installed_cfdon_equals_the_reference(case, monkeypatch):
    """The genuine `cfdon_score` -> `cfdon` -> `compute_cfd` with seeded factor tables in place of
    the model files, plain and with the package installed (scores then come from
    hawk_result_cfdon, computed while the table was on the device)."""
    scoring, cs = load_scoring()
    drv = load_driver()
    tables = synthetic_cfd_dicts(7)
    monkeypatch.setattr(cs, "load_mismatch_pam_scores", lambda debug: tables)
    monkeypatch.setattr(sys.modules["crisprhawk.scoring"], "cfdon_score", scoring.cfdon_score)
    want, _, _ = run_cfdon(drv, scoring, case)
    assert any(w[-1] not in ("NA", "1.0") for w in want) and any(w[-1] == "NA" for w in want)
    fake_backend.activate(monkeypatch)
    hawk.install()
    try:
        assert scoring.cfdon_score.__module__ == "crispr_hawk_b200.scoring"
        got, guides, out = run_cfdon(drv, scoring, case)
        assert out is guides and guides.hawk.get("cfdon") is not None  # the device column was used
        assert got == want
    finally:
        hawk.uninstall()
    assert scoring.cfdon_score.__module__ == "crisprhawk.scoring"


def test_cfdon_key_errors_surface_like_the_reference(monkeypatch):
    """A factor the tables do not hold: the reference's KeyError is wrapped by `cfdon_score` into
    CrisprHawkCfdScoreError (scoring.py:370-379); so is the device's HAWK_ECFD."""
    scoring, cs = load_scoring()
    drv = load_driver()
    case = CFD_CASES[1]
    full = synthetic_cfd_dicts(7)
    tables = synthetic_cfd_dicts(7, drop=[k for k in full[0] if k.endswith(",5")])
    monkeypatch.setattr(cs, "load_mismatch_pam_scores", lambda debug: tables)
    from crisprhawk.crisprhawk_error import CrisprHawkCfdScoreError

    with pytest.raises(CrisprHawkCfdScoreError):
        run_cfdon(drv, scoring, case)
    fake_backend.activate(monkeypatch)
    hawk.install()
    try:
        with pytest.raises(CrisprHawkCfdScoreError):
            run_cfdon(drv, scoring, case)
    finally:
        hawk.uninstall()


FEATURE_SEAM_CASES = [CASES[0], CASES[-3], CASES[-2], CASES[-1]]


@pytest.mark.parametrize("case", FEATURE_SEAM_CASES, ids=[c.name for c in FEATURE_SEAM_CASES])
def test_installed_scorer_inputs_equal_the_reference(case, monkeypatch):
    """The genuine `scoring._extract_guide_sequences` / `_extract_guide_sequences_sgdesigner`
    (scoring.py:50-84) on the annotated guides, plain and with the package installed (the strings
    then come from hawk_result_featurize, cut while the table was on the device); a list the
    package does not know goes to the reference's own functions."""
    scoring, _ = load_scoring()
    drv = load_driver()
    monkeypatch.setattr(sys.modules["crisprhawk.scoring"], "cfdon_score", scoring.cfdon_score)
    _, _, guides = run_driver(drv, case)
    want4, want0 = scoring._extract_guide_sequences(guides), scoring._extract_guide_sequences_sgdesigner(guides)
    assert len(want4) > 20 and len(want4[0]) == case.guidelen + len(case.pam) + 7
    fake_backend.activate(monkeypatch)
    hawk.install()
    try:
        assert scoring._extract_guide_sequences.__module__ == "crispr_hawk_b200.scoring"
        _, _, got = run_driver(drv, case)
        scored = case.pam != "NNGRRT"  # SaCas9 has no scorer (scoring.py:845-857): nothing is precomputed for it
        assert (got.hawk.get("kmers") is not None) == scored and got.built == 0
        assert scoring._extract_guide_sequences(got) == want4
        assert scoring._extract_guide_sequences_sgdesigner(got) == want0
        assert (got.built == 0) == scored  # device columns: no Guide object needed; else the reference's own loop
        plain = list(got)  # not a list of this package: the reference's loop over the objects
        assert scoring._extract_guide_sequences(plain) == want4
    finally:
        hawk.uninstall()
    assert scoring._extract_guide_sequences.__module__ == "crisprhawk.scoring"
