"""N1: haplotypes materialised on the device from edit lists (crispr_hawk_b200/haplotypes.py,
hawk_batch_create_from_edits) against the reference's own builder (live, CPU) and against the
text path (GPU)."""

import numpy as np
import pytest

from crispr_hawk_b200 import haplotypes as HN
from crispr_hawk_b200 import marshal, synth
from oracle import refshim


def cohort_edits(c):
    ref = c.ref.tobytes().decode()
    out = []
    for h in range(c.n_hap):
        edits = []
        for s in c.hap_sites[c.hap_off[h] : c.hap_off[h + 1]]:
            p = int(c.site_pos[s])
            o = int(c.site_altoff[s])
            alt = c.alt_pool[o : o + int(c.site_altlen[s])].tobytes().decode()
            edits.append(HN.Edit(c.region_start + p, ref[p : p + int(c.site_reflen[s])], alt))
        out.append(edits)
    return ref, out


def small(seed=31, n_alt=7):
    return synth.make_cohort(bed_len=2500, n_alt_hap=n_alt, n_sites=110, mean_alts_per_hap=25, seed=seed,
                             snv_frac=0.55, ins_frac=0.25, max_indel=7)  # fmt: skip


def test_segment_map_matches_dict_semantics():
    c = small()
    d = synth.derive(c)
    haps = synth.synth_haplotypes(c)
    for h, hap in enumerate(haps):
        s0, s1 = int(d.seg.seg_off[h]), int(d.seg.seg_off[h + 1])
        pm = HN.SegmentMap(d.seg.seg_rel[s0:s1].astype(np.int64), d.seg.seg_gen[s0:s1].astype(np.int64),
                           d.seg.seg_step[s0:s1], int(d.lens[h]))  # fmt: skip
        assert [pm[i] for i in range(0, len(pm), 37)] == [hap.posmap[i] for i in range(0, len(pm), 37)]
        rev = hap.posmap_rev
        for g in list(range(c.region_start, c.region_stop + 1, 53)) + [c.bed_start, c.bed_stop]:
            assert pm.last_index_of(g) == rev.get(g)


@pytest.mark.ref
def test_edit_semantics_equal_reference_builder():
    """What the device is asked to build (texts via the same edit rules on the host, segments,
    scan bounds) equals what the reference's add_variants_phased builds from the equivalent VCF."""
    ref = refshim.load()
    c = small(seed=44, n_alt=8)
    lines, samples = synth.to_vcf_lines(c)
    region, ref_haps = refshim.build_case(c.ref.tobytes().decode(), c.bed_start, c.bed_stop, lines, samples, True)
    by_text = {h.sequence.sequence: h for h in ref_haps}
    texts = synth.materialize_host(c)
    d = synth.derive(c)
    for h, t in enumerate(texts):
        rh = by_text[t]  # the reference collapses identical haplotypes; every one of ours must exist
        s0, s1 = int(d.seg.seg_off[h]), int(d.seg.seg_off[h + 1])
        pm = HN.SegmentMap(d.seg.seg_rel[s0:s1].astype(np.int64), d.seg.seg_gen[s0:s1].astype(np.int64),
                           d.seg.seg_step[s0:s1], len(t))  # fmt: skip
        assert pm.values().tolist() == [rh.posmap[i] for i in range(len(rh))]
        eh = HN.EditHaplotype(None, h, len(t), pm, region.start, region.stop, "x", "x", {}, "x")
        assert eh.scan_bounds(region.start, region.stop, 3) == tuple(
            ref.search_guides.compute_scan_start_stop(rh, region.start, region.stop, 3)
        )


@pytest.mark.gpu
@pytest.mark.parametrize("pam,G,right", [("NGG", 20, False), ("TTTV", 23, True)])
def test_device_materialised_search_equals_text_search(pam, G, right):
    import crispr_hawk_b200 as hawk
    from oracle import hawk_oracle as O

    c = small(seed=52, n_alt=9)
    ref_text, edits = cohort_edits(c)
    haps, packed = hawk.build_phased(ref_text, c.region_start, edits)
    texts = synth.materialize_host(c)
    assert [h.text() for h in haps] == texts
    region = synth.SynthRegion(c)
    got = hawk.search(pam, region, haps, packed, G, right, True, True, 0, True)
    # the same search from host texts (the drop-in path) and from the oracle
    thaps = synth.synth_haplotypes(c, texts)
    for th, eh in zip(thaps, haps):
        th.samples, th.variants, th.id = eh.samples, eh.variants, eh.id
    want = hawk.search(pam, region, thaps, None, G, right, True, True, 0, True)
    key = lambda g: (g.start, g.stop, g.strand, g.sequence, g.samples, g.variants, g.hapid, tuple(sorted(g.posmap.items())))  # noqa: E731
    assert [key(g) for g in got] == [key(g) for g in want]
    ora = O.search(pam, c.region_start, c.region_stop, [O.OracleHap.from_object(h) for h in thaps], G, right, True, True)
    assert [(g.start, g.stop, g.strand, g.sequence) for g in got] == [(g.start, g.stop, g.strand, g.sequence) for g in ora]
    assert len(got) > 50


@pytest.mark.gpu
def test_bad_edits_are_rejected():
    import crispr_hawk_b200 as hawk
    from crispr_hawk_b200 import _cabi

    ref = "ACGT" * 100
    with pytest.raises(ValueError):  # REF allele does not match the reference text
        hawk.build_phased(ref, 1000, [[], [HN.Edit(1010, "T", "A")]])
    with pytest.raises(_cabi.HawkLibraryError):  # overlapping edits
        hawk.build_phased(ref, 1000, [[HN.Edit(1008, "ACGT", "A"), HN.Edit(1010, "G", "T")]])


# --------------------------------------------------------------------------- plan on arrays
def random_records(seed, n_samples=9, n_sites=120, L=4000, dup=True):
    import types

    rng = np.random.default_rng(seed)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, L))
    start = 7000
    names = [f"S{i}" for i in range(n_samples)]
    pos = np.sort(rng.choice(np.arange(20, L - 40, 12), n_sites, replace=False))
    recs = []
    for p in pos.tolist():
        kind = rng.random()
        if kind < 0.6:
            r, a = ref[p], "ACGT"[("ACGT".index(ref[p]) + int(rng.integers(1, 4))) % 4]
        elif kind < 0.8:
            r, a = ref[p], ref[p] + "".join("ACGT"[i] for i in rng.integers(0, 4, int(rng.integers(1, 6))))
        else:
            r, a = ref[p : p + int(rng.integers(2, 7))], ref[p]
        c0 = {n for n in names if rng.random() < 0.3}
        c1 = {n for n in names if rng.random() < 0.3} if rng.random() < 0.7 else set(c0)
        vt = "snp" if len(r) == 1 and len(a) == 1 else "indel"
        recs.append(types.SimpleNamespace(position=start + p, ref=r, alt=[a], afs=[round(float(rng.random()), 3)],
                                          samples=[(c0, c1)], vtype=[vt], id=[f"chr1-{start + p}-{r}/{a}"]))  # fmt: skip
        if dup and rng.random() < 0.05:  # the same record twice (carriers differ)
            recs.append(types.SimpleNamespace(position=start + p, ref=r, alt=[a], afs=[0.5], samples=[({names[0]}, set())],
                                              vtype=[vt], id=[f"chr1-{start + p}-{r}/{a}"]))  # fmt: skip
    return ref, start, names, recs


@pytest.mark.parametrize("seed", range(8))
def test_plan_on_arrays_equals_plan_on_objects(seed):
    """plan_phased_arrays (one pass over the records, numpy per chromosome copy) against
    plan_phased (Python per variant and copy): same haplotypes, order, strings and dicts."""
    ref, start, names, recs = random_records(seed, dup=False)
    slow = HN.plan_phased(ref, start, names, recs)
    T, fast = HN.plan_phased_arrays(ref, start, names, recs)
    assert len(slow) == len(fast) > 3
    for a, b in zip(slow, fast):
        e = b["e_idx"]
        assert [(x.pos, len(x.ref), len(x.alt)) for x in a["edits"]] == list(zip(T.pos[e].tolist(), T.reflen[e].tolist(), T.altlen[e].tolist()))
        assert [x.alt.upper() for x in a["edits"]] == [bytes(T.pool[o : o + n]).decode() for o, n in zip(T.altoff[e].tolist(), T.altlen[e].tolist())]
        assert (a["samples"], a["variants"], a["afs"]) == (b["samples"], b["variants"], b["afs"])
        assert list(a["afs"]) == list(b["afs"])  # insertion order too


def test_unusual_inputs_leave_the_array_plan():
    import types

    ref, start, names, recs = random_records(3)
    with pytest.raises(HN._Unusual):  # duplicate records on one copy overlap: plan_phased's business
        HN.plan_phased_arrays(ref, start, names, recs + [types.SimpleNamespace(
            position=recs[0].position, ref=recs[0].ref, alt=recs[0].alt, afs=[0.1], samples=recs[0].samples, vtype=recs[0].vtype, id=recs[0].id)])
    ref2, start2, names2, recs2 = random_records(4, dup=False)
    bad = recs2[5]
    bad.ref = ("A" if bad.ref[0] != "A" else "C") + bad.ref[1:]
    with pytest.raises(HN._Unusual):  # REF allele does not match the reference text
        HN.plan_phased_arrays(ref2, start2, names2, recs2)
    with pytest.raises(ValueError):  # ... and plan_phased raises the reference's message
        HN.plan_phased(ref2, start2, names2, recs2)
    ref3, start3, names3, recs3 = random_records(5, dup=False)
    recs3[2].ref, recs3[2].alt = ref3[recs3[2].position - start3 : recs3[2].position - start3 + 2], ["TT"]
    with pytest.raises(HN._Unusual):  # complex substitution
        HN.plan_phased_arrays(ref3, start3, names3, recs3)
