"""Seeded synthetic workloads of the BASELINE.json shapes.

A *cohort* is a random ACGT reference plus variant sites (SNV / insertion /
deletion) and, per haplotype, the sorted list of sites it carries -- what a
phased VCF of that shape would say. From it this module derives, with numpy on
the host, everything the scan needs besides the haplotype texts: lengths, slot
layout, run-length position maps and scan bounds, following the reference's
conventions (haplotype.py:90-159,185-252; search_guides.py:49-84). The texts
themselves are materialised on the device (`hawk_materialize_dev`) for the big
configurations, or with numpy (`materialize_host`) for test-sized ones.

Sites are spaced so no two edits of a haplotype overlap and none touches the BED
boundaries, hence every haplotype is one the reference's
own builder would accept.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import marshal

PADDING = marshal.PADDING
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class Cohort:
    ref: np.ndarray  # uint8 ASCII, padded region [bed_start - 100, bed_stop + 100]
    region_start: int  # genomic coordinate of ref[0] (padded start)
    region_stop: int  # padded stop
    bed_start: int
    bed_stop: int
    site_pos: np.ndarray  # int32, reference index of the anchor base, ascending
    site_reflen: np.ndarray  # int32
    site_altlen: np.ndarray  # int32
    site_altoff: np.ndarray  # int64 into alt_pool
    alt_pool: np.ndarray  # uint8 upper-case ASCII
    hap_off: np.ndarray  # int64 n_hap + 1 (CSR over haplotypes; haplotype 0 = REF, no edits)
    hap_sites: np.ndarray  # int32 site indices, ascending within a haplotype
    seed: int = 0
    _derived: dict = field(default_factory=dict)

    @property
    def n_hap(self) -> int:
        return len(self.hap_off) - 1


def make_cohort(bed_len: int, n_alt_hap: int, n_sites: int, mean_alts_per_hap: float, seed: int,
                snv_frac: float = 0.9, ins_frac: float = 0.05, max_indel: int = 10,
                bed_start: int = 10001, hap_block: int = 0) -> Cohort:  # fmt: skip
    """Region of `bed_len` bases, `n_alt_hap` non-reference haplotypes (+ REF as haplotype 0),
    `n_sites` cohort-wide sites on a jittered grid, allele frequencies from a 1/x spectrum
    scaled so a haplotype carries `mean_alts_per_hap` alternate alleles on average.
    `hap_block` selects an independent block of haplotypes over the SAME reference and
    sites (rank r of a multi-GPU run scans block r of the cohort)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    L = bed_len + 2 * PADDING
    ref = _BASES[rng.integers(0, 4, L)]
    region_start = bed_start - PADDING
    bed_stop = bed_start + bed_len - 1
    # jittered grid keeps >= max_indel + 2 bases between anchors; stay clear of the ends
    margin = PADDING + max_indel + 14
    usable = L - 2 * margin
    n_sites = int(min(n_sites, usable // (max_indel + 3)))
    if n_sites > 0:
        pitch = usable / n_sites
        jitter = rng.random(n_sites) * max(pitch - (max_indel + 2), 0.0)
        site_pos = (margin + np.arange(n_sites) * pitch + jitter).astype(np.int32)
    else:
        site_pos = np.zeros(0, np.int32)
    kind = rng.random(n_sites)
    k = rng.integers(1, max_indel + 1, n_sites).astype(np.int32)
    is_snv = kind < snv_frac
    is_ins = (~is_snv) & (kind < snv_frac + ins_frac)
    is_del = ~(is_snv | is_ins)
    reflen = np.where(is_del, k + 1, 1).astype(np.int32)
    altlen = np.where(is_ins, k + 1, 1).astype(np.int32)
    altoff = np.zeros(n_sites + 1, np.int64)
    np.cumsum(altlen, out=altoff[1:])
    pool = _BASES[rng.integers(0, 4, int(altoff[-1]))]
    anchors = ref[site_pos] if n_sites else np.zeros(0, np.uint8)
    # first ALT character: the anchor base for indels, a different base for SNVs
    first = np.where(
        is_snv, _BASES[(np.searchsorted(_BASES, anchors) + rng.integers(1, 4, n_sites)) % 4], anchors
    )
    if n_sites:
        pool[altoff[:-1]] = first
    # allele frequencies ~ 1/x on [1e-3, 0.5], scaled to the requested mean carrier count
    af = np.exp(rng.uniform(np.log(1e-3), np.log(0.5), n_sites))
    if n_sites and mean_alts_per_hap > 0:
        af *= mean_alts_per_hap / af.sum()
    af = np.clip(af, 0.0, 0.9)
    hap_off = np.zeros(n_alt_hap + 2, np.int64)
    chunks: List[np.ndarray] = []
    if hap_block:
        rng = np.random.Generator(np.random.PCG64([seed, hap_block]))
    for h in range(n_alt_hap):  # row-wise Bernoulli draws keep memory bounded
        carried = np.flatnonzero(rng.random(n_sites) < af).astype(np.int32)
        chunks.append(carried)
        hap_off[h + 2] = hap_off[h + 1] + len(carried)
    hap_sites = np.concatenate(chunks) if chunks else np.zeros(0, np.int32)
    return Cohort(ref, region_start, region_start + L - 1, bed_start, bed_stop, site_pos, reflen,
                  altlen, altoff[:-1].copy(), pool, hap_off, hap_sites.astype(np.int32), seed)  # fmt: skip


@dataclass
class Derived:
    """Everything about the haplotypes except their text (host numpy arrays)."""

    lens: np.ndarray  # int32
    slot_off: np.ndarray  # int64 n_hap + 1
    total_slots: int
    edit_outpos: np.ndarray  # int32 per edit (CSR = cohort.hap_off)
    seg: marshal.SegmentTable
    is_ref: np.ndarray  # uint8
    variant_bases: int  # lower-case bases over all haplotypes

    def n_hap_total(self) -> int:
        return len(self.lens)


def derive(c: Cohort) -> Derived:
    if "d" in c._derived:
        return c._derived["d"]
    n_hap = c.n_hap
    sites = c.hap_sites
    delta = (c.site_altlen - c.site_reflen)[sites].astype(np.int64)
    csum = np.concatenate(([0], np.cumsum(delta)))
    start_sum = csum[c.hap_off[:-1]]  # cumulative delta before each haplotype's first edit
    per_edit_hap = np.repeat(np.arange(n_hap), np.diff(c.hap_off))
    shift_before = csum[:-1] - start_sum[per_edit_hap]  # length change from earlier edits
    outpos = (c.site_pos[sites].astype(np.int64) + shift_before).astype(np.int32)
    lens = (len(c.ref) + (csum[c.hap_off[1:]] - csum[c.hap_off[:-1]])).astype(np.int32)
    slot_off, total = marshal.layout(lens)
    # run-length posmap: a deletion of k makes the base after the anchor jump by k; an
    # insertion of k makes k bases repeat the anchor's coordinate (step 0) and then resume
    altlen = c.site_altlen[sites]
    reflen = c.site_reflen[sites]
    anchor_gen = c.region_start + c.site_pos[sites].astype(np.int64)
    is_ins = altlen > 1
    is_del = reflen > 1
    # every haplotype starts with (0, region_start, step 1)
    n_seg_edit = is_ins.astype(np.int64) * 2 + is_del.astype(np.int64)
    seg_counts = np.ones(n_hap, np.int64)
    np.add.at(seg_counts, per_edit_hap, n_seg_edit)
    seg_off = np.concatenate(([0], np.cumsum(seg_counts)))
    n_seg = int(seg_off[-1])
    seg_rel = np.zeros(n_seg, np.int32)
    seg_gen = np.zeros(n_seg, np.int32)
    seg_step = np.ones(n_seg, np.uint8)
    seg_gen[seg_off[:-1]] = c.region_start
    # position of each edit's first segment inside its haplotype's segment list
    e_csum = np.concatenate(([0], np.cumsum(n_seg_edit)))
    e_first = seg_off[per_edit_hap] + 1 + (e_csum[:-1] - e_csum[c.hap_off[:-1]][per_edit_hap])
    ins = np.flatnonzero(is_ins)
    seg_rel[e_first[ins]] = outpos[ins] + 1
    seg_gen[e_first[ins]] = anchor_gen[ins]
    seg_step[e_first[ins]] = 0
    seg_rel[e_first[ins] + 1] = outpos[ins] + altlen[ins]
    seg_gen[e_first[ins] + 1] = anchor_gen[ins] + 1
    dl = np.flatnonzero(is_del)
    seg_rel[e_first[dl]] = outpos[dl] + 1
    seg_gen[e_first[dl]] = anchor_gen[dl] + reflen[dl]
    seg = marshal.SegmentTable(seg_off.astype(np.int64), seg_rel, seg_gen, seg_step)
    is_ref = (np.diff(c.hap_off) == 0).astype(np.uint8)
    d = Derived(lens, slot_off, total, outpos, seg, is_ref, int(altlen.sum()))
    c._derived["d"] = d
    return d


def scan_bounds(c: Cohort, pamlen: int):
    """compute_scan_start_stop (search_guides.py:49-84) for every haplotype: the BED
    boundaries are never touched by an edit, so posmap_rev is the shifted index."""
    d = derive(c)
    sites = c.hap_sites
    pos = c.site_pos[sites].astype(np.int64)
    delta = (c.site_altlen - c.site_reflen)[sites].astype(np.int64)
    per_edit_hap = np.repeat(np.arange(c.n_hap), np.diff(c.hap_off))

    def rel_index(g_rel: int) -> np.ndarray:
        shift = np.zeros(c.n_hap, np.int64)
        np.add.at(shift, per_edit_hap, np.where(pos < g_rel, delta, 0))
        return g_rel + shift

    a = rel_index(c.bed_start - c.region_start)
    b = rel_index(c.bed_stop - c.region_start) - pamlen + 1
    return a.astype(np.int32), b.astype(np.int32)


def materialize_host(c: Cohort) -> List[str]:
    """Haplotype texts with numpy/Python (small cohorts only)."""
    ref = c.ref.tobytes().decode()
    out = []
    for h in range(c.n_hap):
        parts, cur = [], 0
        for s in c.hap_sites[c.hap_off[h] : c.hap_off[h + 1]]:
            p = int(c.site_pos[s])
            parts.append(ref[cur:p])
            o = int(c.site_altoff[s])
            parts.append(c.alt_pool[o : o + int(c.site_altlen[s])].tobytes().decode().lower())
            cur = p + int(c.site_reflen[s])
        parts.append(ref[cur:])
        out.append("".join(parts))
    return out


def to_vcf_lines(c: Cohort, contig: str = "chr1"):
    """Phased VCF data lines + sample names describing the cohort (for feeding the live
    reference in tests): ALT haplotypes 2i+1, 2i+2 are the two copies of sample i."""
    n_alt = c.n_hap - 1
    n_samples = (n_alt + 1) // 2
    samples = [f"S{i + 1}" for i in range(n_samples)]
    carriers = [set() for _ in range(len(c.site_pos))]
    for h in range(1, c.n_hap):
        for s in c.hap_sites[c.hap_off[h] : c.hap_off[h + 1]]:
            carriers[int(s)].add(h - 1)
    ref = c.ref.tobytes().decode()
    lines = []
    for s in range(len(c.site_pos)):
        if not carriers[s]:
            continue
        p = int(c.site_pos[s])
        o = int(c.site_altoff[s])
        alt = c.alt_pool[o : o + int(c.site_altlen[s])].tobytes().decode()
        gts = []
        for i in range(n_samples):
            gts.append(f"{int(2 * i in carriers[s])}|{int(2 * i + 1 in carriers[s])}")
        lines.append("\t".join([contig, str(c.region_start + p), ".", ref[p : p + int(c.site_reflen[s])],
                                alt, ".", "PASS", "AF=0.1", "GT"] + gts))  # fmt: skip
    return lines, samples


class SynthHap:
    """Duck-typed haplotype (haplotype.py:23-77) over a cohort member, for the oracle and
    for `crispr_hawk_b200.search` in tests."""

    class _Seq:
        def __init__(self, text):
            self.sequence = text

        def __len__(self):
            return len(self.sequence)

    def __init__(self, c: Cohort, h: int, text: str, posmap_vals: np.ndarray):
        self.sequence = SynthHap._Seq(text)
        vals = posmap_vals.tolist()
        self.posmap = dict(enumerate(vals))
        self.posmap_rev = {g: i for i, g in enumerate(vals)}
        self.start, self.stop = c.region_start, c.region_stop
        n = int(c.hap_off[h + 1] - c.hap_off[h])
        self.samples = "REF" if n == 0 else f"S{(h + 1) // 2}:{'1|0' if h % 2 else '0|1'}"
        self.variants = "NA" if n == 0 else ",".join(f"v{int(s)}" for s in c.hap_sites[c.hap_off[h] : c.hap_off[h + 1]])
        self.afs = {}
        self.variant_alleles = {}
        self.id = f"hap{h}"

    def __len__(self):
        return len(self.sequence)


def synth_haplotypes(c: Cohort, texts: Optional[List[str]] = None) -> List[SynthHap]:
    d = derive(c)
    texts = texts or materialize_host(c)
    out = []
    for h in range(c.n_hap):
        s0, s1 = int(d.seg.seg_off[h]), int(d.seg.seg_off[h + 1])
        vals = marshal.eval_segments(d.seg.seg_rel[s0:s1], d.seg.seg_gen[s0:s1], d.seg.seg_step[s0:s1],
                                     np.arange(int(d.lens[h])))  # fmt: skip
        out.append(SynthHap(c, h, texts[h], vals))
    return out


class SynthRegion:
    def __init__(self, c: Cohort, contig: str = "chr1"):
        self.contig, self.start, self.stop = contig, c.region_start, c.region_stop
        self.coordinates = f"{contig}:{c.region_start}-{c.region_stop}"


# ---------------------------------------------------------------- device materialisation
def materialize_device(c: Cohort, ctx=None, device=None):
    """Haplotype texts in the slot layout as a torch uint8 CUDA tensor (total_slots bytes)."""
    import ctypes as C

    import torch

    from . import _cabi

    lib = _cabi.load_library()
    d = derive(c)
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    sites = c.hap_sites
    g_ref = t(np.concatenate((c.ref, np.zeros(32, np.uint8))))  # block copies read whole aligned words
    g_off = t(c.hap_off)
    g_pos = t(c.site_pos[sites])
    g_rl = t(c.site_reflen[sites])
    g_al = t(c.site_altlen[sites])
    g_ao = t(c.site_altoff[sites])
    g_op = t(d.edit_outpos)
    g_pool = t(c.alt_pool if len(c.alt_pool) else np.zeros(1, np.uint8))
    g_slot = t(d.slot_off)
    g_len = t(d.lens)
    out = torch.empty(d.total_slots, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    p = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
    _cabi.check(
        lib.hawk_materialize_dev(
            C.c_void_p(stream), p(g_ref), len(c.ref), p(g_off), p(g_pos), p(g_rl), p(g_al), p(g_ao),
            p(g_op), p(g_pool), p(g_slot), p(g_len), c.n_hap, d.total_slots, len(sites), int(d.lens.max()), p(out),
        ),  # fmt: skip
        "hawk_materialize_dev",
    )
    torch.cuda.synchronize(dev)
    return out


CONFIGS = {
    # BASELINE.md section 4
    "c1": dict(bed_len=5_000, n_alt_hap=20, n_sites=40, mean_alts=8, seed=1, snv=0.70, ins=0.15,
               max_indel=5, pam="NGG", guidelen=20, right=False),
    "c2": dict(bed_len=1_000_000, n_alt_hap=5008, n_sites=10_000, mean_alts=1000, seed=2, snv=0.90,
               ins=0.05, max_indel=10, pam="NGG", guidelen=20, right=False),
    "c3": dict(bed_len=1_000_000, n_alt_hap=5008, n_sites=10_000, mean_alts=1000, seed=2, snv=0.90,
               ins=0.05, max_indel=10, pam="TTTV", guidelen=23, right=True),
    # unphased (gnomAD-style pseudo-samples): generated by synth_unphased.make_unphased_cohort
    "c4": dict(bed_len=10_000_000, unphased=True, pitch=8, seed=4, snv=0.88, multi=0.05, max_indel=5, n_samples=10,
               pam="NNGRRT", guidelen=21, right=False),
    "c5shard": dict(bed_len=50_000_000, n_alt_hap=625, n_sites=500_000, mean_alts=50_000, seed=5,
                    snv=0.90, ins=0.05, max_indel=10, pam="NGG", guidelen=20, right=False),
}  # fmt: skip


def config_cohort(name: str, scale: float = 1.0, seed_offset: int = 0, n_alt_hap: Optional[int] = None,
                  hap_block: int = 0) -> Cohort:
    k = CONFIGS[name]
    if k.get("unphased"):
        from . import synth_unphased

        return synth_unphased.make_unphased_cohort(max(200, int(k["bed_len"] * scale)), k["seed"] + seed_offset, k["pitch"],
                                                   k["snv"], k["multi"], k["max_indel"], k["n_samples"])  # fmt: skip
    bed_len = max(200, int(k["bed_len"] * scale))
    n_sites = max(1, int(k["n_sites"] * scale))
    mean = k["mean_alts"] * scale
    return make_cohort(bed_len, k["n_alt_hap"] if n_alt_hap is None else n_alt_hap, n_sites, mean,
                       k["seed"] + seed_offset, k["snv"], k["ins"], k["max_indel"], hap_block=hap_block)  # fmt: skip
