"""Seeded synthetic workload of the BASELINE config-4 shape: an UNPHASED cohort.

gnomAD-style input (converter.py:19-37): ten population pseudo-samples whose genotypes
only say "allele observed in this population", hence unphased. For such input the
reference builds (haplotypes.py:669-714, add_variants_unphased)

  * one REF haplotype;
  * per sample one full-length haplotype carrying ALL its SNVs as lower-case IUPAC codes
    (haplotype.py:254-329; code = REF base + every ALT the sample carries), collapsed by
    sequence (haplotypes.py:274-294);
  * per indel inside the BED interval, per carrier sample, one <= 201-base *window*
    haplotype (haplotypes.py:535-669): the reference text 100 bases either side of the
    anchor, the carrier's overlapping SNVs as IUPAC codes, then the indel; collapsed by
    sequence inside the indel's group. One quirk is kept: in a deletion window the
    position map is initialised with the final (shorter) length (haplotype.py:90-104 with
    chains < 0), so SNVs in the last k positions of the window are silently skipped
    (haplotype.py:268-271, the KeyError branch).

This module derives exactly those haplotypes with numpy -- texts in the C-ABI slot
layout, run-length position maps, `variant_alleles` tables, scan bounds -- at sizes the
reference's O(length) dict rewrites per variant cannot reach. tests/test_synth_unphased.py
checks it haplotype by haplotype against the reference's own builder on small regions.
Sites keep >= max_indel + 1 bases between anchors, indels stay clear of the BED
boundaries and SNV sites never sit on an indel anchor.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import marshal

PADDING = marshal.PADDING
WINDOW_FLANK = 100  # haplotypes.py:548-549
GNOMAD_POPS = ["afr", "ami", "amr", "asj", "eas", "fin", "nfe", "mid", "sas", "remaining"]  # converter.py:19-30
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_LETTER_OF_MASK = np.frombuffer(b"?ACMGRSVTWYHKDBN", dtype=np.uint8)  # nibble -> IUPAC letter
_NIB_OF_ASCII = np.zeros(256, np.uint8)
for _ch, _n in marshal._NIBBLE.items():
    _NIB_OF_ASCII[ord(_ch)] = _n
    _NIB_OF_ASCII[ord(_ch.lower())] = _n

SNV, INS, DEL = 0, 1, 2


@dataclass
class UnphasedCohort:
    ref: np.ndarray  # uint8 ASCII, padded region
    region_start: int
    region_stop: int
    bed_start: int
    bed_stop: int
    samples: List[str]
    site_pos: np.ndarray  # int32 reference index of the site / anchor base, ascending
    site_kind: np.ndarray  # uint8 SNV / INS / DEL
    site_k: np.ndarray  # int32 indel length (0 for SNVs)
    snv_alt: np.ndarray  # uint8 (n_sites, 2): one-hot nibble of ALT 1 / ALT 2 (0 = no such allele)
    ins_off: np.ndarray  # int64 (n_sites + 1) into ins_pool: the inserted bases (after the anchor)
    ins_pool: np.ndarray  # uint8 upper-case ASCII
    carry: np.ndarray  # uint16 (n_sites, 2): samples carrying ALT 1 / ALT 2 (bit s = samples[s])
    seed: int = 0
    _derived: dict = field(default_factory=dict)


def make_unphased_cohort(bed_len: int, seed: int, pitch: int = 8, snv_frac: float = 0.88, multi_frac: float = 0.05,
                         max_indel: int = 5, n_samples: int = 10, bed_start: int = 10001,
                         carrier_probs=(0.1, 0.3, 0.6), carrier_weights=(0.45, 0.35, 0.20)) -> UnphasedCohort:  # fmt: skip
    """One site every `pitch` bases on a jittered grid (1 / 8 bp, 88 % SNV,
    12 % indel of 1..max_indel bases, 5 % of the SNVs multi-allelic); every (site, population)
    pair is a carrier with the site's own probability."""
    if pitch < max_indel + 2:
        raise ValueError("pitch must leave room for the longest deletion between anchors")
    rng = np.random.Generator(np.random.PCG64(seed))
    L = bed_len + 2 * PADDING
    ref = _BASES[rng.integers(0, 4, L)]
    margin = 24
    n_sites = max(0, (L - 2 * margin - max_indel - 3) // pitch)
    jitter = rng.integers(0, pitch - max_indel - 1 + 1, n_sites) if n_sites else np.zeros(0, np.int64)
    pos = (margin + pitch * np.arange(n_sites) + jitter).astype(np.int32)
    r = rng.random(n_sites)
    kind = np.where(r < snv_frac, SNV, np.where(r < snv_frac + (1 - snv_frac) / 2, INS, DEL)).astype(np.uint8)
    # indels only well inside the BED interval (the reference ignores those of the padding, and
    # an indel across a BED boundary would move the scan bounds)
    inside = (pos >= PADDING + pitch) & (pos <= L - PADDING - 1 - max_indel - pitch)
    kind[~inside] = SNV
    k = np.where(kind == SNV, 0, rng.integers(1, max_indel + 1, n_sites)).astype(np.int32)
    refn = np.searchsorted(_BASES, ref[pos]) if n_sites else np.zeros(0, np.int64)
    d1 = rng.integers(1, 4, n_sites)
    d2 = 1 + (d1 - 1 + rng.integers(1, 3, n_sites)) % 3  # another base, different from ALT 1 as well
    multi = (kind == SNV) & (rng.random(n_sites) < multi_frac)
    snv_alt = np.zeros((n_sites, 2), np.uint8)
    snv_alt[:, 0] = np.where(kind == SNV, 1 << ((refn + d1) % 4), 0)
    snv_alt[:, 1] = np.where(multi, 1 << ((refn + d2) % 4), 0)
    ins_len = np.where(kind == INS, k, 0).astype(np.int64)
    ins_off = np.concatenate(([0], np.cumsum(ins_len)))
    ins_pool = _BASES[rng.integers(0, 4, int(ins_off[-1]))]
    p_site = rng.choice(np.asarray(carrier_probs), size=n_sites, p=np.asarray(carrier_weights))
    draw = rng.random((n_sites, 2, n_samples))
    bits = (1 << np.arange(n_samples)).astype(np.uint16)
    c1 = ((draw[:, 0, :] < p_site[:, None]) * bits).sum(axis=1).astype(np.uint16)
    c2 = ((draw[:, 1, :] < 0.5 * p_site[:, None]) * bits).sum(axis=1).astype(np.uint16)
    # every site is observed somewhere (a VCF line without carriers would not exist)
    lonely = c1 == 0
    c1[lonely] = bits[rng.integers(0, n_samples, int(lonely.sum()))]
    c2[~multi] = 0
    carry = np.stack([c1, c2], axis=1)
    names = GNOMAD_POPS[:n_samples] if n_samples <= len(GNOMAD_POPS) else [f"P{i}" for i in range(n_samples)]
    return UnphasedCohort(ref, bed_start - PADDING, bed_start - PADDING + L - 1, bed_start, bed_start + bed_len - 1,
                          list(names), pos, kind, k, snv_alt, ins_off, ins_pool, carry, seed)  # fmt: skip


def to_vcf_lines(c: UnphasedCohort, contig: str = "chr1"):
    """Unphased VCF data lines (GT 0/0, 0/1, 0/2, 1/2) describing the cohort, for the live
    reference in tests. AF is a per-allele constant: frequencies play no role on this path."""
    ref = c.ref.tobytes().decode()
    lines = []
    ns = len(c.samples)
    for j in range(len(c.site_pos)):
        p, kd, k = int(c.site_pos[j]), int(c.site_kind[j]), int(c.site_k[j])
        if kd == SNV:
            alts = [chr(_LETTER_OF_MASK[int(a)]) for a in c.snv_alt[j] if a]
            refa = ref[p]
        elif kd == INS:
            o = int(c.ins_off[j])
            alts = [ref[p] + c.ins_pool[o : o + k].tobytes().decode()]
            refa = ref[p]
        else:
            alts = [ref[p]]
            refa = ref[p : p + k + 1]
        gts = []
        for s in range(ns):
            a1, a2 = (int(c.carry[j, 0]) >> s) & 1, (int(c.carry[j, 1]) >> s) & 1
            gts.append("1/2" if a1 and a2 else "0/1" if a1 else "0/2" if a2 else "0/0")
        lines.append("\t".join([contig, str(c.region_start + p), ".", refa, ",".join(alts), ".", "PASS",
                                "AF=" + ",".join("0.1" for _ in alts), "GT"] + gts))  # fmt: skip
    return lines, list(c.samples)


@dataclass
class UnphasedSet:
    """Every haplotype of the cohort in the flat form the C-ABI (and the C oracle) take."""

    ascii: np.ndarray  # uint8 slot space
    slot_off: np.ndarray  # int64 n_hap + 1
    lens: np.ndarray  # int32
    total_slots: int
    seg: marshal.SegmentTable
    alleles: marshal.AlleleTable
    is_ref: np.ndarray  # uint8
    hap_start: np.ndarray  # int32 genomic bounds of the haplotype (region or window)
    hap_stop: np.ndarray
    hap_kind: np.ndarray  # uint8 0 REF, 1 full-length SNV haplotype, 2 indel window
    hap_site: np.ndarray  # int32 indel site of a window haplotype, else -1
    hap_samples: np.ndarray  # uint16 bitmask of the samples collapsed into the haplotype
    hap_rep: np.ndarray  # int32 first of them (its variant list is the haplotype's)
    anchor_rel: np.ndarray  # int32 window haplotypes: relative index of the anchor, else -1
    dlen: np.ndarray  # int32 window haplotypes: ALT length - REF length, else 0
    variant_bases: int

    @property
    def n_hap(self) -> int:
        return len(self.lens)

    def n_hap_total(self) -> int:
        return len(self.lens)

    def take(self, idx) -> "UnphasedSet":
        """The haplotypes `idx` (any order) as a set of their own: what a search over exactly
        those haplotypes is fed (oracle subsets of a full-size run, the reference's ordering)."""
        idx = np.asarray(idx, np.int64)
        lens = self.lens[idx]
        slot_off, total = marshal.layout(lens)
        buf = np.zeros(total, np.uint8)
        for k, h in enumerate(idx.tolist()):
            o, n = int(self.slot_off[h]), int(lens[k])
            buf[slot_off[k] : slot_off[k] + n] = self.ascii[o : o + n]
        so = self.seg.seg_off
        own, sk = _ragged_ranges(so[idx], so[idx + 1])
        seg = marshal.SegmentTable(np.concatenate(([0], np.cumsum(so[idx + 1] - so[idx]))).astype(np.int64),
                                   self.seg.seg_rel[sk], self.seg.seg_gen[sk], self.seg.seg_step[sk])  # fmt: skip
        A = self.alleles
        own, vj = _ragged_ranges(A.va_off[idx], A.va_off[idx + 1])
        _, ve = _ragged_ranges(A.va_ent_off[vj], A.va_ent_off[vj + 1])
        n_ent = (A.va_ent_off[vj + 1] - A.va_ent_off[vj]).astype(np.int64)
        alleles = marshal.AlleleTable(np.concatenate(([0], np.cumsum(A.va_off[idx + 1] - A.va_off[idx]))).astype(np.int64),
                                      A.va_idx[vj], np.concatenate(([0], np.cumsum(n_ent))).astype(np.int64), A.va_ref[ve])  # fmt: skip
        return UnphasedSet(buf, slot_off, lens, total, seg, alleles, self.is_ref[idx], self.hap_start[idx],
                           self.hap_stop[idx], self.hap_kind[idx], self.hap_site[idx], self.hap_samples[idx],
                           self.hap_rep[idx], self.anchor_rel[idx], self.dlen[idx],
                           int(((buf >= ord("a")) & (buf <= ord("z"))).sum()))  # fmt: skip

    def scan_bounds(self, c: UnphasedCohort, pamlen: int):
        """compute_scan_start_stop (search_guides.py:49-84) from the window geometry: no edit
        touches a BED boundary, and an indel only shifts the indices behind its anchor."""
        g_lo = np.maximum(c.bed_start, self.hap_start).astype(np.int64)
        g_hi = np.minimum(c.bed_stop, self.hap_stop).astype(np.int64)
        a = g_lo - self.hap_start
        b = g_hi - self.hap_start + self.dlen - pamlen + 1
        return a.astype(np.int32), b.astype(np.int32)


def _snv_texts(c: UnphasedCohort):
    """Per sample: the reference with every SNV site the sample carries replaced by the
    lower-case IUPAC code of REF + carried ALTs (haplotype.py:287-297)."""
    ns = len(c.samples)
    L = len(c.ref)
    T = np.empty((ns, L), np.uint8)
    is_snv = c.site_kind == SNV
    refnib = _NIB_OF_ASCII[c.ref[c.site_pos]] if len(c.site_pos) else np.zeros(0, np.uint8)
    for s in range(ns):
        T[s] = c.ref
        a1 = ((c.carry[:, 0] >> s) & 1).astype(bool) & is_snv
        a2 = ((c.carry[:, 1] >> s) & 1).astype(bool) & is_snv
        any_ = a1 | a2
        mask = refnib | np.where(a1, c.snv_alt[:, 0], 0).astype(np.uint8) | np.where(a2, c.snv_alt[:, 1], 0).astype(np.uint8)
        T[s, c.site_pos[any_]] = _LETTER_OF_MASK[mask[any_]] | 0x20
    return T


def _ragged_ranges(lo: np.ndarray, hi: np.ndarray):
    """(owner, value) for value in [lo[i], hi[i]) for every i."""
    n = (hi - lo).clip(min=0).astype(np.int64)
    owner = np.repeat(np.arange(len(lo)), n)
    start = np.concatenate(([0], np.cumsum(n)))[:-1]
    val = np.arange(int(n.sum())) - np.repeat(start, n) + np.repeat(lo.astype(np.int64), n)
    return owner, val


_HASH_MUL = np.uint64(0x9E3779B97F4A7C15)


def derive_unphased(c: UnphasedCohort, row_block: int = 1 << 15) -> UnphasedSet:
    if "u" in c._derived:
        return c._derived["u"]
    ns = len(c.samples)
    L = len(c.ref)
    T = _snv_texts(c)
    is_snv = c.site_kind == SNV
    snv_carry = np.where(is_snv, c.carry[:, 0] | c.carry[:, 1], 0).astype(np.uint16)

    # ---- full-length SNV haplotypes, collapsed by sequence in sample order -------------------
    snv_rep: List[int] = []
    snv_mask: List[int] = []
    seen = {}
    for s in range(ns):
        if not ((snv_carry >> s) & 1).any():
            continue  # samples without variants are dropped (haplotypes.py:186)
        key = T[s].tobytes()
        if key in seen:
            snv_mask[seen[key]] |= 1 << s
        else:
            seen[key] = len(snv_rep)
            snv_rep.append(s)
            snv_mask.append(1 << s)
    del seen

    # ---- window haplotypes: (indel, carrier sample) rows -------------------------------------
    gpos = c.region_start + c.site_pos.astype(np.int64)
    indel = np.flatnonzero((c.site_kind != SNV) & (gpos >= c.bed_start) & (gpos < c.bed_stop) & (c.carry[:, 0] != 0))
    bits = (c.carry[indel, 0][:, None] >> np.arange(ns)[None, :]) & 1
    r_site_i, r_sample = np.nonzero(bits)  # row-major: site order, then sample order
    r_site = indel[r_site_i]
    p = c.site_pos[r_site].astype(np.int64)
    k = c.site_k[r_site].astype(np.int64)
    is_ins = c.site_kind[r_site] == INS
    dlen = np.where(is_ins, k, -k)
    w0 = np.maximum(0, p - WINDOW_FLANK)
    w1 = np.minimum(L - 1, p - dlen + WINDOW_FLANK)  # inclusive (haplotypes.py:549)
    a = p - w0
    rlen = (w1 - w0 + 1 + dlen).astype(np.int64)
    n_rows = len(r_site)
    lmax = int(rlen.max()) if n_rows else 0
    skip_from = np.where(is_ins, L, w1 - k + 1)  # deletion quirk: SNVs at ref index >= this are not applied
    texts = np.zeros((n_rows, lmax), np.uint8)
    cols = np.arange(lmax, dtype=np.int64)[None, :]
    flatT = T.reshape(-1)
    for r0 in range(0, n_rows, row_block):
        sl = slice(r0, min(n_rows, r0 + row_block))
        a_, k_, d_, w0_, ins_ = a[sl, None], k[sl, None], dlen[sl, None], w0[sl, None], is_ins[sl, None]
        src = np.where(cols <= a_, w0_ + cols, w0_ + cols - d_)
        src = np.clip(src, 0, L - 1)
        from_ref = src >= skip_from[sl, None]
        ch = np.where(from_ref, c.ref[src], flatT[r_sample[sl, None].astype(np.int64) * L + src])
        ch = np.where(cols == a_, c.ref[np.clip(w0_ + a_, 0, L - 1)] | 0x20, ch)  # the anchor: first ALT character
        in_ins = ins_ & (cols > a_) & (cols <= a_ + k_)
        io = np.clip(c.ins_off[r_site[sl]][:, None] + (cols - a_ - 1), 0, max(len(c.ins_pool) - 1, 0))
        if len(c.ins_pool):
            ch = np.where(in_ins, c.ins_pool[io] | 0x20, ch)
        ch = np.where(cols < rlen[sl, None], ch, 0)
        texts[sl] = ch
    # collapse identical texts inside each indel's group, first (lowest sample) wins
    if n_rows:
        pw = np.cumprod(np.full(lmax, _HASH_MUL, np.uint64))
        hsh = (texts.astype(np.uint64) * pw[None, :]).sum(axis=1, dtype=np.uint64)
        key = np.empty(n_rows, dtype=[("site", np.int64), ("h", np.uint64)])
        key["site"], key["h"] = r_site, hsh
        _, first, inv = np.unique(key, return_index=True, return_inverse=True)
        rep_row = first[inv]
        if not (texts == texts[rep_row]).all():
            raise RuntimeError("hash collision while collapsing window haplotypes")
        keep = rep_row == np.arange(n_rows)
        w_mask = np.zeros(n_rows, np.uint16)
        np.bitwise_or.at(w_mask, rep_row, (1 << r_sample).astype(np.uint16))
    else:
        keep = np.zeros(0, bool)
        w_mask = np.zeros(0, np.uint16)
    kr = np.flatnonzero(keep)
    n_win = len(kr)

    # ---- haplotype list: REF, SNV haplotypes, windows ----------------------------------------
    n_snv = len(snv_rep)
    n_hap = 1 + n_snv + n_win
    lens = np.concatenate(([L], np.full(n_snv, L), rlen[kr])).astype(np.int32)
    slot_off, total = marshal.layout(lens)
    buf = np.zeros(total, np.uint8)
    buf[slot_off[0] : slot_off[0] + L] = c.ref
    for i, s in enumerate(snv_rep):
        o = int(slot_off[1 + i])
        buf[o : o + L] = T[s]
    if n_win:
        wt = texts[kr]
        valid = cols < rlen[kr, None]
        dst = slot_off[1 + n_snv : 1 + n_snv + n_win, None] + cols
        buf[dst[valid]] = wt[valid]
    variant_bases = int(((buf >= ord("a")) & (buf <= ord("z"))).sum())
    hap_start = np.concatenate((np.full(1 + n_snv, c.region_start), c.region_start + w0[kr])).astype(np.int32)
    hap_stop = np.concatenate((np.full(1 + n_snv, c.region_stop), c.region_start + w1[kr])).astype(np.int32)
    hap_kind = np.concatenate(([0], np.ones(n_snv, np.uint8), np.full(n_win, 2, np.uint8))).astype(np.uint8)
    hap_site = np.concatenate((np.full(1 + n_snv, -1), r_site[kr])).astype(np.int32)
    hap_samples = np.concatenate(([0], np.asarray(snv_mask, np.uint16), w_mask[kr])).astype(np.uint16)
    hap_rep = np.concatenate(([-1], np.asarray(snv_rep, np.int32), r_sample[kr])).astype(np.int32)
    anchor_rel = np.concatenate((np.full(1 + n_snv, -1), a[kr])).astype(np.int32)
    hap_dlen = np.concatenate((np.zeros(1 + n_snv, np.int64), dlen[kr])).astype(np.int32)
    is_ref = np.zeros(n_hap, np.uint8)
    is_ref[0] = 1

    # ---- run-length position maps (haplotype.py:138-159) -------------------------------------
    w_ins = is_ins[kr]
    seg_cnt = np.concatenate((np.ones(1 + n_snv, np.int64), np.where(w_ins, 3, 2)))
    seg_off = np.concatenate(([0], np.cumsum(seg_cnt)))
    n_seg = int(seg_off[-1])
    seg_rel = np.zeros(n_seg, np.int32)
    seg_gen = np.zeros(n_seg, np.int32)
    seg_step = np.ones(n_seg, np.uint8)
    seg_gen[seg_off[:-1]] = hap_start
    if n_win:
        f = seg_off[1 + n_snv : -1]  # first segment of every window haplotype
        gs, a_k, k_k = hap_start[1 + n_snv :].astype(np.int64), a[kr], k[kr]
        ii = np.flatnonzero(w_ins)
        seg_rel[f[ii] + 1] = a_k[ii] + 1  # inserted bases repeat the anchor's coordinate
        seg_gen[f[ii] + 1] = gs[ii] + a_k[ii]
        seg_step[f[ii] + 1] = 0
        seg_rel[f[ii] + 2] = a_k[ii] + 1 + k_k[ii]
        seg_gen[f[ii] + 2] = gs[ii] + a_k[ii] + 1
        dd = np.flatnonzero(~w_ins)
        seg_rel[f[dd] + 1] = a_k[dd] + 1  # the base after the anchor jumps over the deleted ones
        seg_gen[f[dd] + 1] = gs[dd] + a_k[dd] + k_k[dd] + 1
    seg = marshal.SegmentTable(seg_off.astype(np.int64), seg_rel, seg_gen, seg_step)

    # ---- variant_alleles tables (haplotype.py:287-291, 160-183) ------------------------------
    refnib = _NIB_OF_ASCII[c.ref[c.site_pos]] if len(c.site_pos) else np.zeros(0, np.uint8)
    # full-length haplotypes: every SNV site the representative sample carries
    s_owner, s_site = [], []
    for i, s in enumerate(snv_rep):
        sites = np.flatnonzero((snv_carry >> s) & 1)
        s_owner.append(np.full(len(sites), 1 + i, np.int64))
        s_site.append(sites)
    # windows: SNV sites inside [w0, w1] the row's sample carries (minus the skipped tail of a
    # deletion window), plus the indel's own entry at the anchor
    lo = np.searchsorted(c.site_pos, w0[kr], side="left")
    hi = np.searchsorted(c.site_pos, w1[kr], side="right")
    own, sj = _ragged_ranges(lo, hi)
    rs = r_sample[kr][own]
    ok = ((snv_carry[sj] >> rs) & 1).astype(bool) & (c.site_pos[sj] < skip_from[kr][own])
    own, sj = own[ok], sj[ok]
    spos = c.site_pos[sj].astype(np.int64)
    w_idx = spos - w0[kr][own] + np.where(spos > p[kr][own], dlen[kr][own], 0)
    # merge SNV entries and anchor entries per window, ascending index
    anchor_owner = np.arange(n_win)
    all_owner = np.concatenate([np.concatenate(s_owner) if s_owner else np.zeros(0, np.int64),
                                1 + n_snv + own, 1 + n_snv + anchor_owner]).astype(np.int64)  # fmt: skip
    all_idx = np.concatenate([c.site_pos[np.concatenate(s_site)].astype(np.int64) if s_site else np.zeros(0, np.int64),
                              w_idx, a[kr]]).astype(np.int64)  # fmt: skip
    all_site = np.concatenate([np.concatenate(s_site) if s_site else np.zeros(0, np.int64), sj, r_site[kr]]).astype(np.int64)
    all_sample = np.concatenate([np.repeat(np.asarray(snv_rep, np.int64), [len(x) for x in s_site]) if s_site else np.zeros(0, np.int64),
                                 rs[ok] if len(ok) else np.zeros(0, np.int64), r_sample[kr]]).astype(np.int64)  # fmt: skip
    order = np.lexsort((all_idx, all_owner))
    all_owner, all_idx, all_site, all_sample = all_owner[order], all_idx[order], all_site[order], all_sample[order]
    site_is_snv = is_snv[all_site]
    n_ent = np.where(site_is_snv,
                     ((c.carry[all_site, 0] >> all_sample) & 1) + ((c.carry[all_site, 1] >> all_sample) & 1), 1).astype(np.int64)  # fmt: skip
    ent_off = np.concatenate(([0], np.cumsum(n_ent)))
    # REF allele nibble of an entry: the reference base for SNVs and insertions (one-base REF
    # allele), 0 for deletions (multi-base REF allele never equals a resolved base)
    ent_ref_site = np.where(c.site_kind[all_site] == DEL, 0, refnib[all_site]).astype(np.uint8)
    va_ref = np.repeat(ent_ref_site, n_ent)
    va_off = np.zeros(n_hap + 1, np.int64)
    np.add.at(va_off, all_owner + 1, 1)
    va_off = np.cumsum(va_off)
    alleles = marshal.AlleleTable(va_off, all_idx.astype(np.int32), ent_off.astype(np.int64), va_ref.astype(np.uint8))

    u = UnphasedSet(buf, slot_off, lens, total, seg, alleles, is_ref, hap_start, hap_stop, hap_kind, hap_site,
                    hap_samples, hap_rep, anchor_rel, hap_dlen, variant_bases)  # fmt: skip
    c._derived["u"] = u
    return u


class UnphasedHap:
    """Duck-typed haplotype (haplotype.py:23-77) over a member of an UnphasedSet, for the
    oracle and for `crispr_hawk_b200.search` in tests (small cohorts only)."""

    class _Seq:
        def __init__(self, text):
            self.sequence = text

        def __len__(self):
            return len(self.sequence)

    def __init__(self, c: UnphasedCohort, u: UnphasedSet, h: int):
        o, n = int(u.slot_off[h]), int(u.lens[h])
        self.sequence = UnphasedHap._Seq(u.ascii[o : o + n].tobytes().decode("ascii"))
        s0, s1 = int(u.seg.seg_off[h]), int(u.seg.seg_off[h + 1])
        vals = marshal.eval_segments(u.seg.seg_rel[s0:s1], u.seg.seg_gen[s0:s1], u.seg.seg_step[s0:s1], np.arange(n)).tolist()
        self.posmap = dict(enumerate(vals))
        self.posmap_rev = {g: i for i, g in enumerate(vals)}
        self.start, self.stop = int(u.hap_start[h]), int(u.hap_stop[h])
        names = [c.samples[s] for s in range(len(c.samples)) if (int(u.hap_samples[h]) >> s) & 1]
        self.samples = "REF" if u.is_ref[h] else ",".join(names)
        self.afs = {}
        self.id = f"h{h}"
        ref = c.ref.tobytes().decode("ascii")
        va = {}
        ids = []
        A = u.alleles
        rep = int(u.hap_rep[h])
        for j in range(int(A.va_off[h]), int(A.va_off[h + 1])):
            idx = int(A.va_idx[j])
            g = self.posmap[idx]
            site = int(np.searchsorted(c.site_pos, g - c.region_start))
            kd, pos = int(c.site_kind[site]), int(c.site_pos[site])
            if kd == SNV:
                ents = [(ref[pos], chr(_LETTER_OF_MASK[int(c.snv_alt[site, t])]), g)
                        for t in (0, 1) if (int(c.carry[site, t]) >> rep) & 1]  # fmt: skip
            elif kd == INS:
                o2, k2 = int(c.ins_off[site]), int(c.site_k[site])
                ents = [(ref[pos], ref[pos] + c.ins_pool[o2 : o2 + k2].tobytes().decode(), g)]
            else:
                ents = [(ref[pos : pos + int(c.site_k[site]) + 1], ref[pos], g)]
            va[idx] = ents
            ids.extend(f"chr1-{g}-{e[0]}/{e[1]}" for e in ents)
        self.variant_alleles = va
        self.variants = "NA" if u.is_ref[h] else ",".join(ids)

    def __len__(self):
        return len(self.sequence)


def unphased_haplotypes(c: UnphasedCohort, u: UnphasedSet = None) -> List[UnphasedHap]:
    """Objects for the haplotypes of `u` (default: all of the cohort's, in generator order)."""
    u = u or derive_unphased(c)
    return [UnphasedHap(c, u, h) for h in range(u.n_hap)]
