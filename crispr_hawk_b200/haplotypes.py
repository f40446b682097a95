"""N1 (next row of the scope table): phased haplotypes materialised on the device.

The reference assembles every haplotype on the host, one variant at a time, rewriting a
Python list of characters and two dicts of `len(haplotype)` entries per indel
(haplotype.py:106-159, 185-252) -- the real wall of a run at BASELINE scale.
Here the host keeps what it is good at (which variants each haplotype copy carries, sample
strings, variant ids, allele frequencies: all metadata) and hands the device only the
*edit lists*; `hawk_batch_create_from_edits` materialises the texts in HBM, packs them and
attaches run-length coordinate maps. Nothing of haplotype length ever crosses PCIe or lives
in a Python object unless somebody asks for it.

`EditHaplotype` is duck-typed like the reference's `Haplotype` (haplotype.py:23-77) as far
as `search()` and its callers read it; `.sequence`, `.posmap` and `.posmap_rev` are lazy.
Edits must be sorted, non-overlapping SNVs / anchored insertions / anchored deletions: what
`add_variants_phased` produces for a phased VCF whose records do not overlap on a haplotype
(SURVEY.md Appendix B)."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _cabi, marshal
from .encoder import PackedRegion

PADDING = marshal.PADDING


@dataclass
class Edit:
    pos: int  # genomic coordinate of the anchor base (VCF POS)
    ref: str  # REF allele
    alt: str  # ALT allele


class SegmentMap:
    """`posmap` of one haplotype as run-length segments: int -> genomic coordinate."""

    def __init__(self, rel: np.ndarray, gen: np.ndarray, step: np.ndarray, length: int):
        self.rel, self.gen, self.step, self.length = rel, gen, step, int(length)

    def __len__(self) -> int:
        return self.length

    def __getitem__(self, i: int) -> int:
        if i < 0 or i >= self.length:
            raise KeyError(i)
        k = int(np.searchsorted(self.rel, i, side="right")) - 1
        return int(self.gen[k] + (i - self.rel[k] if self.step[k] else 0))

    def values(self) -> np.ndarray:
        return marshal.eval_segments(self.rel, self.gen, self.step, np.arange(self.length))

    def last_index_of(self, g: int) -> Optional[int]:
        """posmap_rev[g] (haplotype.py:104,159: the LAST index carrying coordinate g), or None
        when the coordinate was deleted."""
        seg_len = np.diff(np.append(self.rel, self.length))
        hit1 = self.step.astype(bool) & (self.gen <= g) & (g < self.gen + seg_len)
        hit0 = (~self.step.astype(bool)) & (self.gen == g)
        best = -1
        if hit1.any():
            k = np.flatnonzero(hit1)
            best = int((self.rel[k] + (g - self.gen[k])).max())
        if hit0.any():
            k = np.flatnonzero(hit0)
            best = max(best, int((self.rel[k] + seg_len[k] - 1).max()))
        return best if best >= 0 else None


class _LazySeq:
    def __init__(self, hap):
        self._hap = hap

    @property
    def sequence(self) -> str:
        return self._hap.text()

    def __len__(self) -> int:
        return len(self._hap)

    def __str__(self) -> str:  # haplotypes_table writes f"{hap.sequence}" (haplotypes.py:836-841)
        return self._hap.text()


class EditHaplotype:
    def __init__(self, batch: "_cabi.Batch", index: int, length: int, posmap: SegmentMap, start: int, stop: int,
                 samples: str, variants: str, afs: Dict[str, float], hapid: str):  # fmt: skip
        self._batch, self._index, self._len = batch, index, int(length)
        self.posmap = posmap
        self.start, self.stop = start, stop
        self.samples, self.variants, self.afs, self.id = samples, variants, afs, hapid
        self.variant_alleles: dict = {}
        self.sequence = _LazySeq(self)
        self._text: Optional[str] = None
        self._rev: Optional[dict] = None

    def __len__(self) -> int:
        return self._len

    def text(self) -> str:
        if self._text is None:
            self._text = self._batch.export_text(self._index)
        return self._text

    def __getitem__(self, idx):
        return list(self.text()[idx]) if isinstance(idx, slice) else self.text()[idx]

    @property
    def posmap_rev(self) -> dict:
        if self._rev is None:  # full dict only on request (the scan itself never needs it)
            self._rev = {int(g): i for i, g in enumerate(self.posmap.values())}
        return self._rev

    def scan_bounds(self, region_start: int, region_stop: int, pamlen: int):
        """compute_scan_start_stop (search_guides.py:49-84) from the segments."""
        stop_g = min(region_stop - PADDING, self.stop)
        i_stop = self.posmap.last_index_of(stop_g)
        if i_stop is None and stop_g == region_stop - PADDING:
            top = int(self.posmap.values().max())
            g = stop_g
            while i_stop is None and g <= top:
                g += 1
                i_stop = self.posmap.last_index_of(g)
        if i_stop is None:
            raise KeyError(stop_g)
        i_start = self.posmap.last_index_of(max(region_start + PADDING, self.start))
        if i_start is None:
            raise KeyError(max(region_start + PADDING, self.start))
        return i_start, i_stop - pamlen + 1


def build_phased(ref_text: str, region_start: int, hap_edits: Sequence[Sequence[Edit]],
                 samples: Optional[Sequence[str]] = None, variants: Optional[Sequence[str]] = None,
                 afs: Optional[Sequence[Dict[str, float]]] = None, ids: Optional[Sequence[str]] = None,
                 contig: str = "chr1", ctx: Optional["_cabi.Context"] = None):  # fmt: skip
    """Materialise haplotypes `REF + edits` on the device. `ref_text` is the padded region's
    reference (upper-case), `region_start` the genomic coordinate of its first base;
    `hap_edits[h]` the edits of haplotype h sorted by position (an empty list = the REF
    haplotype). Returns (haplotypes, PackedRegion) ready for `search()`."""
    ctx = ctx or _cabi.Context.default()
    n = len(hap_edits)
    ref = np.frombuffer(ref_text.encode("ascii"), np.uint8)
    edit_off = np.zeros(n + 1, np.int64)
    pos, rl, al, pool = [], [], [], []
    for h, edits in enumerate(hap_edits):
        for e in edits:
            if ref_text[e.pos - region_start : e.pos - region_start + len(e.ref)].upper() != e.ref.upper():
                raise ValueError(f"Mismatching reference alleles in VCF and reference sequence at {e.pos}")  # haplotype.py:206-210
            pos.append(e.pos - region_start)
            rl.append(len(e.ref))
            al.append(len(e.alt))
            pool.append(e.alt.upper())
        edit_off[h + 1] = len(pos)
    al_arr = np.asarray(al, np.int32)
    ao_arr = np.concatenate(([0], np.cumsum(al_arr)[:-1])).astype(np.int64) if len(al) else np.zeros(0, np.int64)
    pool_arr = np.frombuffer("".join(pool).encode("ascii"), np.uint8) if pool else np.zeros(0, np.uint8)
    batch = _cabi.Batch.from_edits(ctx, ref, region_start, edit_off, np.asarray(pos, np.int32), np.asarray(rl, np.int32),
                                   al_arr, ao_arr, pool_arr)  # fmt: skip
    region_stop = region_start + len(ref_text) - 1
    haps: List[EditHaplotype] = []
    for h, edits in enumerate(hap_edits):
        rel, gen, step = [0], [region_start], [1]
        shift = 0
        for e in edits:
            op = e.pos - region_start + shift
            if len(e.alt) > 1:
                rel += [op + 1, op + len(e.alt)]
                gen += [e.pos, e.pos + 1]
                step += [0, 1]
            elif len(e.ref) > 1:
                rel.append(op + 1)
                gen.append(e.pos + len(e.ref))
                step.append(1)
            shift += len(e.alt) - len(e.ref)
        pm = SegmentMap(np.asarray(rel, np.int64), np.asarray(gen, np.int64), np.asarray(step, np.uint8), int(batch.lens[h]))
        is_ref = len(edits) == 0
        haps.append(EditHaplotype(
            batch, h, int(batch.lens[h]), pm, region_start, region_stop,
            (samples[h] if samples else ("REF" if is_ref else f"hap{h}")),
            (variants[h] if variants else ("NA" if is_ref else ",".join(f"{contig}-{e.pos}-{e.ref}/{e.alt}" for e in edits))),
            (afs[h] if afs else {}), (ids[h] if ids else f"hap{h}"),
        ))  # fmt: skip
    packed = PackedRegion(batch, [None] * n)
    packed.owners = [id(h) for h in haps]
    return haps, packed


# --------------------------------------------------------------------------- drop-in seam (N1)
# Mirror of crisprhawk.haplotypes.add_variants_phased (haplotypes.py:716-751): the reference's
# own VariantRecord lists in, haplotypes out -- but as EditHaplotypes over ONE device batch
# built from edit lists, instead of one host rewrite of the sequence and of two
# len(haplotype)-entry dicts per variant (haplotype.py:106-159, 185-252). install() rebinds
# it in `crisprhawk.haplotypes`; `add_variants` (:754-792) finds it through the module globals.
class UnsupportedShape(Exception):
    """The region's records do not have the shape the device builder takes; the reference's
    own builder is used for it."""


_reference_add_variants_phased = None  # set by install()


def _copy_edits(variants, ref_text: str, region_start: int):
    """One chromosome copy: `_sort_variants` order for the ids (SNPs by position, then indels by
    position, haplotype.py:494-512), position order for the edits; REF alleles checked against
    the reference text like haplotype.py:206-210."""
    snps = sorted((v for v in variants if v.vtype[0] == "snp"), key=lambda v: v.position)
    indels = sorted((v for v in variants if v.vtype[0] != "snp"), key=lambda v: v.position)
    ordered = snps + indels
    ids = [v.id[0] for v in ordered]
    afs = {v.id[0]: v.afs[0] for v in ordered}
    edits = sorted((Edit(int(v.position), v.ref, v.alt[0]) for v in ordered), key=lambda e: e.pos)
    end = region_start + len(ref_text)
    last_stop = -1
    for e in edits:
        if len(e.ref) > 1 and len(e.alt) > 1:
            raise UnsupportedShape("complex substitution")
        if e.pos <= last_stop:
            raise UnsupportedShape("records overlap on one chromosome copy")
        if e.pos < region_start or e.pos + len(e.ref) > end:
            raise UnsupportedShape("record runs past the region")
        i = e.pos - region_start
        if ref_text[i : i + len(e.ref)] != e.ref:
            raise ValueError(
                f"Mismatching reference alleles in VCF and reference sequence "
                f"at position {e.pos} ({ref_text[i : i + len(e.ref)]} - {e.ref})"
            )  # haplotype.py:206-210
        if e.ref[0].upper() != e.alt[0].upper() and (len(e.ref) > 1 or len(e.alt) > 1):
            raise UnsupportedShape("indel without an anchor base")
        last_stop = e.pos + len(e.ref) - 1
    return tuple((e.pos, e.ref, e.alt) for e in edits), edits, ",".join(ids), afs


def plan_phased(ref_text: str, region_start: int, samples: Sequence[str], variants):
    """compute_haplotypes_phased (:143-171) + _solve_haplotypes_phased (:297-333) +
    collapse_haplotypes (:274-294) on edit lists: returns the collapsed haplotypes as dicts
    {edits, samples, variants, afs}, REF first, in the reference's first-seen order. Haplotype
    copies are merged when their edit lists are equal -- the reference merges on equal
    sequences, which differs only for un-normalised indel records spelling the same text."""
    per_sample = {s: ([], []) for s in samples}
    for v in variants:
        for copy in (0, 1):
            for smp in v.samples[0][copy]:
                per_sample[smp][copy].append(v)
    groups: Dict[tuple, dict] = {(): dict(edits=[], members=[], variants="NA", afs={})}
    for smp in samples:
        c0, c1 = per_sample[smp]
        if not c0 and not c1:
            continue  # samples without variants are dropped (:170)
        built = [_copy_edits(c, ref_text, region_start) for c in (c0, c1)]
        if built[0][0] == built[1][0]:  # ishomozygous (:228): one haplotype, S:1|1 (haplotype.py:331-354)
            entries = [(built[0], f"{smp}:1|1")]
        else:
            entries = [(built[0], f"{smp}:1|0"), (built[1], f"{smp}:0|1")]
        for (key, edits, ids, afs), label in entries:
            g = groups.get(key)
            if g is None:
                g = groups[key] = dict(edits=edits, members=[], variants=ids, afs=afs)
            g["members"].append(label)
    out = []
    for key, g in groups.items():
        is_ref = key == ()
        out.append(dict(edits=g["edits"], samples="REF" if is_ref else ",".join(dict.fromkeys(g["members"])),
                        variants="NA" if is_ref else g["variants"], afs={} if is_ref else g["afs"]))  # fmt: skip
    return out


# ---- the same plan on arrays: one pass over the records, numpy per chromosome copy -------------
# plan_phased / build_phased above touch every (variant, haplotype copy) pair in Python several
# times (three sorts, an Edit object, four checks each): 0.7 s for 400 copies of 1,000 variants.
# Everything that depends on the record alone is computed once per record here; a copy is then a
# few numpy calls on an index array. Inputs with anything unusual -- a record the device builder
# does not take, a REF allele that does not match, overlapping records on a copy -- leave this
# path untouched and go through the functions above, which raise what has to be raised in the
# reference's order.
class _Unusual(Exception):
    pass


class _RecordTable:
    def __init__(self, variants, ref_text: str, region_start: int):
        n = len(variants)
        self.pos = np.fromiter((int(v.position) for v in variants), np.int64, n)
        refs = [v.ref for v in variants]
        alts = [v.alt[0] for v in variants]
        self.reflen = np.fromiter((len(r) for r in refs), np.int32, n)
        self.altlen = np.fromiter((len(a) for a in alts), np.int32, n)
        self.is_snp = np.fromiter((v.vtype[0] == "snp" for v in variants), np.bool_, n)
        self.ids = [v.id[0] for v in variants]
        self.afs = [v.afs[0] for v in variants]
        end = region_start + len(ref_text)
        ok = ~((self.reflen > 1) & (self.altlen > 1)) & (self.pos >= region_start) & (self.pos + self.reflen <= end)
        if not ok.all():
            raise _Unusual
        seen: Dict[tuple, int] = {}
        canon = np.empty(n, np.int32)
        for k, (p, r, a) in enumerate(zip(self.pos.tolist(), refs, alts)):
            i = p - region_start
            if ref_text[i : i + len(r)] != r or (r[0].upper() != a[0].upper() and (len(r) > 1 or len(a) > 1)):
                raise _Unusual
            canon[k] = seen.setdefault((p, r, a), k)
        self.canon = canon  # first record with the same (position, REF, ALT): the grouping key of a copy
        self.altoff = np.concatenate(([0], np.cumsum(self.altlen)[:-1])).astype(np.int64) if n else np.zeros(0, np.int64)
        self.pool = np.frombuffer("".join(a.upper() for a in alts).encode("ascii"), np.uint8) if n else np.zeros(0, np.uint8)


def _copy_arrays(T: _RecordTable, c: np.ndarray):
    """_copy_edits on record indices: (grouping key, edit order, id order)."""
    snp = T.is_snp[c]
    s_idx, i_idx = c[snp], c[~snp]
    ids_idx = np.concatenate((s_idx[np.argsort(T.pos[s_idx], kind="stable")], i_idx[np.argsort(T.pos[i_idx], kind="stable")]))
    e_idx = ids_idx[np.argsort(T.pos[ids_idx], kind="stable")]
    ep = T.pos[e_idx]
    if len(ep) > 1 and bool((ep[1:] <= ep[:-1] + T.reflen[e_idx][:-1] - 1).any()):
        raise _Unusual  # records overlap on one chromosome copy
    return T.canon[e_idx].tobytes(), e_idx, ids_idx


def plan_phased_arrays(ref_text: str, region_start: int, samples: Sequence[str], variants):
    """plan_phased with the haplotypes' edits as index arrays into a record table. Returns
    (record table, [{e_idx, samples, variants, afs}]), REF first; raises _Unusual for inputs that
    belong to plan_phased."""
    T = _RecordTable(variants, ref_text, region_start)
    per_sample = {s: ([], []) for s in samples}
    for r, v in enumerate(variants):
        for copy in (0, 1):
            for smp in v.samples[0][copy]:
                per_sample[smp][copy].append(r)
    empty = np.zeros(0, np.int32)
    groups: Dict[bytes, dict] = {b"": dict(e_idx=empty, ids_idx=empty, members=[])}
    for smp in samples:
        c0, c1 = per_sample[smp]
        if not c0 and not c1:
            continue  # samples without variants are dropped (:170)
        built = [_copy_arrays(T, np.asarray(c, np.int32)) for c in (c0, c1)]
        if built[0][0] == built[1][0]:  # ishomozygous (:228)
            entries = [(built[0], f"{smp}:1|1")]
        else:
            entries = [(built[0], f"{smp}:1|0"), (built[1], f"{smp}:0|1")]
        for (key, e_idx, ids_idx), label in entries:
            g = groups.get(key)
            if g is None:
                g = groups[key] = dict(e_idx=e_idx, ids_idx=ids_idx, members=[])
            g["members"].append(label)
    out = []
    for key, g in groups.items():
        is_ref = key == b""
        order = g["ids_idx"].tolist()
        out.append(dict(e_idx=g["e_idx"], samples="REF" if is_ref else ",".join(dict.fromkeys(g["members"])),
                        variants="NA" if is_ref else ",".join(T.ids[i] for i in order),
                        afs={} if is_ref else {T.ids[i]: T.afs[i] for i in order}))  # fmt: skip
    return T, out


def build_phased_arrays(ref_text: str, region_start: int, T: _RecordTable, plan, contig: str = "chr1",
                        ctx: Optional["_cabi.Context"] = None):  # fmt: skip
    """build_phased for a plan of plan_phased_arrays: the CSR edit arrays by fancy indexing, the
    run-length maps by numpy per haplotype."""
    ctx = ctx or _cabi.Context.default()
    n = len(plan)
    ref = np.frombuffer(ref_text.encode("ascii"), np.uint8)
    counts = np.fromiter((len(p["e_idx"]) for p in plan), np.int64, n)
    edit_off = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)
    e_all = np.concatenate([p["e_idx"] for p in plan]) if n else np.zeros(0, np.int32)
    batch = _cabi.Batch.from_edits(ctx, ref, region_start, edit_off, (T.pos[e_all] - region_start).astype(np.int32),
                                   T.reflen[e_all], T.altlen[e_all], T.altoff[e_all], T.pool if len(T.pool) else np.zeros(1, np.uint8))  # fmt: skip
    region_stop = region_start + len(ref_text) - 1
    haps: List[EditHaplotype] = []
    for h, p in enumerate(plan):
        e = p["e_idx"]
        pos_g, rl, al = T.pos[e], T.reflen[e].astype(np.int64), T.altlen[e].astype(np.int64)
        d = al - rl
        op = pos_g - region_start + (np.cumsum(d) - d)
        ins, dele = al > 1, (rl > 1) & (al <= 1)
        nseg = np.where(ins, 2, np.where(dele, 1, 0))
        at = np.cumsum(nseg) - nseg + 1
        m = int(nseg.sum()) + 1
        rel, gen, step = np.zeros(m, np.int64), np.full(m, region_start, np.int64), np.ones(m, np.uint8)
        i, j = np.flatnonzero(ins), np.flatnonzero(dele)
        rel[at[i]], gen[at[i]], step[at[i]] = op[i] + 1, pos_g[i], 0
        rel[at[i] + 1], gen[at[i] + 1] = op[i] + al[i], pos_g[i] + 1
        rel[at[j]], gen[at[j]] = op[j] + 1, pos_g[j] + rl[j]
        pm = SegmentMap(rel, gen, step, int(batch.lens[h]))
        haps.append(EditHaplotype(batch, h, int(batch.lens[h]), pm, region_start, region_stop, p["samples"], p["variants"],
                                  p["afs"], f"hap{h}"))  # fmt: skip
    packed = PackedRegion(batch, [None] * n)
    packed.owners = [id(h) for h in haps]
    return haps, packed


def add_variants_phased(haplotypes, region, vcfs, variants, phased: bool, debug: bool):
    """Drop-in for crisprhawk.haplotypes.add_variants_phased (haplotypes.py:716-751)."""
    ref_text = region.sequence.sequence
    try:
        if len(haplotypes) != 1 or not ref_text.isupper():
            raise UnsupportedShape("not a single reference haplotype")
        try:
            T, plan = plan_phased_arrays(ref_text, region.start, vcfs[region.contig].samples, variants)
            haps, packed = build_phased_arrays(ref_text, region.start, T, plan, contig=region.contig)
        except _Unusual:
            plan = plan_phased(ref_text, region.start, vcfs[region.contig].samples, variants)
            haps, packed = build_phased(ref_text, region.start, [p["edits"] for p in plan], [p["samples"] for p in plan],
                                        [p["variants"] for p in plan], [p["afs"] for p in plan],
                                        [f"hap{i}" for i in range(len(plan))], contig=region.contig)  # fmt: skip
    except UnsupportedShape:
        if _reference_add_variants_phased is None:
            raise
        return _reference_add_variants_phased(haplotypes, region, vcfs, variants, phased, debug)
    ref_afs = getattr(haplotypes[0], "afs", None)
    if ref_afs is not None:
        haps[0].afs = ref_afs
    haps[0]._region_pack = packed  # crispr_hawk_b200.encoder.encode_region hands it to search()
    return haps
