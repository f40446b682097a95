"""Host-side marshalling: reference-style haplotype objects -> flat arrays of the C-ABI.

Pure data movement (numpy); no scan arithmetic happens here. The haplotype
duck type is the reference's `Haplotype` (haplotype.py:23-77): `.sequence.sequence`,
`.posmap`, `.posmap_rev`, `.start/.stop`, `.samples`, `.variant_alleles`.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np

SLOT_ALIGN = 128  # HAWK_SLOT_ALIGN
SLOT_GAP = 128  # HAWK_SLOT_GAP: zero slots before the first and after every haplotype
CHUNK = 32
PADDING = 100  # region_constructor.py:21
GUIDESEQPAD = 10  # guide.py:21

_NIBBLE = {
    "A": 1, "C": 2, "G": 4, "T": 8, "R": 5, "Y": 10, "S": 6, "W": 9,
    "K": 12, "M": 3, "B": 14, "D": 13, "H": 11, "V": 7, "N": 15,
}  # fmt: skip  (encoder.py:18-34)


def hap_text(hap) -> str:
    seq = hap.sequence
    return seq.sequence if hasattr(seq, "sequence") else str(seq)


def layout(lengths: Sequence[int]) -> Tuple[np.ndarray, int]:
    """Aligned exclusive prefix of the haplotype lengths (mirror of hawk_layout)."""
    lens = np.asarray(lengths, dtype=np.int64)
    padded = (lens + SLOT_ALIGN - 1) // SLOT_ALIGN * SLOT_ALIGN + SLOT_GAP
    off = np.full(len(lens) + 1, SLOT_GAP, dtype=np.int64)
    off[1:] += np.cumsum(padded)
    return off, int(off[-1])


class InvalidSequence(ValueError):
    def __init__(self, hap_index: int, position: int, char: str):
        super().__init__(f"invalid character {char!r} at {position} of haplotype {hap_index}")
        self.hap_index, self.position, self.char = hap_index, position, char


def stage_ascii(texts: Sequence[str]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Copy haplotype texts into the zero-initialised slot space (1 byte per slot)."""
    lens = np.array([len(t) for t in texts], dtype=np.int32)
    off, total = layout(lens)
    buf = np.zeros(total, dtype=np.uint8)
    for h, t in enumerate(texts):
        try:
            raw = t.encode("ascii")
        except UnicodeEncodeError as e:
            raise InvalidSequence(h, e.start, t[e.start]) from e
        if b"\x00" in raw:  # NUL marks an unused slot on the device; never a base
            raise InvalidSequence(h, raw.index(b"\x00"), "\x00")
        buf[off[h] : off[h] + len(raw)] = np.frombuffer(raw, dtype=np.uint8)
    return buf, off, lens


def posmap_values(hap) -> np.ndarray:
    pm = hap.posmap
    if isinstance(pm, dict):
        # haplotype.py:101-103,138-159: keys are 0..L-1 inserted in ascending order
        vals = np.fromiter(pm.values(), dtype=np.int64, count=len(pm))
        n = len(vals)
        if n and (next(iter(pm)) != 0 or next(reversed(pm)) != n - 1):
            vals = np.array([pm[i] for i in range(n)], dtype=np.int64)
        return vals
    return np.asarray(pm, dtype=np.int64)


def posmap_segments(vals: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Run-length encode a posmap (haplotype.py:90-104,138-159) into the C-ABI's
    segments. A new segment starts wherever consecutive coordinates do not differ
    by +1: after a deletion (jump) and at every inserted base (inserted bases
    repeat the anchor's coordinate, haplotype.py:151-158). All emitted segments
    have step 1; the ABI's step-0 form is only used by generators that know the
    insertion runs."""
    n = len(vals)
    if n == 0:
        return (np.zeros(1, np.int32), np.zeros(1, np.int32), np.ones(1, np.uint8))
    brk = np.flatnonzero(np.diff(vals) != 1) + 1
    rel = np.concatenate(([0], brk)).astype(np.int64)
    return rel.astype(np.int32), vals[rel].astype(np.int32), np.ones(len(rel), np.uint8)


def eval_segments(rel, gen, step, idx):
    k = np.searchsorted(rel, idx, side="right") - 1
    return gen[k] + step[k].astype(np.int64) * (np.asarray(idx) - rel[k])


def scan_bounds(hap, region_start: int, region_stop: int, pamlen: int) -> Tuple[int, int]:
    """Mirror of compute_scan_start_stop (search_guides.py:49-84): haplotype-relative
    [start, stop) of the PAM scan, from the haplotype's own position maps."""
    if hasattr(hap, "scan_bounds"):  # device-materialised haplotypes answer from their segments
        return hap.scan_bounds(region_start, region_stop, pamlen)
    rev = hap.posmap_rev
    stop_g = min(region_stop - PADDING, hap.stop)
    if stop_g == region_stop - PADDING and stop_g not in rev:
        top = max(rev.keys())
        for g in range(stop_g, top + 1):
            if g in rev:
                stop_g = g
                break
    scan_stop = rev[stop_g] - pamlen + 1
    scan_start = rev[max(region_start + PADDING, hap.start)]
    return scan_start, scan_stop


@dataclass
class AlleleTable:
    va_off: np.ndarray  # n_hap + 1
    va_idx: np.ndarray  # sites
    va_ent_off: np.ndarray  # sites + 1
    va_ref: np.ndarray  # entries (nibble of a single-base REF allele, else 0)


def allele_table(haps) -> AlleleTable:
    """Flatten `variant_alleles` (haplotype.py:287-291): rel index -> [(ref, alt, pos)]."""
    va_off = [0]
    va_idx: List[int] = []
    ent_off = [0]
    refs: List[int] = []
    for h in haps:
        va: Dict[int, list] = getattr(h, "variant_alleles", None) or {}
        for idx in sorted(va):
            va_idx.append(int(idx))
            for entry in va[idx]:
                ref = entry[0]
                refs.append(_NIBBLE.get(ref, 0) if len(ref) == 1 else 0)
            ent_off.append(len(refs))
        va_off.append(len(va_idx))
    return AlleleTable(
        np.asarray(va_off, np.int64),
        np.asarray(va_idx, np.int32),
        np.asarray(ent_off, np.int64),
        np.asarray(refs, np.uint8),
    )


@dataclass
class SegmentTable:
    seg_off: np.ndarray
    seg_rel: np.ndarray
    seg_gen: np.ndarray
    seg_step: np.ndarray


def segment_table(haps) -> SegmentTable:
    offs = [0]
    rels, gens, steps = [], [], []
    for h in haps:
        r, g, s = posmap_segments(posmap_values(h))
        rels.append(r)
        gens.append(g)
        steps.append(s)
        offs.append(offs[-1] + len(r))
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)  # noqa: E731
    return SegmentTable(
        np.asarray(offs, np.int64), cat(rels, np.int32), cat(gens, np.int32), cat(steps, np.uint8)
    )


def pam_nibbles(seq: str) -> List[int]:
    return [_NIBBLE[c] for c in seq.upper()]


# --------------------------------------------------------------------------- N2: variant tables
def normalise_variant(ref: str, alt: str, pos: int):
    """variant.py:456-486 (adjust_multiallelic): the form annotation._parse_variant compares in."""
    if len(ref) == len(alt):
        return ref[0], alt[0], pos
    if len(ref) > len(alt):
        return ref[len(alt) - 1 :], alt[-1], pos + len(alt) - 1
    return ref[-1], alt[len(ref) - 1 :], pos + len(ref) - 1


class VariantTable:
    """Per-haplotype variant lists for hawk_batch_set_variants, plus the ids in table order."""

    def __init__(self, var_off, var_pos, var_reflen, var_altlen, var_altoff, alt_pool, ids, ambiguous=False):
        self.var_off, self.var_pos, self.var_reflen = var_off, var_pos, var_reflen
        self.var_altlen, self.var_altoff, self.alt_pool, self.ids = var_altlen, var_altoff, alt_pool, ids
        # a haplotype with two variants at one normalised position: annotation._create_variants_map
        # keeps whichever its set iteration yields last (annotation.py:96, 129-160), i.e. the
        # reference's own answer depends on Python's hash order -- such lists are not annotated
        # on the device (the seam hands them to the reference's functions)
        self.ambiguous = ambiguous


def variant_table(haps) -> VariantTable:
    """Parse every haplotype's `variants` string ('NA' or 'chrom-pos-ref/alt,...',
    variant.py:436-453) into the normalised, position-sorted table the device walks."""
    off, pos, rl, al, ao, pool, ids = [0], [], [], [], [], [], []
    n_pool = 0
    ambiguous = False
    for h in haps:
        rows = []
        if h.variants and h.variants != "NA":
            # annotation.py:96 splits into a set: duplicates collapse
            for vid in dict.fromkeys(h.variants.split(",")):
                parts = vid.split("-")
                ref, alt = parts[2].split("/")
                r2, a2, p2 = normalise_variant(ref, alt, int(parts[1]))
                rows.append((p2, vid, len(r2), a2))
        rows.sort(key=lambda t: t[0])
        ambiguous = ambiguous or any(rows[k][0] == rows[k - 1][0] for k in range(1, len(rows)))
        ids.append([t[1] for t in rows])
        for p2, _, r_len, a2 in rows:
            pos.append(p2)
            rl.append(r_len)
            al.append(len(a2))
            ao.append(n_pool)
            pool.append(a2)
            n_pool += len(a2)
        off.append(len(pos))
    return VariantTable(
        np.asarray(off, np.int64), np.asarray(pos, np.int32), np.asarray(rl, np.int32), np.asarray(al, np.int32),
        np.asarray(ao, np.int64), np.frombuffer("".join(pool).encode("ascii"), np.uint8).copy() if pool else np.zeros(0, np.uint8),
        ids, ambiguous,
    )  # fmt: skip
