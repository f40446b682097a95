"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the post-search pure functions on guides
(N2 of the scope table): `polish_guide_variants`, `annotate_variants_afs`,
`reverse_guides`, `gc_content` of the reference's annotation.py. Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU arm may import it.

Pinned against the unmodified reference: tests/golden/annot.json.gz holds the reference's own
outputs for every phased golden case (tests/golden/make_golden_annot.py), and
tests/test_annot_oracle.py re-runs the live reference where it exists. `gc_fraction` is
Biopython's (1.83, ambiguous="remove"), which is not installed here: it is restated from its
documentation (G+C+S over A+C+G+T+S+W, case-insensitive), the one unpinned piece.

Scope: phased / variant-free searches, haplotypes that carry at most one variant per
normalised position (what a phased VCF gives a haplotype copy). With two
variants at one position the reference's own result depends on Python's set iteration order
(annotation.py:266-274: `offset` leaks from one variant id to the next)."""

from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

GUIDESEQPAD = 10  # guide.py:21

# utils.py:46-79 (RC): IUPAC-aware complement, case preserved
_RC = {
    "A": "T", "C": "G", "G": "C", "T": "A", "U": "A", "R": "Y", "Y": "R", "M": "K", "K": "M",
    "H": "D", "D": "H", "B": "V", "V": "B", "N": "N", "S": "S", "W": "W",
}  # fmt: skip
RC = dict(_RC)
RC.update({k.lower(): v.lower() for k, v in _RC.items()})


def normalise_variant(ref: str, alt: str, pos: int) -> Tuple[str, str, int]:
    """variant.py:456-486 (adjust_multiallelic)."""
    if len(ref) == len(alt):
        return ref[0], alt[0], pos
    if len(ref) > len(alt):  # deletion
        return ref[len(alt) - 1 :], alt[-1], pos + len(alt) - 1
    return ref[-1], alt[len(ref) - 1 :], pos + len(ref) - 1  # insertion


def parse_variant(variant_id: str) -> Tuple[str, int, str, str]:
    """annotation.py:54-73 (_parse_variant): 'chrom-pos-ref/alt' -> normalised fields."""
    parts = variant_id.split("-")
    ref, alt = parts[2].split("/")
    ref_, alt_, pos_ = normalise_variant(ref, alt, int(parts[1]))
    return parts[0], pos_, ref_, alt_


def _find_insertion_stop(seg: str) -> int:
    """annotation.py:178-194: index of the first upper-case character, 0 when there is none."""
    assert not seg[0].isupper() and not all(nt.isupper() for nt in seg)
    return next((i for i, nt in enumerate(seg) if nt.isupper()), 0)


def _check_insertion(seg: str, alt: str, posrel: int, pos: int, stop: int, is_snv: bool) -> bool:
    """annotation.py:197-226."""
    if is_snv:
        return False
    if posrel == 0 and alt.endswith(seg.upper()[: _find_insertion_stop(seg)]):
        return True
    return bool(pos == stop and alt.startswith(seg.upper()))


def _check_snv(seg: str, alt: str) -> bool:
    """annotation.py:229-243."""
    return seg.islower() and seg.upper() == alt


def polish_guide_variants(guidepam: str, posmap: Sequence[int], stop: int, variants: Sequence[str]) -> List[str]:
    """annotation.py:246-281: the haplotype's variants that are visible in this guide, sorted.
    `guidepam` = the guide's core (guide + PAM as it lies on the forward strand), `posmap` its
    per-base genomic coordinates, `stop` the guide's stop coordinate."""
    vmap = [(parse_variant(v)[1], v) for v in variants]
    positions = {p for p, _ in vmap}
    out = set()
    for i in range(len(guidepam)):
        offset = 0
        p = posmap[i]
        if p in positions:
            for v in [v for q, v in vmap if q == p]:
                _, _, ref, alt = parse_variant(v)
                is_snv = len(ref) == len(alt)
                if not is_snv:
                    offset = abs(len(ref) - len(alt)) if len(ref) < len(alt) else 0
                seg = guidepam[i : i + offset + 1]
                if _check_insertion(seg, alt, i, p, stop, is_snv) or _check_snv(seg, alt):
                    out.add(v)
    return sorted(out)


def format_af(af: float) -> str:
    """annotation.py:316-331 (_format_af)."""
    s = f"{af:.10f}".rstrip("0").rstrip(".")
    decimal_digits = len(s.split(".")) if "." in s else 0
    return f"{af:.6e}" if decimal_digits > 3 else str(round(af, 6))


def gc_fraction(seq: str):
    """Bio.SeqUtils.gc_fraction(seq) with the default ambiguous="remove" (Biopython 1.83)."""
    s = seq.upper()
    gc = sum(s.count(c) for c in "CGS")
    n = gc + sum(s.count(c) for c in "ATW")
    return gc / n if n else 0


def reverse_complement(text: str) -> str:
    """guide.py:245-255 / utils.py:123-137."""
    return "".join(RC[c] for c in text[::-1])


def annotate_guide(sequence: str, guidelen: int, pamlen: int, strand: int, right: bool, stop: int,
                   posmap: Sequence[int], variants: str, afs: Dict[str, float]):  # fmt: skip
    """annotation.py:563-572 for one guide as `search()` returned it: (variants string,
    afs_str, sequence after reverse_guides, `right` after it, gc string)."""
    core = sequence[GUIDESEQPAD:-GUIDESEQPAD]
    if variants == "NA":  # _is_reference_guide, annotation.py:104-126
        v = "NA"
    else:
        v = ",".join(polish_guide_variants(core, posmap, stop, variants.split(",")))
    if v != "NA":  # annotate_variants_afs, annotation.py:334-365
        vals = [format_af(afs[x]) if str(afs[x]) != "nan" else "NA" for x in v.split(",")]
    else:
        vals = ["NA"]
    # Guide.afs_str setter, guide.py:311-328
    afs_str = "NA" if not vals or (len(set(vals)) == 1 and vals[0] == "NA") else ",".join(vals)
    if strand == 1:  # reverse_guides, annotation.py:27-51
        sequence = reverse_complement(sequence)
        right = not right
    core = sequence[GUIDESEQPAD:-GUIDESEQPAD]
    guide = core[pamlen:] if right else core[:-pamlen]
    gc = gc_fraction(guide)
    if not isinstance(gc, float):  # Guide.gc setter, guide.py:598-618 -> CrisprHawkGcContentError
        raise ValueError("GC content calculation failed")
    return v, afs_str, sequence, right, str(gc)
