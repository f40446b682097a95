"""N2 (next row of the scope table): the post-search pure functions of the reference's
annotation.py -- `_annotate_variants` / `polish_guide_variants` (:246-315),
`annotate_variants_afs` (:334-365), `reverse_guides` (:27-51), `gc_content` (:513-541) --
for the guide table of a phased / variant-free search.

The per-base work (which variants are visible in which guide, reverse complements, GC counts)
runs on the device over the resident table (`hawk_result_annotate`); what is left here is
string assembly: joining the sorted variant ids, looking up and formatting their allele
frequencies, `str()` of the GC fraction. The reference does all of it per `Guide` in Python
and re-parses the haplotype's whole variant list for every guide."""

from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

from . import marshal
from .errors import error_class, exception_handler


def format_af(af: float) -> str:
    """annotation.py:316-331 (_format_af)."""
    s = f"{af:.10f}".rstrip("0").rstrip(".")
    decimal_digits = len(s.split(".")) if "." in s else 0
    return f"{af:.6e}" if decimal_digits > 3 else str(round(af, 6))


class AmbiguousVariants(ValueError):
    """Two variants at one (normalised) position of a haplotype: the reference's result depends on
    its set iteration order there (annotation.py:96, 129-160), so there is nothing to be bit-exact
    with; the seam hands such a list to the reference's own functions."""


def annotate_table(table: Dict[str, np.ndarray], res, batch, haplotypes, right: bool,
                   vt: Optional[marshal.VariantTable] = None, debug: bool = True) -> Dict[str, list]:  # fmt: skip
    """Columns of annotation.annotate_guides' first four steps for every row of `table`
    (the dict `search_table` returned, rows in emission order): `variants`, `afs_str`,
    `sequence` (after reverse_guides), `right` (after it) and `gc` -- the values the
    reference leaves in Guide.variants / .afs_str / .sequence / .right / .gc."""
    from . import _cabi

    if not getattr(batch, "has_variants", False):
        vt = vt or marshal.variant_table(haplotypes)
        if vt.ambiguous:
            raise AmbiguousVariants("a haplotype carries two variants at one position")
        batch.set_variants(vt)
    elif vt is None:
        vt = marshal.variant_table(haplotypes)
    try:
        ann = res.annotate(batch)
    except _cabi.HawkLibraryError as e:
        if e.code == _cabi.HAWK_EASSERT:
            raise AssertionError(str(e)) from e  # annotation.py:191
        raise
    n = len(table["hap"])
    off, idx = ann["gv_off"], ann["gv_idx"]
    hap, strand = table["hap"], table["strand"]
    variants: List[str] = []
    afs_str: List[str] = []
    for i in range(n):
        h = haplotypes[int(hap[i])]
        if h.variants == "NA":  # _is_reference_guide, annotation.py:104-126
            variants.append("NA")
            afs_str.append("NA")
            continue
        ids = sorted(vt.ids[int(hap[i])][j] for j in idx[off[i] : off[i + 1]])
        v = ",".join(ids)
        variants.append(v)
        vals = [format_af(h.afs[x]) if str(h.afs[x]) != "nan" else "NA" for x in v.split(",")]
        afs_str.append("NA" if not vals or (len(set(vals)) == 1 and vals[0] == "NA") else ",".join(vals))  # guide.py:311-328
    seq = [bytes(r).decode("ascii") for r in ann["rc_text"]]
    num, den = ann["gc_num"], ann["gc_den"]
    if n and int(den.min()) == 0:
        # gc_fraction returns the int 0 and Guide.gc refuses it (guide.py:598-618, annotation.py:531-538)
        exception_handler(error_class("CrisprHawkGcContentError"), "GC content calculation failed", 65, debug, None)
    gc = [str(int(a) / int(b)) for a, b in zip(num, den)]
    rp = [((not right) if s == 1 else bool(right)) != (s == 1) for s in strand.tolist()]
    return {"variants": variants, "afs_str": afs_str, "sequence": seq, "right": rp, "gc": gc}


def report_groups(table, res, haplotypes):
    """Row groups of reports._collapse_report_entries (reports.py:958-1008) from the resident
    table: `hawk_result_collapse`, then the split into index arrays (report_rows.groups_of)."""
    from . import report_rows

    is_ref = np.array([h.samples == "REF" for h in haplotypes], dtype=np.uint8)
    perm, head, collision = res.collapse(is_ref)
    key_of = None
    if collision:
        hap, strand, start, stop = table["hap"], table["strand"], table["start"], table["stop"]
        core = res.table()["text"][:, 10 : 10 + res.window - 20]  # GUIDESEQPAD either side

        def key_of(i):
            return (int(start[i]), int(stop[i]), int(strand[i]), int(is_ref[hap[i]]), core[i].tobytes())

    return report_rows.groups_of(perm, head, collision, key_of)


# --------------------------------------------------------------------------- drop-in seam (N2)
# Mirrors of the four per-guide loops annotation.annotate_guides runs right after search()
# (annotation.py:563-572), same names, signatures and return values. On a list that came from
# crispr_hawk_b200.search (a GuideList with its device-resident table) the first call computes
# every column on the device and the four functions only assign them to the Guide objects; any
# other list goes to the reference's own function (install() keeps it), untouched.
_reference = {}


def _columns(guides, debug: bool):
    link = getattr(guides, "hawk", None)
    if link is None:
        return None
    if "cols" not in link:
        if link.get("res") is None or not getattr(link["res"], "handle", None):
            return None
        try:
            link["cols"] = annotate_table(link["table"], link["res"], link["batch"], link["haplotypes"], link["right"], debug=debug)
        except AmbiguousVariants:
            link["cols"] = None  # sticky: every later step of this list goes to the reference too
            link["res"].close()
            link["res"] = None
            return None
        link["groups"] = report_groups(link["table"], link["res"], link["haplotypes"])  # for the report's row collapse
        from . import scoring

        link["cfdon"] = scoring.cfdon_column(link, debug)  # N4, for scoring.cfdon_score (None: not applicable)
        link["kmers"] = scoring.kmer_columns(link)  # N4, the learned scorers' input strings (scoring.py:50-84)
        link["res"].close()  # the table has served its purpose: release the device memory
        link["res"] = None
    if link["cols"] is None:
        return None
    return link["cols"], link["order"]


def _fallback(name, *args):
    fn = _reference.get(name)
    if fn is None:
        raise RuntimeError(f"crispr_hawk_b200.annotation.{name}: not a crispr_hawk_b200 guide list and no reference "
                           "implementation installed (there is no CPU path here)")  # fmt: skip
    return fn(*args)


def _stage(guides, cols, order, apply):
    """Run one per-guide step: a lazy GuideList takes it as a stage (objects that exist are
    updated now, the others when they are built); a plain list is walked as the reference does."""
    if hasattr(guides, "add_stage"):
        guides.add_stage(lambda g, k: apply(g, int(order[k])))
    else:
        for g, i in zip(guides, order.tolist()):
            apply(g, i)
    return guides


def _annotate_variants(guides, verbosity: int, debug: bool):
    """annotation.py:284-315."""
    got = _columns(guides, debug)
    if got is None:
        return _fallback("_annotate_variants", guides, verbosity, debug)
    cols, order = got

    def apply(g, i):
        g.variants = cols["variants"][i]

    return _stage(guides, cols, order, apply)


def annotate_variants_afs(guides, verbosity: int):
    """annotation.py:334-365."""
    got = _columns(guides, True)
    if got is None:
        return _fallback("annotate_variants_afs", guides, verbosity)
    cols, order = got

    def apply(g, i):
        g.afs_str = cols["afs_str"][i].split(",")  # the setter joins again (guide.py:311-328)

    return _stage(guides, cols, order, apply)


def reverse_guides(guides, verbosity: int):
    """annotation.py:27-51: the reverse-complemented text comes from the device."""
    got = _columns(guides, True)
    if got is None:
        return _fallback("reverse_guides", guides, verbosity)
    cols, order = got

    def apply(g, i):
        if g.strand == 1:
            g._sequence = cols["sequence"][i]
            g._right = cols["right"][i]
            (getattr(g, "_compute_pamguide_sequences", None) or g._split)()  # guide.py:255

    return _stage(guides, cols, order, apply)


def gc_content(guides, verbosity: int, debug: bool):
    """annotation.py:513-541."""
    got = _columns(guides, debug)
    if got is None:
        return _fallback("gc_content", guides, verbosity, debug)
    cols, order = got

    def apply(g, i):
        g.gc = float(cols["gc"][i])  # the setter stores str(value) (guide.py:598-618)

    return _stage(guides, cols, order, apply)


SEAM = ("_annotate_variants", "annotate_variants_afs", "reverse_guides", "gc_content")
