"""N2, second half: the row collapse of the report -- reports._collapse_report_entries
(reports.py:958-1008) with the aggregations of reports.py:767-858, 912-955.

The reference groups the per-guide DataFrame with pandas on (chr, start, stop, sgRNA_sequence,
pam, strand, [scores,] gc_content, origin). Every one of those columns but (start, stop,
strand, origin) is a function of the guide + PAM text of the row, so the groups are the classes
of (start, stop, strand, origin, text) -- which the device computes over the resident table
(`hawk_result_collapse`: hash, two radix sorts, head flags). What is left here is what pandas
would do per group: take the key columns of the first row, join the sample sets and haplotype
ids, and order the groups by the key columns."""

from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Sequence

import numpy as np

IUPAC_SETS = {"A": "A", "C": "C", "G": "G", "T": "T", "R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT", "M": "AC",
              "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT"}  # fmt: skip

# the groupby keys (reports.py:983-989 for a Cas system without a scorer branch) and the aggregated
# fields (reports.py:930-939); pandas emits the keys first -- minus `origin`, which is also an
# aggregated field and appears there
GROUP_COLS = ["chr", "start", "stop", "sgRNA_sequence", "pam", "strand", "gc_content", "origin"]
AGG_COLS = ["pam_class", "origin", "samples", "variant_id", "af", "target", "haplotype_id"]
COLUMNS = [c for c in GROUP_COLS if c != "origin"] + AGG_COLS


def pam_class(pam_text: str) -> str:
    """reports.py:64-77 (_compute_pam_class): NGG -> [ACGT]GG."""
    return "".join(nt if nt in "ACGT" else f"[{IUPAC_SETS[nt]}]" for nt in pam_text)


def polish_samples_phased(samples: str) -> str:
    """reports.py:767-790: per sample, the element-wise max of its phased genotypes, samples in
    first-appearance order."""
    if "|" not in samples:
        return samples
    m: Dict[str, List[int]] = defaultdict(lambda: [0, 0])
    for e in samples.split(","):
        sample, gt = e.split(":")
        a1, a2 = map(int, gt.split("|"))
        m[sample][0] = max(m[sample][0], a1)
        m[sample][1] = max(m[sample][1], a2)
    return ",".join(f"{s}:{a}|{b}" for s, (a, b) in m.items())


def collapse_samples(samples: Sequence[str]) -> str:
    """reports.py:793-810."""
    return "" if not len(samples) else polish_samples_phased(",".join(sorted(set(",".join(samples).split(",")))))


def collapse_haplotype_ids(hapids: Sequence[str]) -> str:
    """reports.py:845-858."""
    return "" if not len(hapids) else ",".join(sorted(set(",".join(hapids).split(","))))


def check_variant_ids(variant_ids: Sequence[str]) -> str:
    """reports.py:828-842: the sorted id set of the group. (Where the rows of a group carry
    different sets the reference pops one in hash order; this takes the first row's.)"""
    v = variant_ids[0]
    return ",".join(sorted(set(v.split(",")))) if v else ""


def groups_of(perm: np.ndarray, head: np.ndarray, collision: bool, key_of=None) -> List[np.ndarray]:
    """Table-row indices of every group (rows in emission order). After a hash collision equal
    keys may sit in two runs: those are joined on the exact key."""
    if len(perm) == 0:
        return []
    cuts = np.flatnonzero(head)
    out = np.split(perm.astype(np.int64), cuts[1:])
    if collision and key_of is not None:
        merged: Dict[tuple, List[np.ndarray]] = {}
        for g in out:
            merged.setdefault(key_of(int(g[0])), []).append(g)
        out = [np.sort(np.concatenate(v)) for v in merged.values()]
    return out


def split_core(sequence: str, right: bool, guidelen: int, pamlen: int, pad: int = 10):
    """guide.py:184-197: (guide, pam) of a padded window text."""
    core = sequence[pad : pad + guidelen + pamlen]
    return (core[pamlen:], core[:pamlen]) if right else (core[:guidelen], core[guidelen:])


def collapse_table(table, groups, cols, haplotypes, contig: str, target: str, pam_text: str, guidelen: int) -> Dict[str, list]:
    """The collapsed report of one region for the score-free column set (Cas systems without a
    scorer branch, scoring.py:845-857), as {column: values}, rows in the reference's order.
    `table` = the search table (emission order), `groups` = groups_of(...), `cols` = the
    annotation columns of crispr_hawk_b200.annotation.annotate_table."""
    pamlen = len(pam_text)
    pclass = pam_class(pam_text)
    hap, strand, start, stop = table["hap"], table["strand"], table["start"], table["stop"]
    rows = []
    for g in groups:
        i = int(g[0])
        guide, pam = split_core(cols["sequence"][i], cols["right"][i], guidelen, pamlen)
        haps = [haplotypes[int(h)] for h in hap[g]]
        samples0 = haps[0].samples
        rows.append((contig, int(start[i]), int(stop[i]), guide, pam, "+" if strand[i] == 0 else "-", cols["gc"][i],
                     pclass, "ref" if samples0 == "REF" else "alt", collapse_samples([h.samples for h in haps]),
                     check_variant_ids([cols["variants"][int(k)] for k in g]), cols["afs_str"][i], target,
                     collapse_haplotype_ids([h.id for h in haps])))  # fmt: skip
    rows.sort(key=lambda r: r[:7] + (r[8],))  # pandas sorts the groupby keys lexicographically
    return {c: [r[k] for r in rows] for k, c in enumerate(COLUMNS)}


# --------------------------------------------------------------------------- drop-in seam
# Mirrors of reports._process_data (:476-531) and reports._collapse_report_entries (:958-1008).
# _process_data is the reference's own (it reads the Guide objects); the mirror only notes, on
# the DataFrame it returns, which device-computed groups its rows fall into. The collapse mirror
# uses them in place of the pandas groupby; without them (or with annotation / off-target
# columns, whose aggregations need the reference's own functions) the call goes to the reference.
_reference = {}


def _process_data(region, guides, *args):
    report = _reference["_process_data"](region, guides, *args)
    link = getattr(guides, "hawk", None)
    if link is not None and link.get("groups") is not None and len(report) == len(link["order"]):
        report.attrs["hawk_groups"] = (link["groups"], link["order"])
    return report


def _collapse_report_entries(report, pam, annotations, gene_annotations, estimate_offtargets):
    got = report.attrs.get("hawk_groups") if hasattr(report, "attrs") else None
    if got is None or annotations or gene_annotations or estimate_offtargets:
        fn = _reference.get("_collapse_report_entries")
        if fn is None:
            raise RuntimeError("crispr_hawk_b200.report_rows._collapse_report_entries: no device-computed groups on this "
                               "report and no reference implementation installed (there is no CPU path here)")  # fmt: skip
        return fn(report, pam, annotations, gene_annotations, estimate_offtargets)
    return collapse_frame(report, *got)


def collapse_frame(report, groups, order):
    """The collapsed DataFrame from device-computed groups: key columns of each group's first
    row, the aggregations of reports.py:912-955, rows ordered by the key columns."""
    where = np.empty(len(order), np.int64)  # table row -> report row (the guide list's order)
    where[order] = np.arange(len(order))
    agg = [c for c in AGG_COLS if c in report.columns]
    keys = [c for c in report.columns if c not in agg]
    if "score_elevationon" in keys:  # joins the keys last (reports.py:990-991)
        keys.remove("score_elevationon")
        keys.append("score_elevationon")
    col = {c: report[c].tolist() for c in ("samples", "variant_id", "haplotype_id")}
    firsts, samples, variants, hapids = [], [], [], []
    for g in groups:
        r = where[g]  # ascending within a group: a bucket's members keep emission order
        firsts.append(int(r[0]))
        samples.append(collapse_samples([col["samples"][k] for k in r]))
        variants.append(check_variant_ids([col["variant_id"][k] for k in r]))
        hapids.append(collapse_haplotype_ids([col["haplotype_id"][k] for k in r]))
    out = report.iloc[firsts].reset_index(drop=True)
    out["samples"], out["variant_id"], out["haplotype_id"] = samples, variants, hapids
    origin_at = keys.index("score_elevationon") if "score_elevationon" in keys else len(keys)
    out = out.sort_values(keys[:origin_at] + ["origin"] + keys[origin_at:], kind="stable").reset_index(drop=True)
    out.attrs = {}
    return out[keys + agg]


SEAM = ("_process_data", "_collapse_report_entries")
