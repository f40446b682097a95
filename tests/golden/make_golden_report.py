"""Golden rows for the N2 row collapse: the reference's own report pipeline (search ->
annotate_guides -> reports._construct_report -> reports._collapse_report_entries,
reports.py:476-610, 958-1008) for a score-free Cas system (SaCas9 NNGRRT: no scorer branch,
scoring.py:845-857), run UNMODIFIED in the build container:

    PYTHONHASHSEED=0 python tests/golden/make_golden_report.py

Stores the haplotypes (like make_golden.py) and the collapsed table's rows.
"""

from __future__ import annotations

import gzip
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402
from tests.golden.make_golden import hap_to_json  # noqa: E402
from tests.synth_cases import make_case  # noqa: E402
from tests.test_real_driver import load_driver  # noqa: E402

COLUMNS = ["chr", "start", "stop", "sgRNA_sequence", "pam", "pam_class", "strand", "gc_content", "origin", "samples",
           "variant_id", "af", "target", "haplotype_id"]  # fmt: skip


def cases():
    out = [make_case(1, bed_len=5000, n_sites=40, n_samples=10, indel_frac=0.3, max_indel=5, multiallelic_frac=0.0,
                     pam="NNGRRT", guidelen=21, right=False, phased=True, name="C1_sacas9")]  # fmt: skip
    for k, (pam, g, right) in enumerate([("NNGRRT", 21, False), ("NNNNGATT", 22, False), ("NNGRRT", 21, False)]):
        out.append(make_case(300 + k, bed_len=900, n_sites=30, n_samples=6, phased=True, pam=pam, guidelen=g, right=right,
                             indel_frac=0.4, name=f"report_{k}_{pam}"))  # fmt: skip
    return out


def run(case):
    drv = load_driver()
    import crisprhawk.reports as R

    region, haps = refshim.build_case(case.ref_text, case.bed_start, case.bed_stop, case.vcf_lines, case.samples, True)
    args = types.SimpleNamespace(guidelen=case.guidelen, right=case.right, verbosity=0, debug=True, annotations=[],
                                 gene_annotations=[], annotation_colnames=[], gene_annotation_colnames=[],
                                 estimate_offtargets=False, compute_elevation=False)  # fmt: skip
    pam = drv.encode_pam(case.pam, case.right, 0, True)
    bits = drv.encode_haplotypes({region: haps}, args)
    guides = drv.guides_search(pam, {region: haps}, bits, True, True, args)
    guides = drv.annotate_guides(guides, args)
    rep = R._construct_report(guides, pam, [], [], [], [], False, False)[region]
    n_before = len(rep)
    col = R._collapse_report_entries(rep, pam, [], [], False)
    rows = [[(int(r[c]) if c in ("start", "stop") else str(r[c])) for c in COLUMNS] for _, r in col.iterrows()]
    return {"name": case.name, "pam": case.pam, "guidelen": case.guidelen, "right": case.right, "contig": case.contig,
            "region_start": region.start, "region_stop": region.stop, "target": str(region.coordinates),
            "haps": [hap_to_json(h) for h in haps], "n_guides": n_before, "columns": COLUMNS, "rows": rows}  # fmt: skip


if __name__ == "__main__":
    os.environ.setdefault("PYTHONHASHSEED", "0")
    out = [run(c) for c in cases()]
    path = os.path.join(HERE, "report.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps(out, separators=(",", ":")).encode())
    print(f"report: {len(out)} cases, {sum(c['n_guides'] for c in out)} guides -> {sum(len(c['rows']) for c in out)} rows, "
          f"{os.path.getsize(path) / 1e3:.0f} kB")
