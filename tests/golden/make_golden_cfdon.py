"""Golden CFDon scores (N4): the reference's own `scoring.cfdon_score` ->
`scores.crisprhawk_scores.cfdon` -> `scores.cfdscore.cfdscore.compute_cfd`, run UNMODIFIED in the
build container on the annotated guides of five phased cases. The reference's model files
(`scores/cfdscore/models/*.pkl`) are not part of its source tree, so the factor tables are seeded
stand-ins with the reference's keys (`tests.test_real_driver.synthetic_cfd_dicts(7)`) -- the
arithmetic under test is the reference's, the factors are data.

    PYTHONHASHSEED=0 python tests/golden/make_golden_cfdon.py

Stores the haplotypes and, per guide in list order, the score string the Guide holds and the raw
float (hex) `cfdon` returned.
"""

from __future__ import annotations

import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden.make_golden import hap_to_json  # noqa: E402
from tests.test_real_driver import CFD_CASES, load_driver, load_scoring, run_driver, synthetic_cfd_dicts  # noqa: E402


def run(case):
    scoring, cs = load_scoring()
    drv = load_driver()
    tables = synthetic_cfd_dicts(7)
    cs.load_mismatch_pam_scores = lambda debug: tables
    region, _, guides = run_driver(drv, case)
    from oracle import refshim

    _, haps = refshim.build_case(case.ref_text, case.bed_start, case.bed_stop, case.vcf_lines, case.samples, True)
    raw = []
    for _, (ref, members) in scoring.group_guides_position(guides, True).items():
        raw.extend(cs.cfdon(ref, members, True))
    out = scoring.cfdon_score(guides, 0, True)
    assert len(raw) == len(out)
    rows = [[g.start, g.strand, g.hapid, g.cfdon_score, float(x).hex()] for g, x in zip(out, raw)]
    return {"name": case.name, "pam": case.pam, "guidelen": case.guidelen, "right": case.right, "contig": case.contig,
            "region_start": region.start, "region_stop": region.stop, "haps": [hap_to_json(h) for h in haps], "rows": rows}  # fmt: skip


if __name__ == "__main__":
    out = {"table_seed": 7, "cases": [run(c) for c in CFD_CASES]}
    path = os.path.join(HERE, "cfdon.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps(out, separators=(",", ":")).encode())
    n = sum(len(c["rows"]) for c in out["cases"])
    scored = sum(r[3] != "NA" for c in out["cases"] for r in c["rows"])
    print(f"cfdon: {len(out['cases'])} cases, {n} guides ({scored} with a REF guide at their key), {os.path.getsize(path) / 1e3:.0f} kB")
