"""Golden vectors for the post-search pure functions (N2): the UNMODIFIED reference's
`_annotate_variants`, `annotate_variants_afs`, `reverse_guides`, `gc_content`
(annotation.py:563-572) run on the guides its own `search()` returns, for every phased /
variant-free case of the golden set. Build container only (needs /root/reference):

    python tests/golden/make_golden_annot.py

Output: tests/golden/annot.json.gz = {case name: [[variants, afs_str, sequence, right, gc], ...]}
in the order of the case's guides, or {"error": exception class} when the reference itself
fails on the case."""

from __future__ import annotations

import gzip
import importlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import refshim  # noqa: E402
from tests.synth_cases import config1_cases, kat_cases, random_cases  # noqa: E402
from make_golden import edge_cases  # noqa: E402


def run(case):
    refshim.load()
    ann = importlib.import_module("crisprhawk.annotation")
    region, haps = refshim.build_case(case.ref_text, case.bed_start, case.bed_stop, case.vcf_lines, case.samples, case.phased)
    _, _, guides = refshim.run_search(region, haps, case.pam, case.guidelen, case.right, case.variants_present, case.phased)
    try:
        guides = ann._annotate_variants(guides, 0, True)
        guides = ann.annotate_variants_afs(guides, 0)
        guides = ann.reverse_guides(guides, 0)
        guides = ann.gc_content(guides, 0, True)
    except BaseException as e:  # the reference asserts / exits on some inputs
        return {"error": type(e).__name__}
    return [[g.variants, g.afs_str, g.sequence, bool(g.right), g.gc] for g in guides]


if __name__ == "__main__":
    os.environ.setdefault("PYTHONHASHSEED", "0")
    out = {}
    for c in kat_cases() + config1_cases() + random_cases(5) + edge_cases():
        if c.variants_present and not c.phased:
            continue
        out[c.name] = run(c)
    path = os.path.join(HERE, "annot.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps(out, separators=(",", ":")).encode())
    n = sum(len(v) for v in out.values() if isinstance(v, list))
    err = [k for k, v in out.items() if isinstance(v, dict)]
    print(f"annot: {len(out)} cases, {n} guides, errors: {err}, {os.path.getsize(path) / 1e3:.0f} kB")
