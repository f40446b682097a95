"""Golden vectors for BASELINE config 4 (unphased, gnomAD density) at the slice size SURVEY.md
8(d) names: 8 x 5 kb, SaCas9 NNGRRT / 21 nt, built and searched by the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    PYTHONHASHSEED=0 python tests/golden/make_golden_c4.py

The slices are cohorts of crispr_hawk_b200.synth_unphased (the generator behind bench.py's c4
workload). For each one the script (1) feeds the cohort's VCF lines to the reference's own
haplotype builder and checks that the generator derived the same haplotypes -- sequence,
position map, variant_alleles, bounds -- (2) runs the reference's encode + search, and
(3) stores what a test needs to reproduce the run without the reference: the seed, the order
in which the reference listed the haplotypes (it iterates Python sets, so the order inside an
indel's group is the interpreter's), the number of guides, a SHA-256 over the canonical text of
the whole guide list and every 97th guide in full. The guide tables themselves would be ~20 MB.
"""

from __future__ import annotations

import gzip
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from crispr_hawk_b200 import synth_unphased as SU  # noqa: E402
from oracle import refshim  # noqa: E402

SLICES = [dict(seed=400 + i, bed_len=5000) for i in range(8)]
PAM, GUIDELEN, RIGHT = "NNGRRT", 21, False
SAMPLE_EVERY = 97


def canon_hap(h):
    pm = h.posmap
    return (h.sequence.sequence, tuple(pm[i] for i in range(len(pm))), h.start, h.stop,
            tuple(sorted((k, tuple((e[0], e[1], e[2]) for e in v)) for k, v in h.variant_alleles.items())))  # fmt: skip


def guide_line(start, stop, strand, sequence, right, hap_index):
    return f"{start}\t{stop}\t{strand}\t{sequence}\t{int(bool(right))}\t{hap_index}\n"


def run_slice(spec):
    c = SU.make_unphased_cohort(spec["bed_len"], spec["seed"])
    mine = SU.unphased_haplotypes(c)
    lines, samples = SU.to_vcf_lines(c)
    region, ref_haps = refshim.build_case(c.ref.tobytes().decode(), c.bed_start, c.bed_stop, lines, samples, False)
    index_of = {}
    for i, h in enumerate(mine):
        index_of.setdefault(canon_hap(h), []).append(i)
    order = []
    for h in ref_haps:
        order.append(index_of[canon_hap(h)].pop(0))  # KeyError / IndexError = generator differs from the reference
    assert sorted(order) == list(range(len(mine)))
    for i, h in enumerate(ref_haps):
        h.id = f"h{order[i]}"
    pam, bits, guides = refshim.run_search(region, ref_haps, PAM, GUIDELEN, RIGHT, True, False)
    sha = hashlib.sha256()
    sample = []
    for k, g in enumerate(guides):
        row = (g.start, g.stop, g.strand, g.sequence, bool(g.right), int(g.hapid[1:]))
        sha.update(guide_line(*row).encode())
        if k % SAMPLE_EVERY == 0:
            sample.append([k] + list(row))
    return {"seed": spec["seed"], "bed_len": spec["bed_len"], "pam": PAM, "guidelen": GUIDELEN, "right": RIGHT,
            "n_haps": len(ref_haps), "hap_order": order, "n_guides": len(guides), "sha256": sha.hexdigest(),
            "sample_every": SAMPLE_EVERY, "sample": sample}  # fmt: skip


if __name__ == "__main__":
    out = [run_slice(s) for s in SLICES]
    path = os.path.join(HERE, "config4_slices.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps(out, separators=(",", ":")).encode())
    print(f"config4_slices: {len(out)} slices, {sum(s['n_guides'] for s in out)} guides, "
          f"{sum(s['n_haps'] for s in out)} haplotypes, {os.path.getsize(path) / 1e3:.0f} kB")
