"""Generate golden input/output vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case it builds haplotypes with the reference's own haplotype builder,
runs the reference's `encode` + `pam_search` + `search`, and stores inputs and
outputs as plain JSON (gzip). The GPU box has no /root/reference; the `-m gpu`
parity tests read these files instead.
"""

from __future__ import annotations

import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402
from tests.synth_cases import config1_cases, config4_cases, kat_cases, make_case, random_cases  # noqa: E402


def hap_to_json(h):
    pm = h.posmap
    return {
        "seq": h.sequence.sequence,
        "posmap": [pm[i] for i in range(len(pm))],
        "start": h.start,
        "stop": h.stop,
        "samples": h.samples,
        "variants": h.variants,
        "afs": {k: (None if v != v else v) for k, v in h.afs.items()},
        "variant_alleles": {str(k): [list(e) for e in v] for k, v in h.variant_alleles.items()},
        "id": h.id,
    }


def guide_to_json(g):
    pm = g.posmap
    return [
        g.start, g.stop, g.strand, g.sequence, bool(g.right), g.samples, g.variants,
        g.hapid, [pm[j] for j in range(len(pm))],
    ]  # fmt: skip


def run_case(case):
    ref = refshim.load()
    region, haps = refshim.build_case(
        case.ref_text, case.bed_start, case.bed_stop, case.vcf_lines, case.samples, case.phased
    )
    pam, bits, guides = refshim.run_search(
        region, haps, case.pam, case.guidelen, case.right, case.variants_present, case.phased
    )
    hits = ref.search_guides.pam_search(pam, region, haps, bits, 0, True)
    bounds = [
        list(ref.search_guides.compute_scan_start_stop(h, region.start, region.stop, len(pam)))
        for h in haps
    ]
    return {
        "name": case.name,
        "pam": case.pam,
        "guidelen": case.guidelen,
        "right": case.right,
        "phased": case.phased,
        "variants_present": case.variants_present,
        "contig": case.contig,
        "region_start": region.start,
        "region_stop": region.stop,
        "pam_bits": pam.bits,
        "pam_bitsrc": pam.bitsrc,
        "pam_rc": pam.pamrc,
        "cas_system": pam.cas_system,
        "haps": [hap_to_json(h) for h in haps],
        "scan_bounds": bounds,
        "pam_hits": [[list(f), list(r)] for f, r in hits],
        "guides": [guide_to_json(g) for g in guides],
    }


def dump(name, cases):
    out = [run_case(c) for c in cases]
    path = os.path.join(HERE, name + ".json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps(out, separators=(",", ":")).encode())
    ng = sum(len(c["guides"]) for c in out)
    print(f"{name}: {len(out)} cases, {ng} guides, {os.path.getsize(path) / 1e3:.0f} kB")


def edge_cases():
    """Targeted shapes: dense variants, long indels, degenerate PAMs, long/short guides."""
    out = []
    for s in range(5):
        out.append(make_case(500 + s, bed_len=300, n_sites=40, n_samples=3, phased=True,
                             pam="NGG", guidelen=20, right=False, indel_frac=0.5, max_indel=8,
                             name=f"dense_ph_{s}"))  # fmt: skip
        out.append(make_case(520 + s, bed_len=300, n_sites=25, n_samples=3, phased=False,
                             pam="NNGRRT", guidelen=21, right=False, indel_frac=0.4, max_indel=6,
                             name=f"dense_un_{s}"))  # fmt: skip
    out.append(make_case(540, bed_len=400, phased=True, pam="NNNNGATT", guidelen=22, right=False, name="nme_ph"))
    out.append(make_case(541, bed_len=400, phased=False, pam="NNNNGATT", guidelen=22, right=False, name="nme_un"))
    out.append(make_case(542, bed_len=400, phased=True, pam="TTTV", guidelen=40, right=True, name="long_guide_ph"))
    out.append(make_case(543, bed_len=400, phased=False, pam="NGG", guidelen=33, right=False, name="g33_un"))
    out.append(make_case(544, bed_len=400, phased=True, pam="NRG", guidelen=8, right=False, name="short_guide_ph"))
    out.append(make_case(545, bed_len=400, phased=True, pam="N", guidelen=20, right=False, n_sites=6, name="pam_N_ph"))
    out.append(make_case(546, bed_len=50, phased=True, pam="NGG", guidelen=20, right=False, n_sites=4, name="tiny_bed_ph"))
    out.append(make_case(547, bed_len=50, phased=False, pam="NGG", guidelen=20, right=True, n_sites=4, name="tiny_bed_un_R"))
    out.append(make_case(548, bed_len=400, phased=True, pam="NGG", guidelen=20, right=False, n_sites=0, name="novariants"))
    out.append(make_case(549, bed_len=1200, phased=True, pam="YTTV", guidelen=23, right=True, n_sites=30, n_samples=8, name="yttv_ph"))
    out.append(make_case(550, bed_len=1200, phased=False, pam="NGK", guidelen=20, right=False, n_sites=30, n_samples=8, name="ngk_un"))
    return out


if __name__ == "__main__":
    os.environ.setdefault("PYTHONHASHSEED", "0")
    if "--only-config4" in sys.argv:
        dump("config4", config4_cases())
        sys.exit(0)
    dump("config4", config4_cases())
    dump("kat", kat_cases())
    dump("config1", config1_cases())
    dump("random", random_cases(5))
    dump("edge", edge_cases())
