"""Golden scorer inputs (N4, second half): the reference's own `scoring._extract_guide_sequences`
and `_extract_guide_sequences_sgdesigner` (scoring.py:50-84) and DeepCpf1's `preprocess`
(scores/deepCpf1/seqdeepcpf1.py:71-92), run UNMODIFIED in the build container on the guides the
genuine driver leaves after `guides_search` + `annotate_guides` (i.e. with strand-1 guides already
reverse-complemented, the state the scorers see them in).

    PYTHONHASHSEED=0 python tests/golden/make_golden_features.py

Stores, per case, the haplotypes and per guide in list order [start, strand, hapid, 4-lead string,
0-lead string]; of the one-hot tensor its shape, a SHA-256 of its float32 bytes and its first row
in full -- or "KeyError" where `preprocess` raises it (a letter other than A, C, G, T).
`seqdeepcpf1.py` imports h5py (absent here) only for its weight loader: an empty stub stands in.
"""

from __future__ import annotations

import gzip
import hashlib
import importlib.util
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden.make_golden import hap_to_json  # noqa: E402
from tests.synth_cases import config1_cases, make_case  # noqa: E402
from tests.test_real_driver import load_driver, load_scoring, run_driver  # noqa: E402

FEATURE_CASES = [config1_cases()[0]] + [
    make_case(150 + k, phased=ph, pam=p, guidelen=g, right=r, bed_len=900, n_sites=40, n_samples=5, indel_frac=0.3)
    for k, (p, g, r, ph) in enumerate([("NGG", 20, False, True), ("TTTV", 23, True, True), ("TTTV", 23, False, True),
                                       ("NNGRRT", 21, False, True), ("NGG", 18, True, True), ("NGG", 20, False, False)])
]  # fmt: skip


def with_n_bases(case, every=97):
    """The same case over a reference with a few N bases, placed where no VCF record's REF allele
    lies (variant.py / haplotype.py:206-210 check it): `preprocess` raises KeyError on such guides."""
    import dataclasses

    g0 = case.bed_start - 100
    taken = set()
    for line in case.vcf_lines:
        f = line.split("\t")
        taken.update(range(int(f[1]) - g0 - 1, int(f[1]) - g0 + len(f[3]) + 1))
    text = list(case.ref_text)
    for i in range(130, len(text) - 130, every):
        if i not in taken:
            text[i] = "N"
    return dataclasses.replace(case, name=case.name + "_N", ref_text="".join(text))


FEATURE_CASES.append(with_n_bases(FEATURE_CASES[1]))


def load_preprocess():
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    import crisprhawk

    path = os.path.join(os.path.dirname(crisprhawk.__file__), "scores", "deepCpf1", "seqdeepcpf1.py")
    spec = importlib.util.spec_from_file_location("ref_seqdeepcpf1", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.preprocess


def run(case):
    scoring, _ = load_scoring()
    drv = load_driver()
    preprocess = load_preprocess()
    region, _, guides = run_driver(drv, case)
    from oracle import refshim

    _, haps = refshim.build_case(case.ref_text, case.bed_start, case.bed_stop, case.vcf_lines, case.samples, case.phased)
    k4 = scoring._extract_guide_sequences(guides)
    k0 = scoring._extract_guide_sequences_sgdesigner(guides)
    try:
        t = preprocess(k4).numpy()
        onehot = {"shape": list(t.shape), "sha256": hashlib.sha256(t.tobytes()).hexdigest(), "row0": t[0].astype(int).tolist()}
    except KeyError:
        onehot = "KeyError"
    rows = [[g.start, g.strand, g.hapid, a, b] for g, a, b in zip(guides, k4, k0)]
    return {"name": case.name, "pam": case.pam, "guidelen": case.guidelen, "right": case.right, "phased": case.phased,
            "contig": case.contig, "region_start": region.start, "region_stop": region.stop,
            "haps": [hap_to_json(h) for h in haps], "rows": rows, "onehot": onehot}  # fmt: skip


if __name__ == "__main__":
    out = {"cases": [run(c) for c in FEATURE_CASES]}
    path = os.path.join(HERE, "features.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps(out, separators=(",", ":")).encode())
    n = sum(len(c["rows"]) for c in out["cases"])
    ke = sum(c["onehot"] == "KeyError" for c in out["cases"])
    print(f"features: {len(out['cases'])} cases, {n} guides, {ke} case(s) where preprocess raises KeyError, {os.path.getsize(path) / 1e3:.0f} kB")
