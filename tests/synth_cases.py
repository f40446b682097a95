"""Seeded generator of small FASTA+VCF-like cases (test support, CPU only).

It produces inputs the *reference's* haplotype builder accepts (SURVEY.md
Appendix B): ACGT-only reference, left-anchored indels, no overlapping indel
spans, variants kept away from the padded region ends, one AF per ALT allele.
The cases are turned into haplotypes either by the live reference
(`oracle/refshim.build_case`, build container only) or loaded back from the
golden fixtures.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

BASES = "ACGT"


@dataclass
class SmallCase:
    name: str
    ref_text: str  # padded FASTA slice, 1-based inclusive [bed_start-100, bed_stop+100]
    bed_start: int
    bed_stop: int
    vcf_lines: List[str]
    samples: List[str]
    phased: bool
    pam: str
    guidelen: int
    right: bool
    contig: str = "chr1"

    @property
    def variants_present(self) -> bool:
        return bool(self.vcf_lines)


def _gt(rng, n_alt, af, phased):
    sep = "|" if phased else "/"
    alleles = []
    for _ in range(2):
        if rng.random() < af:
            alleles.append(str(int(rng.integers(1, n_alt + 1))))
        else:
            alleles.append("0")
    if not phased and n_alt == 1 and alleles == ["1", "1"] and rng.random() < 0.5:
        alleles = ["0", "1"]
    return sep.join(alleles)


def make_case(
    seed: int,
    bed_len: int = 600,
    n_sites: int = 12,
    n_samples: int = 4,
    phased: bool = True,
    pam: str = "NGG",
    guidelen: int = 20,
    right: bool = False,
    indel_frac: float = 0.3,
    multiallelic_frac: float = 0.15,
    max_indel: int = 4,
    bed_start: int = 10001,
    edge_margin: int = 24,
    name: str = "",
) -> SmallCase:
    rng = np.random.default_rng(seed)
    pad = 100
    total = bed_len + 2 * pad
    ref = "".join(BASES[i] for i in rng.integers(0, 4, total))
    g0 = bed_start - pad  # genomic coordinate of ref[0]
    bed_stop = bed_start + bed_len - 1
    samples = [f"S{i + 1}" for i in range(n_samples)]
    lines = []
    # candidate sites anywhere in the padded region (variants in the padding ARE
    # applied by the reference, variant.py:791-793), away from the ends
    lo, hi = g0 + edge_margin, g0 + total - 1 - edge_margin - max_indel
    taken_until = -1
    positions = np.sort(rng.choice(np.arange(lo, hi), size=min(n_sites, hi - lo), replace=False))
    for gpos in positions:
        gpos = int(gpos)
        if gpos <= taken_until:
            continue
        i = gpos - g0
        refb = ref[i]
        r = rng.random()
        af = float(rng.choice([0.5, 0.25, 0.125, 0.05]))
        if r < indel_frac / 2:  # insertion
            k = int(rng.integers(1, max_indel + 1))
            alt = refb + "".join(BASES[j] for j in rng.integers(0, 4, k))
            refa, alts, afs = refb, [alt], [af]
            taken_until = gpos  # next site may be adjacent
        elif r < indel_frac:  # deletion
            k = int(rng.integers(1, max_indel + 1))
            refa, alts, afs = ref[i : i + k + 1], [refb], [af]
            taken_until = gpos + k  # nothing may be anchored inside the deleted span
        else:  # SNV, sometimes multi-allelic
            others = [b for b in BASES if b != refb]
            if rng.random() < multiallelic_frac:
                a = list(rng.choice(others, size=2, replace=False))
                refa, alts, afs = refb, [str(x) for x in a], [af, af / 2]
            else:
                refa, alts, afs = refb, [str(rng.choice(others))], [af]
            taken_until = gpos
        carr = float(rng.choice([0.2, 0.4, 0.7]))
        gts = [_gt(rng, len(alts), carr, phased) for _ in samples]
        if all(set(g.replace("|", "/").split("/")) == {"0"} for g in gts):
            gts[int(rng.integers(0, n_samples))] = ("1|0" if phased else "0/1")
        lines.append(
            "\t".join(
                [
                    "chr1", str(gpos), ".", refa, ",".join(alts), ".", "PASS",
                    "AF=" + ",".join(f"{x:g}" for x in afs), "GT",
                ]
                + gts
            )  # fmt: skip
        )
    return SmallCase(
        name=name or f"seed{seed}_{'ph' if phased else 'un'}_{pam}_{guidelen}_{'R' if right else 'L'}",
        ref_text=ref,
        bed_start=bed_start,
        bed_stop=bed_stop,
        vcf_lines=lines,
        samples=samples,
        phased=phased,
        pam=pam,
        guidelen=guidelen,
        right=right,
    )


KAT_REF = (
    "AGACTTTCAAAGATATGCTGGGTAGAGGTCGAGGTTATTATTTGTTACCAATTCTCATTGTGTTTCGGAA"
    "CTTGCGTTTTAGGTATGTCTTAGTGACTCTAAATACCAAGGCAGTCCTCGATCCGTTCCTAATAAGGAAT"
    "GGTGATTCCCTGTCATACCAATCTACCCCCTGTTATGCGCGTTTGTCGTTAGACCAATGTCAGCGCAGCG"
    "GCAGATCAAGCAGGAGGCGGAATGTAAACAGAAGGTATGCTTAGGTGGATAGGGAGTGAGCAACAAACGG"
)
_KAT_VCF = [
    "chr1\t1020\t.\tG\tA\t.\tPASS\tAF=0.25\tGT\t0|1\t1|1",
    "chr1\t1040\t.\tT\tTGG\t.\tPASS\tAF=0.5\tGT\t1|0\t0|0",
    "chr1\t1060\t.\tAAT\tA\t.\tPASS\tAF=0.125\tGT\t0|0\t0|1",
]


def kat_cases() -> List[SmallCase]:
    """The four known-answer cases of SURVEY.md Appendix A."""
    s2 = ["S1", "S2"]
    return [
        SmallCase("KAT1", KAT_REF, 1001, 1080, [], [], False, "NGG", 20, False),
        SmallCase("KAT2", KAT_REF, 1001, 1080, [], [], False, "TTN", 23, True),
        SmallCase("KAT3", KAT_REF, 1001, 1080, list(_KAT_VCF), s2, True, "NGG", 20, False),
        SmallCase(
            "KAT4", KAT_REF, 1001, 1080, [l.replace("|", "/") for l in _KAT_VCF], s2, False,
            "NGG", 20, False,
        ),  # fmt: skip
    ]


PAM_SETUPS = [
    ("NGG", 20, False),
    ("TTTV", 23, True),
    ("NNGRRT", 21, False),
    ("TTN", 23, True),
]


def random_cases(n_seeds: int, seed0: int = 100, **kw) -> List[SmallCase]:
    out = []
    for s in range(seed0, seed0 + n_seeds):
        for phased in (True, False):
            for pam, g, right in PAM_SETUPS:
                out.append(make_case(s, phased=phased, pam=pam, guidelen=g, right=right, **kw))
    return out


def config1_cases() -> List[SmallCase]:
    """5 kb, 40 sites (70/15/15 SNV/ins/del, 1-5 bp), 10 samples, seed 1."""
    kw = dict(bed_len=5000, n_sites=40, n_samples=10, indel_frac=0.3, max_indel=5,
              multiallelic_frac=0.0, pam="NGG", guidelen=20, right=False)  # fmt: skip
    return [
        make_case(1, phased=True, name="C1_phased", **kw),
        make_case(1, phased=False, name="C1_unphased", **kw),
    ]


GNOMAD_POPS = ["afr", "amr", "asj", "eas", "fin", "mid", "nfe", "sas", "ami", "remaining"]


def config4_cases() -> List[SmallCase]:
    """BASELINE config 4 in miniature: SaCas9 NNGRRT / 21 nt over gnomAD-density variants
    (one site per ~8 bp, 88 % SNV / 12 % indel, 5 % multi-allelic), unphased, the converter's
    10 population pseudo-samples (converter.py:19-37). Slices this small are all the
    reference can resolve in reasonable time."""
    out = []
    for k, (seed, bed_len) in enumerate([(4, 1200), (41, 1600), (42, 800)]):
        c = make_case(seed, bed_len=bed_len, n_sites=bed_len // 8, n_samples=10, phased=False, pam="NNGRRT",
                      guidelen=21, right=False, indel_frac=0.12, multiallelic_frac=0.05, max_indel=5,
                      name=f"C4_gnomad_{k}")  # fmt: skip
        c.samples[:] = GNOMAD_POPS
        out.append(c)
    return out
