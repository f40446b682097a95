"""TEST INFRASTRUCTURE ONLY -- loader for the *live* CRISPR-HAWK reference.

Only `tests/` and `tests/golden/make_golden.py` may import this module. It puts
the unmodified reference (``/root/reference/src``; never copied into this repo)
on ``sys.path`` and satisfies three I/O-only third-party imports that are not
installed in the build container:

* ``colorama``  -- imported by crisprhawk/utils.py:13, exception_handlers.py:11
* ``pysam``     -- imported by crisprhawk/sequence.py:13, bedfile.py:17, variant.py:13
* ``Bio``       -- imported by crisprhawk/annotation.py:21 (gc_fraction only)

None of the stubs touches the arithmetic of the hot path. ``/root/reference``
does not exist on the GPU box, so everything that uses this module is skipped
there (``available()`` is False).
"""

from __future__ import annotations

import os
import sys
import types

REFERENCE_SRC = os.environ.get("HAWK_REFERENCE_SRC", "/root/reference/src")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "crisprhawk"))


def _install_stubs() -> None:
    if "colorama" not in sys.modules:
        try:
            import colorama  # noqa: F401
        except ImportError:
            col = types.ModuleType("colorama")

            class _Codes:
                def __getattr__(self, name):
                    return ""

            col.Fore = _Codes()
            col.Back = _Codes()
            col.Style = _Codes()
            col.init = lambda *a, **k: None
            sys.modules["colorama"] = col
    if "pysam" not in sys.modules:
        try:
            import pysam  # noqa: F401
        except ImportError:
            pysam = types.ModuleType("pysam")
            for name in (
                "TabixFile",
                "FastaFile",
                "VariantFile",
                "VariantHeader",
                "VariantRecord",
            ):
                setattr(pysam, name, object)
            pysam.tabix_index = lambda *a, **k: None
            pysam.faidx = lambda *a, **k: None
            pysam.utils = types.ModuleType("pysam.utils")
            pysam.utils.SamtoolsError = Exception
            sys.modules["pysam"] = pysam
            sys.modules["pysam.utils"] = pysam.utils
    if "Bio" not in sys.modules:
        try:
            import Bio  # noqa: F401
        except ImportError:
            bio = types.ModuleType("Bio")
            su = types.ModuleType("Bio.SeqUtils")

            def gc_fraction(seq, ambiguous="remove"):
                s = str(seq).upper()
                gc = sum(s.count(c) for c in "CGS")
                n = gc + sum(s.count(c) for c in "ATW")
                return gc / n if n else 0

            su.gc_fraction = gc_fraction
            bio.SeqUtils = su
            sys.modules["Bio"] = bio
            sys.modules["Bio.SeqUtils"] = su


_loaded = None


def load():
    """Import the reference's hot-path modules; returns a namespace object."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_SRC}")
    _install_stubs()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    # `import crisprhawk` runs crisprhawk/__init__.py only; the workflow driver
    # (crisprhawk.crisprhawk) drags matplotlib & friends in and is not needed.
    import importlib

    ns = types.SimpleNamespace()
    for mod in (
        "encoder",
        "pam",
        "search_guides",
        "guide",
        "haplotype",
        "haplotypes",
        "variant",
        "region",
        "sequence",
        "coordinate",
        "utils",
        "region_constructor",
    ):
        setattr(ns, mod, importlib.import_module(f"crisprhawk.{mod}"))
    _loaded = ns
    return ns


class FakeVCF:
    """The only attribute of `VCF` the haplotype builder reads (haplotypes.py:759)."""

    def __init__(self, samples):
        self.samples = list(samples)


def build_case(
    ref_text: str,
    bed_start: int,
    bed_stop: int,
    vcf_lines,
    samples,
    phased: bool,
    contig: str = "chr1",
    padding: int = 100,
):
    """Build (region, haplotypes) through the reference's own classes, in memory.

    `ref_text` is the padded FASTA slice [bed_start - padding, bed_stop + padding]
    (1-based inclusive, sequence.py:342). `vcf_lines` are tab-separated VCF data
    lines. Haplotype ids are assigned deterministically as h0, h1, ... (the
    reference draws random ids, haplotypes.py:793-815).
    """
    ref = load()
    coord = ref.coordinate.Coordinate(contig, bed_start, bed_stop, padding)
    region = ref.region.Region(ref.sequence.Sequence(ref_text, False), coord)
    haps = ref.haplotypes.initialize_haplotypes(ref.region.RegionList([region]), True)
    variants = []
    for line in vcf_lines:
        v = ref.variant.VariantRecord(True)
        v.read_vcf_line(line.split("\t"), list(samples), phased)
        variants.extend(v.split())
    hl = haps[region]
    if vcf_lines:
        vcfs = {contig: FakeVCF(samples)}
        if phased:
            hl = ref.haplotypes.add_variants_phased(
                hl, region, vcfs, variants, True, True
            )
        else:
            hl = ref.haplotypes.add_variants_unphased(
                hl, region, vcfs, variants, False, True
            )
    for i, h in enumerate(hl):
        h.id = f"h{i}"
    return region, hl


def run_search(region, haps, pamseq, guidelen, right, variants_present, phased):
    """Run the reference's encode + search on reference objects."""
    ref = load()
    pam = ref.pam.PAM(pamseq, right, True)
    pam.encode(0)
    bits = [ref.encoder.encode(h.sequence.sequence, 0, True) for h in haps]
    guides = ref.search_guides.search(
        pam, region, haps, bits, guidelen, right, variants_present, phased, 0, True
    )
    return pam, bits, guides
