"""The drop-in seam: install() rebinds the three names crisprhawk/crisprhawk.py resolves at
call time (crisprhawk.py:18 encode, :29 search, :64 encode_haplotypes) and uninstall() puts
the originals back. Needs libhawkscan.so (it is loaded up front so a missing library fails
here, not in the middle of a run) but no GPU."""

import types

import pytest

import crispr_hawk_b200 as hawk
from crispr_hawk_b200 import _cabi, encoder, search_guides


def _fake_driver():
    m = types.ModuleType("crisprhawk.crisprhawk")
    m.encode = lambda *a: "ref-encode"
    m.search = lambda *a: "ref-search"
    m.encode_haplotypes = lambda *a: "ref-encode-haplotypes"
    return m


def test_install_rebinds_and_uninstall_restores():
    drv = _fake_driver()
    orig = (drv.encode, drv.search, drv.encode_haplotypes)
    assert hawk.install(drv) is drv
    assert drv.encode is encoder.encode
    assert drv.search is search_guides.search
    assert drv.encode_haplotypes is encoder.encode_haplotypes
    hawk.install(drv)  # idempotent
    hawk.uninstall(drv)
    assert (drv.encode, drv.search, drv.encode_haplotypes) == orig


def test_install_fails_loudly_without_the_library(monkeypatch, tmp_path):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "libhawkscan.so"))
    with pytest.raises(_cabi.HawkLibraryError):
        hawk.install(_fake_driver())


def test_signatures_match_the_reference():
    import inspect

    assert list(inspect.signature(hawk.encode).parameters) == ["sequence", "verbosity", "debug"]
    assert list(inspect.signature(hawk.search).parameters) == [
        "pam", "region", "haplotypes", "haplotypes_bits", "guidelen", "right", "variants_present", "phased",
        "verbosity", "debug",
    ]  # fmt: skip  (search_guides.py:510-521)
    assert list(inspect.signature(hawk.encode_haplotypes).parameters) == ["haplotypes", "args"]  # crisprhawk.py:64


@pytest.mark.ref
def test_signatures_equal_live_reference():
    import inspect

    from oracle import refshim

    ref = refshim.load()
    assert list(inspect.signature(ref.search_guides.search).parameters) == list(inspect.signature(hawk.search).parameters)
    assert list(inspect.signature(ref.encoder.encode).parameters) == list(inspect.signature(hawk.encode).parameters)


def test_library_exports_every_declared_symbol():
    """Every function include/hawkscan.h declares is exported by libhawkscan.so and bound by
    the ctypes layer (no compute calls: works without a GPU)."""
    import ctypes
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "hawkscan.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(hawk_[a-z0-9_]+)\s*\(", text))
    declared -= {"hawk_ctx", "hawk_batch", "hawk_result", "hawk_params"}
    assert len(declared) >= 30
    lib = _cabi.load_library()
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in hawkscan.h but not exported"
        assert name in _cabi.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.hawk_abi_version() == _cabi.ABI_VERSION
    assert lib.hawk_strerror(_cabi.HAWK_EIUPAC).decode() == "non-IUPAC character"
    # host-only helpers work without a device
    import numpy as np

    lens = np.array([5, 300, 0], np.int32)
    off = np.zeros(4, np.int64)
    total = ctypes.c_int64()
    assert lib.hawk_layout(_cabi.ptr(lens, ctypes.c_int32), 3, _cabi.ptr(off, ctypes.c_int64), ctypes.byref(total)) == 0
    from crispr_hawk_b200 import marshal

    want, wtotal = marshal.layout(lens)
    assert off.tolist() == want.tolist() and total.value == wtotal


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU the product path must fail loudly, never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(_cabi.HawkLibraryError) as ei:
        _cabi.Context(0)
    assert "no CPU path" in str(ei.value) or "CUDA" in str(ei.value)
    with pytest.raises(_cabi.HawkLibraryError):
        hawk.encode("ACGT", 0, True)


def test_install_rebinds_the_annotation_seam_and_falls_through_for_foreign_lists():
    """N2: the four per-guide loops of annotation.annotate_guides (annotation.py:563-572) are
    rebound; a list that did not come from crispr_hawk_b200.search goes to the originals."""
    from crispr_hawk_b200 import annotation as ann

    drv = _fake_driver()
    amod = types.ModuleType("crisprhawk.annotation")
    calls = []
    for name in ann.SEAM:
        setattr(amod, name, (lambda n: (lambda guides, *a: calls.append(n) or guides))(name))
    orig = {n: getattr(amod, n) for n in ann.SEAM}
    hawk.install(drv, annotation_module=amod)
    for n in ann.SEAM:
        assert getattr(amod, n) is getattr(ann, n)
    plain = ["g1", "g2"]
    assert amod._annotate_variants(plain, 0, True) is plain
    assert amod.annotate_variants_afs(plain, 0) is plain
    assert amod.reverse_guides(plain, 0) is plain
    assert amod.gc_content(plain, 0, True) is plain
    assert calls == list(ann.SEAM)
    hawk.uninstall(drv, annotation_module=amod)
    assert {n: getattr(amod, n) for n in ann.SEAM} == orig
    with pytest.raises(RuntimeError):  # no reference installed and not our list: no CPU path here
        ann.gc_content(plain, 0, True)


def test_annotation_seam_signatures():
    import inspect

    from crispr_hawk_b200 import annotation as ann

    assert list(inspect.signature(ann._annotate_variants).parameters) == ["guides", "verbosity", "debug"]
    assert list(inspect.signature(ann.annotate_variants_afs).parameters) == ["guides", "verbosity"]
    assert list(inspect.signature(ann.reverse_guides).parameters) == ["guides", "verbosity"]
    assert list(inspect.signature(ann.gc_content).parameters) == ["guides", "verbosity", "debug"]


def test_live_device_tables_are_bounded():
    """search() keeps a region's device table alive for the annotation seam; crisprhawk.py searches
    every region before annotating any, so the total is capped and the oldest links are released."""
    from crispr_hawk_b200.search_guides import _LiveTables

    class Res:
        def __init__(self):
            self.closed = False

        def close(self):
            self.closed = True

    live = _LiveTables(cap_bytes=100)
    links = [dict(res=Res()) for _ in range(5)]
    keep = [lk["res"] for lk in links]
    for lk in links[:3]:
        live.add(lk, 40)
    assert keep[0].closed and links[0]["res"] is None  # 120 > 100: the oldest went
    assert not keep[1].closed and not keep[2].closed and live.total == 80
    links[1]["res"].close()
    links[1]["res"] = None  # annotated (annotation._columns releases the table itself)
    live.add(links[3], 40)
    assert live.total == 80 and not keep[2].closed and not keep[3].closed
    live.add(links[4], 500)  # one table larger than the cap stays (the newest is never dropped)
    assert keep[2].closed and keep[3].closed and not keep[4].closed and len(live.links) == 1


def test_two_variants_at_one_position_go_to_the_reference(monkeypatch):
    """N2's guard: a haplotype with two variants at one normalised position (an SNV and an insertion
    at the same anchor) has no well-defined annotation in the reference (its variant map keeps
    whichever a Python set yields last, annotation.py:96, 129-160). The table builder flags it, the
    device annotation refuses it, and the seam reports 'not mine' so the reference's functions run."""
    from crispr_hawk_b200 import annotation, marshal

    class H:
        def __init__(self, variants):
            self.variants = variants

    plain = marshal.variant_table([H("NA"), H("chr1-100-A/G,chr1-140-AT/A"), H("chr1-100-A/G,chr1-101-C/CT")])
    assert not plain.ambiguous
    clash = marshal.variant_table([H("NA"), H("chr1-100-A/G,chr1-100-A/AT")])
    assert clash.ambiguous and clash.var_pos.tolist() == [100, 100]

    closed = []

    class Res:
        handle = object()

        def close(self):
            closed.append(True)

    class Batch:
        has_variants = False

        def set_variants(self, vt):
            raise AssertionError("an ambiguous table must not reach the device")

    class L(list):
        pass

    guides = L()
    guides.hawk = {"res": Res(), "batch": Batch(), "table": {"hap": []}, "haplotypes": [H("chr1-5-A/G,chr1-5-A/AT")], "right": False,
                   "order": []}  # fmt: skip
    assert annotation._columns(guides, True) is None and closed == [True]
    assert guides.hawk["cols"] is None and guides.hawk["res"] is None
    assert annotation._columns(guides, True) is None  # sticky
    called = []
    monkeypatch.setitem(annotation._reference, "gc_content", lambda g, v, d: called.append("ref") or g)
    assert annotation.gc_content(guides, 0, True) is guides and called == ["ref"]


def test_scorer_strings_fall_back_when_the_list_changed(monkeypatch):
    """scoring._extract_guide_sequences hands a list to the reference's own function when the device
    columns do not describe it any more (a caller filtered the list) or were never computed."""
    import numpy as np

    from crispr_hawk_b200 import scoring

    class L(list):
        pass

    calls = []
    monkeypatch.setitem(scoring._reference, "_extract_guide_sequences", lambda g: calls.append(len(g)) or ["ref"] * len(g))
    k4 = np.frombuffer(b"ACGTAC" * 3, np.uint8).reshape(3, 6)
    guides = L(["g0", "g1", "g2"])
    guides.hawk = {"kmers": {4: k4, 0: k4[:, 4:]}, "order": np.array([2, 0, 1])}
    assert scoring._extract_guide_sequences(guides) == ["ACGTAC"] * 3 and calls == []
    assert scoring._extract_guide_sequences_sgdesigner.__name__ == "_extract_guide_sequences_sgdesigner"
    short = L(["g0", "g1"])
    short.hawk = guides.hawk
    assert scoring._extract_guide_sequences(short) == ["ref", "ref"] and calls == [2]
    none = L(["g0"])
    none.hawk = {"kmers": None, "order": np.array([0])}
    assert scoring._extract_guide_sequences(none) == ["ref"] and calls == [2, 1]
    assert scoring._extract_guide_sequences(["x", "y"]) == ["ref", "ref"]  # a plain list
