"""GPU replacement of `search_guides.search` (search_guides.py:510-548), same
signature, same `List[Guide]` (order included).

Division of labour: scan bounds are read from the haplotypes' own position maps
(host dict lookups, :49-84); PAM matching on both strands, the in-range and
REF-core filters, unphased IUPAC resolution, genomic start/stop, redundancy
removal, emission-order merge, bucket ids and window text all run on the GPU
behind `hawk_search` (include/hawkscan.h); this module turns the returned table
into `Guide` objects. There is no CPU implementation of the scan here."""

from __future__ import annotations

from time import time
from typing import List, Sequence, Tuple

import numpy as np

from . import _cabi, marshal
from .encoder import PackedRegion, encode_region
from .errors import error_class, exception_handler, print_verbosity
from .guide import guide_class
from .pam import pam_patterns

GUIDESEQPAD = marshal.GUIDESEQPAD


class _LiveTables:
    """Device-resident tables kept alive behind GuideLists (crisprhawk.py searches every region
    before it annotates any). Bounded: beyond `cap_bytes` of table memory the oldest links are
    released -- their lists then go to the reference's own annotation functions."""

    def __init__(self, cap_bytes=None):
        self.cap_bytes, self.total, self.links = cap_bytes, 0, []

    def _cap(self) -> int:
        if self.cap_bytes is None:  # half of the device's memory, asked once
            try:
                self.cap_bytes = int(_cabi.Context.default().info()["total_mem"]) // 2
            except Exception:
                self.cap_bytes = 16 << 30
        return self.cap_bytes

    def add(self, link: dict, nbytes: int) -> None:
        self.links = [(lk, nb) for lk, nb in self.links if lk.get("res") is not None]
        self.total = sum(nb for _, nb in self.links)
        self.links.append((link, nbytes))
        self.total += nbytes
        while self.total > self._cap() and len(self.links) > 1:
            old, nb = self.links.pop(0)
            if old.get("res") is not None:
                old["res"].close()
                old["res"] = None
            self.total -= nb


LIVE_TABLES = _LiveTables()


class GuideList(list):
    """`search()`'s return value: the reference's `List[Guide]` (buckets in first-seen order,
    search_guides.py:306-369) whose Guide objects are built on first access -- N3, the guide
    table as the wire format. The list carries the table it came from (`hawk`: columns,
    `_cabi.Result`, batch, haplotypes, table order) for the N2 mirrors in
    crispr_hawk_b200.annotation, which register their columns as *stages* here instead of
    touching every object: a Guide built later gets the stages applied at birth, one built
    earlier is updated when a stage arrives. Indexing, slicing and iteration build only what
    they touch; every other list operation (`+`, `sort`, `copy`, `in`, pickling ...) builds the
    remaining objects first and then behaves like the plain list it is."""

    def __init__(self, n: int, make, hawk=None):
        super().__init__([None] * n)
        self._make, self._pending, self.hawk = make, n, hawk
        self._stages = []  # [(apply(guide, final_index), ...)] in the order annotate_guides ran them

    # ---- element access: build on demand ----
    def _build(self, k: int):
        g = self._make(k)
        for apply in self._stages:
            apply(g, k)
        list.__setitem__(self, k, g)
        self._pending -= 1
        return g

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[j] for j in range(*k.indices(len(self)))]
        g = list.__getitem__(self, k)
        if g is None and self._pending:
            g = self._build(k if k >= 0 else k + len(self))
        return g

    def __iter__(self):
        for k in range(len(self)):
            yield self[k]

    def add_stage(self, apply) -> None:
        """Register one of annotate_guides' per-guide steps: applied now to the guides that
        exist, at construction to the others."""
        if self._pending:
            for k in range(len(self)):
                g = list.__getitem__(self, k)
                if g is not None:
                    apply(g, k)
            self._stages.append(apply)
        else:
            for k, g in enumerate(list.__iter__(self)):
                apply(g, k)

    def realise(self) -> "GuideList":
        """Build every remaining Guide (afterwards this is an ordinary list)."""
        if self._pending:
            for k in range(len(self)):
                if list.__getitem__(self, k) is None:
                    self._build(k)
            self._stages = []
        return self

    @property
    def built(self) -> int:
        return len(self) - self._pending


def _realised(name):
    base = getattr(list, name)

    def method(self, *a, **kw):
        self.realise()
        return base(self, *a, **kw)

    method.__name__ = name
    return method


for _name in ("__add__", "__radd__", "__iadd__", "__mul__", "__rmul__", "__imul__", "__contains__", "__reversed__", "__eq__",
              "__ne__", "__lt__", "__le__", "__gt__", "__ge__", "__reduce_ex__", "__repr__", "__setitem__", "__delitem__", "copy", "count",
              "index", "sort", "reverse", "pop", "remove", "insert", "append", "extend", "clear"):  # fmt: skip
    if hasattr(list, _name):
        setattr(GuideList, _name, _realised(_name))
GuideList.__hash__ = None


def _packed_for(haplotypes, haplotypes_bits, verbosity: int, debug: bool) -> PackedRegion:
    if isinstance(haplotypes_bits, PackedRegion) and haplotypes_bits.matches(haplotypes):
        return haplotypes_bits
    # bit lists produced elsewhere (e.g. the reference's own encoder) carry no
    # case plane: pack from the haplotype texts
    return encode_region(haplotypes, verbosity, debug)


def _prepare(pam, region, haplotypes, packed: PackedRegion, guidelen, right, unphased):
    fwd, rc = pam_patterns(pam)
    params = _cabi.make_params(fwd, rc, guidelen, right, unphased)
    bounds = [marshal.scan_bounds(h, region.start, region.stop, len(fwd)) for h in haplotypes]
    a = np.array([b[0] for b in bounds], dtype=np.int32)
    b = np.array([b[1] for b in bounds], dtype=np.int32)
    return params, a, b


def _split_hits(recs: np.ndarray, n_hap: int) -> List[List[int]]:
    hap = (recs >> np.uint64(32)).astype(np.int64)
    pos = (recs & np.uint64(0xFFFFFFFF)).astype(np.int64)
    cuts = np.searchsorted(hap, np.arange(n_hap + 1))
    return [pos[cuts[h] : cuts[h + 1]].tolist() for h in range(n_hap)]


def pam_search(pam, region, haplotypes, haplotypes_bits, verbosity: int, debug: bool) -> List[Tuple[List[int], List[int]]]:
    """search_guides.py:102-131 on the GPU: raw PAM occurrences per haplotype and strand."""
    packed = _packed_for(haplotypes, haplotypes_bits, verbosity, debug)
    params, a, b = _prepare(pam, region, haplotypes, packed, 1, False, False)
    res = _cabi.pam_search(packed.batch.ctx, packed.batch, params, a, b)
    fwd = _split_hits(res.hits(0), len(haplotypes))
    rev = _split_hits(res.hits(1), len(haplotypes))
    res.close()
    for h, (f, r) in zip(haplotypes, zip(fwd, rev)):
        print_verbosity(f"Searching PAM occurrences in haplotype {h.samples}", verbosity, 3)
        print_verbosity(
            f"Found {len(f) + len(r)} PAM occurrences ({len(f)} on 5'-3'; {len(r)} on 3'-5')", verbosity, 3
        )
    return list(zip(fwd, rev))


def search_table(pam, region, haplotypes, haplotypes_bits, guidelen: int, right: bool,
                 variants_present: bool, phased: bool, verbosity: int = 0, debug: bool = False,
                 want_text: bool = True):  # fmt: skip
    """Run the device pipeline and return (table dict, Result) without building Guides.
    `want_text=False` leaves the window-text column on the device (48 of the 69 bytes a row
    takes across PCIe): for a phased / variant-free search it is a slice of the haplotype text
    the caller already holds (search_guides.py:134-160)."""
    packed = _packed_for(haplotypes, haplotypes_bits, verbosity, debug)
    batch = packed.batch
    unphased = bool(variants_present and not phased)
    params, a, b = _prepare(pam, region, haplotypes, packed, guidelen, right, unphased)
    if not batch.has_posmap:
        batch.set_posmap(marshal.segment_table(haplotypes))
    if unphased and not batch.has_alleles:
        batch.set_alleles(marshal.allele_table(haplotypes))
    is_ref = np.array([h.samples == "REF" for h in haplotypes], dtype=np.uint8)
    try:
        res = _cabi.search(batch.ctx, batch, params, a, b, is_ref)
    except _cabi.HawkLibraryError as e:
        if e.code == _cabi.HAWK_EALLELES:
            # the reference dies with a bare KeyError here (search_guides.py:207-213)
            raise KeyError("ambiguity code without variant_alleles entry inside a guide window") from e
        if e.code == _cabi.HAWK_EDUPREF:
            exception_handler(
                error_class("CrisprHawkCfdScoreError"),
                "Duplicate REF guide at position ? CFDon calculation failed",
                65, debug, e,
            )  # fmt: skip
        raise
    res.batch_ref = batch  # the batch the table was computed from (kept alive with the result)
    return res.table(want_text=want_text), res


def search(pam, region, haplotypes, haplotypes_bits, guidelen: int, right: bool,
           variants_present: bool, phased: bool, verbosity: int, debug: bool) -> list:  # fmt: skip
    print_verbosity(f"Searching guide candidates in {region.coordinates}", verbosity, 3)
    if verbosity >= 3:
        pam_search(pam, region, haplotypes, haplotypes_bits, verbosity, debug)
    start_t = time()
    unphased = bool(variants_present and not phased)
    # unphased rows hold resolved strings; otherwise the text is a window of the haplotype's own
    table, res = search_table(pam, region, haplotypes, haplotypes_bits, guidelen, right,
                              variants_present, phased, verbosity, debug, want_text=unphased)  # fmt: skip
    Guide = guide_class()
    pamlen = len(pam_patterns(pam)[0])
    span = guidelen + pamlen
    # remove_redundant_guides returns buckets in first-seen order, members in emission
    # order (:306-369): a stable sort of the emission-ordered table by bucket id
    order = np.argsort(table["bucket"], kind="stable")
    hap_i, strand, pos = table["hap"], table["strand"], table["pos"]
    starts, stops = table["start"], table["stop"]
    texts = table.get("text")
    hap_texts = {}

    def make(k: int):
        """The k-th guide of the reference's list (guide.py:64-120 arguments, :488-503)."""
        i = int(order[k])
        hi = int(hap_i[i])
        h = haplotypes[hi]
        s = int(strand[i])
        p = int(pos[i])
        rp = (not right) if s == 1 else bool(right)  # :538
        pivot = p if rp else p - guidelen
        pm = h.posmap
        gpm = {j: pm[pivot + j] for j in range(span)}  # retrieve_guide_posmap :283-303
        if texts is not None:
            seq = texts[i].tobytes().decode("ascii")
        else:  # extract_guide_sequence :134-160
            t = hap_texts.get(hi)
            if t is None:
                t = hap_texts[hi] = marshal.hap_text(h)
            w0 = pivot - GUIDESEQPAD
            seq = t[w0 : w0 + span + 2 * GUIDESEQPAD]
        return Guide(int(starts[i]), int(stops[i]), seq, guidelen, pamlen, s, h.samples, h.variants, h.afs, gpm, debug, rp, h.id)

    print_verbosity(f"Guides retrieved in {time() - start_t:.2f}s", verbosity, 3)
    if not unphased:
        # N2 seam: the device-resident table travels with the list, so the mirrors of
        # annotation.py's per-guide loops (crispr_hawk_b200.annotation) can run on it
        link = dict(table=table, res=res, batch=res.batch_ref, haplotypes=haplotypes, right=bool(right), order=order, pam=pam)
        LIVE_TABLES.add(link, res.device_bytes() + getattr(res.batch_ref, "device_bytes", lambda: 0)())
        return GuideList(len(order), make, link)
    res.close()
    return GuideList(len(order), make, None)
