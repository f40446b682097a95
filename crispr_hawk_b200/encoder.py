"""GPU replacements of `encoder.encode` (encoder.py:48-57) and of
`crisprhawk.encode_haplotypes` (crisprhawk.py:64-81).

`encode()` keeps the reference's signature and returns the nibble list it would
return (computed by the K1 pack kernel). `encode_haplotypes()` packs all
haplotypes of a region in one launch and returns device-resident handles that
`search()` consumes directly; `haplotypes_bits` is opaque to every other caller
of the reference."""

from __future__ import annotations

import os
from time import time
from typing import Dict, List, Sequence

import numpy as np

from . import _cabi, marshal
from .errors import error_class, exception_handler, print_verbosity


def _raise_iupac(texts: Sequence[str], hap: int, pos: int, debug: bool):
    nt = texts[hap][pos].upper()
    exception_handler(
        error_class("CrisprHawkIupacTableError"),
        f"The nucleotide {nt} at position {pos} is not a IUPAC character",
        os.EX_DATAERR,
        debug,
    )


def pack_texts(texts: Sequence[str], debug: bool, ctx=None) -> "_cabi.Batch":
    """Stage haplotype texts in the slot layout and run K1 on them."""
    ctx = ctx or _cabi.Context.default()
    try:
        buf, off, lens = marshal.stage_ascii(texts)
    except marshal.InvalidSequence as e:
        _raise_iupac(texts, e.hap_index, e.position, debug)
    try:
        return _cabi.Batch(ctx, buf, off, lens)
    except _cabi.HawkLibraryError as e:
        if e.code != _cabi.HAWK_EIUPAC:
            raise
        slot = e.bad_slot
        hap = int(np.searchsorted(off, slot, side="right") - 1)
        _raise_iupac(texts, hap, int(slot - off[hap]), debug)


class PackedRegion:
    """`haplotypes_bits[region]`: behaves like the reference's List[List[int]]
    (nibble lists are downloaded on demand) and carries the device batch."""

    def __init__(self, batch: "_cabi.Batch", texts: Sequence[str]):
        self.batch = batch
        self._texts = list(texts)
        self._cache: Dict[int, List[int]] = {}

    def __len__(self) -> int:
        return self.batch.n_hap

    def __getitem__(self, i: int) -> List[int]:
        if i < 0:
            i += len(self)
        if i not in self._cache:
            self._cache[i] = self.batch.export_nibbles(i).tolist()
        return self._cache[i]

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    owners = None  # ids of the haplotype objects the batch was built for (device-materialised)

    def matches(self, haplotypes) -> bool:
        if self.owners is not None:
            return [id(h) for h in haplotypes] == self.owners
        return len(haplotypes) == len(self._texts) and all(
            marshal.hap_text(h) is t or marshal.hap_text(h) == t for h, t in zip(haplotypes, self._texts)
        )


class EncodedBits(list):
    """Return value of `encode()`: the reference's nibble list, plus the packed batch."""

    batch = None


def encode(sequence: str, verbosity: int, debug: bool) -> List[int]:
    print_verbosity(f"Encoding sequence {sequence} in bits", verbosity, 3)
    start = time()
    batch = pack_texts([sequence], debug)
    bits = EncodedBits(batch.export_nibbles(0).tolist() if len(sequence) else [])
    bits.batch = batch
    assert len(bits) == len(sequence)
    print_verbosity(f"Encoding completed in {time() - start:.2f}s", verbosity, 3)
    return bits


def encode_region(haplotypes, verbosity: int, debug: bool) -> PackedRegion:
    pack = getattr(haplotypes[0], "_region_pack", None) if len(haplotypes) else None
    if pack is not None and pack.matches(haplotypes):
        return pack  # N1: the haplotypes were materialised on the device from edit lists, packed already
    texts = [marshal.hap_text(h) for h in haplotypes]
    return PackedRegion(pack_texts(texts, debug), texts)


def encode_haplotypes(haplotypes, args):
    """Drop-in for crisprhawk.crisprhawk.encode_haplotypes (crisprhawk.py:64-81)."""
    print_verbosity("Encoding haplotypes in bits", args.verbosity, 1)
    start = time()
    out = {region: encode_region(haps, args.verbosity, args.debug) for region, haps in haplotypes.items()}
    print_verbosity(f"Haplotype encoding completed in {time() - start:.2f}s", args.verbosity, 2)
    return out
