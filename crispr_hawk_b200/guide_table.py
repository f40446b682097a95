"""N3 (next row of the scope table): the guide table as the wire format between the scan and
its consumers, instead of a Python `List[Guide]` (guide.py:64-120).

`GuideTable` holds the structure-of-arrays columns the device produced -- rows in the
reference's final order (first-seen `(start, strand)` buckets, members in emission order) --
plus the N2 columns when they were computed, and hands them out the way the consumers take
them: whole columns for batched scorers (`scoring.py:49-84` slices every guide's sequence in a
Python loop), `Guide` objects one at a time, built on demand, for code that wants objects."""

from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

from . import marshal
from .guide import guide_class
from .pam import pam_patterns

GUIDESEQPAD = marshal.GUIDESEQPAD


class GuideTable:
    def __init__(self, table: Dict[str, np.ndarray], haplotypes, pam, guidelen: int, right: bool,
                 annotation: Optional[Dict[str, list]] = None, debug: bool = False):  # fmt: skip
        """`table`: columns of `search_table` / `hawk_result_fetch` / the streamed search
        (emission order); `annotation`: columns of `annotate_table` for the same rows."""
        self.order = np.argsort(table["bucket"], kind="stable")  # :306-369 -> final order
        self.table, self.haplotypes, self.annotation = table, haplotypes, annotation
        self.guidelen, self.pamlen, self.right, self.debug = int(guidelen), len(pam_patterns(pam)[0]), bool(right), debug
        self._cache: Dict[int, object] = {}

    def __len__(self) -> int:
        return len(self.order)

    # ---- columns, final order ----
    def column(self, name: str) -> np.ndarray:
        return self.table[name][self.order]

    def sequences(self) -> np.ndarray:
        """(n, window) uint8: the padded window text; after N2 the reverse-complemented one."""
        if self.annotation is not None:
            txt = np.frombuffer("".join(self.annotation["sequence"]).encode("ascii"), np.uint8)
            return txt.reshape(len(self.table["hap"]), -1)[self.order]
        return self.table["text"][self.order]

    def scorer_sequences(self, sgdesigner: bool = False) -> List[str]:
        """scoring.py:49-84 (_extract_guide_sequences[_sgdesigner]) for all guides at once:
        4 (or 0) bases upstream of the guide to 3 bases downstream of the PAM, upper-cased."""
        seq = self.sequences()
        lo = GUIDESEQPAD if sgdesigner else GUIDESEQPAD - 4
        part = seq[:, lo : seq.shape[1] - GUIDESEQPAD + 3] & np.uint8(0xDF)  # ASCII letters: upper-case
        return [bytes(r).decode("ascii") for r in part]

    # ---- lazy Guide views ----
    def guide(self, k: int):
        """The k-th guide of the reference's list as a `Guide` object (built on first use)."""
        if k in self._cache:
            return self._cache[k]
        i = int(self.order[k])
        t = self.table
        h = self.haplotypes[int(t["hap"][i])]
        s = int(t["strand"][i])
        rp = (not self.right) if s == 1 else self.right  # search_guides.py:538
        G, P = self.guidelen, self.pamlen
        pivot = int(t["pos"][i]) - (0 if rp else G)
        pm = h.posmap
        g = guide_class()(int(t["start"][i]), int(t["stop"][i]), t["text"][i].tobytes().decode("ascii"), G, P, s,
                          h.samples, h.variants, h.afs, {j: pm[pivot + j] for j in range(G + P)}, self.debug, rp, h.id)  # fmt: skip
        a = self.annotation
        if a is not None:  # what annotation.py:563-572 leaves in the object
            g.variants = a["variants"][i]
            g.afs_str = a["afs_str"][i].split(",")
            if s == 1:
                g.reverse_complement()
            g.gc = float(a["gc"][i])
        self._cache[k] = g
        return g

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self.guide(j) for j in range(*k.indices(len(self)))]
        if k < 0:
            k += len(self)
        if not 0 <= k < len(self):
            raise IndexError(k)
        return self.guide(k)

    def __iter__(self):
        return (self.guide(k) for k in range(len(self)))
