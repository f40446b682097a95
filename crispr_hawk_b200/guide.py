"""`Guide` records returned by `search()`.

When the reference package is importable, `guide_class()` returns its own
`crisprhawk.guide.Guide` so downstream annotation / scoring / reporting receive
exactly the objects they expect. Otherwise a field-compatible mirror of
guide.py:24-120 is used (stand-alone use, GPU-box tests)."""

from __future__ import annotations

from typing import Dict

GUIDESEQPAD = 10  # guide.py:21

_COMPLEMENT = str.maketrans("ACGTURYMKHDBVNSWacgturymkhdbvnsw", "TGCAAYRKMDHVBNSWtgcaayrkmdhvbnsw")


class Guide:
    """Mirror of the reference's Guide (guide.py:64-120): padded window text plus
    coordinates, strand, haplotype annotations and the per-base position map."""

    def __init__(self, position_start: int, position_stop: int, sequence: str, guidelen: int,
                 pamlen: int, direction: int, samples: str, variants: str, afs: Dict[str, float],
                 posmap: Dict[int, int], debug: bool, right: bool, hapid: str) -> None:  # fmt: skip
        self._debug = debug
        self._guidelen, self._pamlen = guidelen, pamlen
        self._start, self._stop = position_start, position_stop
        self._sequence = sequence
        self._right = right
        self._split()
        self._direction = direction
        self._samples, self._variants, self._afs = samples, variants, afs
        self._posmap = posmap
        self._hapid = hapid
        # guide.py:199-212
        self._guide_id = f"{self._start}_{self._stop}_{self._direction}_{self._hapid}_{self._guideseq}"
        for score in ("azimuth", "rs3", "cfdon", "elevationon", "deepcpf1", "ooframe", "plmcrispr",
                      "crispron", "sgdesigner"):  # fmt: skip
            setattr(self, f"_{score}_score", "NA")
        self._cfd = self._gc = self._offtargets_num = "NA"
        self._funcann, self._geneann = [], []

    def _split(self) -> None:
        core = self._sequence[GUIDESEQPAD:-GUIDESEQPAD]  # guide.py:184-197
        if self._right:
            self._pamseq, self._guideseq = core[: self._pamlen], core[self._pamlen :]
        else:
            self._pamseq, self._guideseq = core[-self._pamlen :], core[: -self._pamlen]

    def reverse_complement(self) -> None:  # guide.py:245-255
        self._sequence = self._sequence[::-1].translate(_COMPLEMENT)
        self._right = not self._right
        self._split()

    def __repr__(self) -> str:
        return (f"<{self.__class__.__name__} object; start={self._start} stop={self._stop} "
                f"sequence={self._sequence} direction={self._direction}>")  # fmt: skip

    def __len__(self) -> int:
        return len(self._sequence)

    start = property(lambda s: s._start)
    stop = property(lambda s: s._stop)
    strand = property(lambda s: s._direction)
    sequence = property(lambda s: s._sequence)
    samples = property(lambda s: s._samples)
    afs = property(lambda s: s._afs)
    hapid = property(lambda s: s._hapid)
    pam = property(lambda s: s._pamseq)
    pamlen = property(lambda s: s._pamlen)
    guide = property(lambda s: s._guideseq)
    guidelen = property(lambda s: s._guidelen)
    right = property(lambda s: s._right)
    guide_id = property(lambda s: s._guide_id)

    @property
    def guidepam(self) -> str:
        return self._pamseq + self._guideseq if self._right else self._guideseq + self._pamseq

    @property
    def variants(self) -> str:
        return self._variants

    @variants.setter
    def variants(self, value: str) -> None:
        self._variants = value

    @property
    def afs_str(self) -> str:  # guide.py:307-328
        return self._afs_

    @afs_str.setter
    def afs_str(self, value) -> None:
        self._afs_ = "NA" if not value or (len(set(value)) == 1 and value[0] == "NA") else ",".join(value)

    @property
    def gc(self) -> str:  # guide.py:594-618
        return self._gc

    @gc.setter
    def gc(self, value: float) -> None:
        if not isinstance(value, float):
            from .errors import CrisprHawkGuideError

            raise CrisprHawkGuideError(f"\n\nOut-of-frame score must be a float, got {type(value).__name__} instead")
        self._gc = str(value)

    @property
    def posmap(self) -> Dict[int, int]:
        return self._posmap

    @posmap.setter
    def posmap(self, value: Dict[int, int]) -> None:
        self._posmap = value


def guide_class():
    try:
        from crisprhawk.guide import Guide as RefGuide  # type: ignore

        return RefGuide
    except Exception:
        return Guide
