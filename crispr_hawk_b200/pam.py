"""PAM description for the scan (mirror of pam.py:46-173).

The reference's own `PAM` object is accepted everywhere (`search()` reads only
`.pam`, `len()`); this mirror exists for stand-alone use and for the tests that
run where the reference is not installed."""

from __future__ import annotations

import os
from typing import List

from . import marshal
from .errors import exception_handler

# Cas systems and their PAM tables (pam.py:18-36, 114-125)
CASX, CPF1, SACAS9, SPCAS9, XCAS9 = 0, 1, 2, 3, 4
_CAS_TABLE = [
    (CASX, {"TTCN"}, None),
    (CPF1, {"TTN", "TTTN", "TYCV", "TATV", "TTTV", "TTTR", "ATTN", "TTTA", "TCTA", "TCCA", "CCCA",
            "YTTV", "TTYN"}, True),
    (SACAS9, {"NNGRRT", "NNNRRT"}, None),
    (SPCAS9, {"NGG", "NGA", "NRG", "NGC"}, False),
    (XCAS9, {"NGK", "NGN", "NNG"}, False),
]  # fmt: skip

_RC = str.maketrans("ACGTRYMKHDBVNSW", "TGCAYRKMDHVBNSW")  # utils.py:46-79


def reverse_complement(seq: str) -> str:
    return seq.upper()[::-1].translate(_RC)


class PAM:
    def __init__(self, pamseq: str, right: bool, debug: bool):
        self._debug = debug
        if any(c.upper() not in marshal._NIBBLE for c in pamseq):
            exception_handler(ValueError, f"Invalid PAM sequence {pamseq}", os.EX_DATAERR, debug)
        self._sequence = pamseq.upper()
        self._sequence_rc = reverse_complement(self._sequence)
        self._cas_system = -1
        for system, pams, need_right in _CAS_TABLE:
            if self._sequence in pams and (need_right is None or need_right == bool(right)):
                self._cas_system = system
                break

    def encode(self, verbosity: int = 0) -> None:
        self._sequence_bits = marshal.pam_nibbles(self._sequence)
        self._sequence_rc_bits = marshal.pam_nibbles(self._sequence_rc)
        self._packed_bits = _pack(self._sequence_bits)
        self._packed_bitsrc = _pack(self._sequence_rc_bits)

    def __len__(self) -> int:
        return len(self._sequence)

    def __str__(self) -> str:
        return self._sequence

    def __repr__(self) -> str:
        return f"<{self.__class__.__name__} object; sequence={self._sequence}>"

    def __eq__(self, other) -> bool:
        return self._sequence == other.pam if hasattr(other, "pam") else NotImplemented

    pam = property(lambda s: s._sequence)
    pamrc = property(lambda s: s._sequence_rc)
    bits = property(lambda s: s._packed_bits)
    bitsrc = property(lambda s: s._packed_bitsrc)
    bits_list = property(lambda s: s._sequence_bits)
    cas_system = property(lambda s: s._cas_system)


def _pack(bits: List[int]) -> int:
    v = 0
    for b in bits:  # first base in the most significant nibble (pam.py:169-173)
        v = (v << 4) | b
    return v


def pam_patterns(pam) -> tuple:
    """(forward nibbles, reverse-complement nibbles) of any PAM-like object or string."""
    seq = pam if isinstance(pam, str) else pam.pam
    rc = getattr(pam, "pamrc", None) or reverse_complement(seq)
    return marshal.pam_nibbles(seq), marshal.pam_nibbles(rc)
