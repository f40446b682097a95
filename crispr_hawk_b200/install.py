"""Drop-in installation: rebind the three names `crisprhawk.crisprhawk` resolves at
call time (crisprhawk.py:18 `encode`, :29 `search`, :64 `encode_haplotypes`) so the
unchanged `crisprhawk search` CLI runs the GPU path."""

from __future__ import annotations

import importlib
from typing import Optional

from . import encoder, search_guides

_saved = {}


def install(module: Optional[object] = None):
    """Rebind in `crisprhawk.crisprhawk` (or the given module object). Returns it."""
    from . import _cabi

    _cabi.load_library()  # fail loudly now, not in the middle of a run
    drv = module or importlib.import_module("crisprhawk.crisprhawk")
    if drv in _saved:
        return drv
    _saved[drv] = {n: getattr(drv, n, None) for n in ("encode", "search", "encode_haplotypes")}
    drv.encode = encoder.encode
    drv.encode_haplotypes = encoder.encode_haplotypes
    drv.search = search_guides.search
    return drv


def uninstall(module: Optional[object] = None) -> None:
    drv = module or importlib.import_module("crisprhawk.crisprhawk")
    for name, fn in _saved.pop(drv, {}).items():
        if fn is not None:
            setattr(drv, name, fn)
