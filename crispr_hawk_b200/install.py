"""Drop-in installation: rebind the three names `crisprhawk.crisprhawk` resolves at
call time (crisprhawk.py:18 `encode`, :29 `search`, :64 `encode_haplotypes`) so the
unchanged `crisprhawk search` CLI runs the GPU path."""

from __future__ import annotations

import importlib
from typing import Optional

from . import encoder, search_guides

_saved = {}


def install(module: Optional[object] = None, annotation_module: Optional[object] = None, annotation: bool = True,
            haplotypes_module: Optional[object] = None, haplotypes: bool = True,
            reports_module: Optional[object] = None, reports: bool = True,
            scoring_module: Optional[object] = None, scoring: bool = True):
    """Rebind in `crisprhawk.crisprhawk` (or the given module object). Returns it.

    With `annotation` (N2) the four per-guide loops `annotation.annotate_guides` runs right
    after `search()` -- `_annotate_variants`, `annotate_variants_afs`, `reverse_guides`,
    `gc_content` (annotation.py:563-572) -- are rebound in `crisprhawk.annotation` (or the
    given module object) too; they act on lists returned by this package's `search` and hand
    every other list to the reference's own functions. With `reports` (N2, second half)
    `reports._process_data` and `reports._collapse_report_entries` (reports.py:476-531, 958-1008)
    are rebound the same way: the row collapse uses the groups the device computed over the
    resident table instead of a pandas groupby. With `scoring` (N4) `scoring.cfdon_score`
    (scoring.py:352-387) is rebound: the scores come from `hawk_result_cfdon`, computed with the
    reference's own factor tables while the table is on the device; so are
    `scoring._extract_guide_sequences` / `_extract_guide_sequences_sgdesigner` (scoring.py:50-84),
    the learned scorers' input strings, cut for the whole batch by `hawk_result_featurize`."""
    from . import _cabi
    from . import report_rows as rep
    from . import scoring as sco
    from . import annotation as ann
    from . import haplotypes as hapmod

    _cabi.load_library()  # fail loudly now, not in the middle of a run
    drv = module or importlib.import_module("crisprhawk.crisprhawk")
    if drv not in _saved:
        _saved[drv] = {n: getattr(drv, n, None) for n in ("encode", "search", "encode_haplotypes")}
        drv.encode = encoder.encode
        drv.encode_haplotypes = encoder.encode_haplotypes
        drv.search = search_guides.search
    if annotation:
        amod = annotation_module
        if amod is None and module is None:
            amod = importlib.import_module("crisprhawk.annotation")
        if amod is not None and amod not in _saved:
            _saved[amod] = {n: getattr(amod, n, None) for n in ann.SEAM}
            for n in ann.SEAM:
                if _saved[amod][n] is not None:
                    ann._reference[n] = _saved[amod][n]
                setattr(amod, n, getattr(ann, n))
    if haplotypes:
        # N1: the phased branch of haplotype assembly (haplotypes.py:716-751) builds its haplotypes
        # on the device from edit lists; shapes it does not take go to the reference's builder
        hmod = haplotypes_module
        if hmod is None and module is None:
            hmod = importlib.import_module("crisprhawk.haplotypes")
        if hmod is not None and hmod not in _saved:
            _saved[hmod] = {"add_variants_phased": getattr(hmod, "add_variants_phased", None)}
            hapmod._reference_add_variants_phased = _saved[hmod]["add_variants_phased"]
            hmod.add_variants_phased = hapmod.add_variants_phased
    if reports:
        rmod = reports_module
        if rmod is None and module is None:
            rmod = importlib.import_module("crisprhawk.reports")
        if rmod is not None and rmod not in _saved:
            _saved[rmod] = {n: getattr(rmod, n, None) for n in rep.SEAM}
            for n in rep.SEAM:
                if _saved[rmod][n] is not None:
                    rep._reference[n] = _saved[rmod][n]
                    setattr(rmod, n, getattr(rep, n))
    if scoring:
        smod = scoring_module
        if smod is None and module is None:
            try:
                smod = importlib.import_module("crisprhawk.scoring")
            except Exception:
                smod = None  # the scorers' own dependencies are missing: nothing to rebind
        if smod is not None and smod not in _saved:
            _saved[smod] = {n: getattr(smod, n, None) for n in sco.SEAM}
            for n in sco.SEAM:
                if _saved[smod][n] is not None:
                    sco._reference[n] = _saved[smod][n]
                    setattr(smod, n, getattr(sco, n))
            try:
                pmod = importlib.import_module("crisprhawk.pam")
                sco._reference["cas9_systems"] = (pmod.SPCAS9, pmod.XCAS9)
                sco._reference["cpf1_system"] = pmod.CPF1
            except Exception:
                pass
    return drv


def uninstall(module: Optional[object] = None, annotation_module: Optional[object] = None,
              haplotypes_module: Optional[object] = None, reports_module: Optional[object] = None,
              scoring_module: Optional[object] = None) -> None:
    from . import annotation as ann
    from . import haplotypes as hapmod
    from . import report_rows as rep
    from . import scoring as sco

    smod = scoring_module
    if smod is None and module is None:
        smod = __import__("sys").modules.get("crisprhawk.scoring")
    if smod is not None:
        for name, fn in _saved.pop(smod, {}).items():
            if fn is not None:
                setattr(smod, name, fn)
        sco._reference.clear()

    rmod = reports_module
    if rmod is None and module is None:
        try:
            rmod = importlib.import_module("crisprhawk.reports")
        except Exception:
            rmod = None
    if rmod is not None:
        for name, fn in _saved.pop(rmod, {}).items():
            if fn is not None:
                setattr(rmod, name, fn)
            rep._reference.pop(name, None)

    hmod = haplotypes_module
    if hmod is None and module is None:
        try:
            hmod = importlib.import_module("crisprhawk.haplotypes")
        except Exception:
            hmod = None
    if hmod is not None:
        for name, fn in _saved.pop(hmod, {}).items():
            if fn is not None:
                setattr(hmod, name, fn)
        hapmod._reference_add_variants_phased = None

    drv = module or importlib.import_module("crisprhawk.crisprhawk")
    for name, fn in _saved.pop(drv, {}).items():
        if fn is not None:
            setattr(drv, name, fn)
    amod = annotation_module
    if amod is None and module is None:
        try:
            amod = importlib.import_module("crisprhawk.annotation")
        except Exception:
            amod = None
    if amod is not None:
        for name, fn in _saved.pop(amod, {}).items():
            if fn is not None:
                setattr(amod, name, fn)
            ann._reference.pop(name, None)
