"""N4 (next row): the CFDon specificity score -- `scoring.cfdon_score` (scoring.py:352-387) with
`group_guides_position` (:303-349), `scores.crisprhawk_scores.cfdon` (:65-87) and
`scores.cfdscore.cfdscore.compute_cfd` (:53-95) -- for the guide table of a phased / variant-free
search.

The reference groups the guides by (start, strand) in a dict, finds each group's REF guide and
multiplies, per guide, the mismatch factors of the positions where it differs from the REF guide
and the factor of its PAM's last two letters. Here the grouping is the table's bucket column, the
REF guide of a key is the key's first row (REF rows are emitted first), and one kernel computes
every row's product in the reference's order in double precision (`hawk_result_cfdon`): the
floats are bit-identical, the host only rounds and formats them like the `Guide.cfdon_score`
setter (guide.py:468-488). The factor tables are the reference's own model files, loaded through
its own loader at run time."""

from __future__ import annotations

import importlib
from typing import Dict, Tuple

import numpy as np

from .errors import error_class, exception_handler

_reference: Dict[str, object] = {}

_DNA = "ACGT"
_RNA = "ACGU"
_RC = {"A": "T", "C": "G", "G": "C", "U": "A"}  # utils.py:46-79 on upper-case RNA letters


def cfd_tables(mmscores: Dict[str, float], pamscores: Dict[str, float]) -> Tuple[np.ndarray, np.ndarray]:
    """The reference's two dicts as dense arrays: mm[i, w, g] = mmscores["r<w>:d<revcomp(g)>,<i+1>"]
    (cfdscore.py:89-93), pam2[a, b] = pamscores["<a><b>"] (:94); NaN where a key is absent."""
    mm = np.full((20, 4, 4), np.nan, np.float64)
    for i in range(20):
        for w in range(4):
            for g in range(4):
                if w != g:
                    mm[i, w, g] = mmscores.get(f"r{_RNA[w]}:d{_RC[_RNA[g]]},{i + 1}", np.nan)
    pam2 = np.array([[pamscores.get(_DNA[a] + _DNA[b], np.nan) for b in range(4)] for a in range(4)], np.float64)
    return mm, pam2


def load_tables(debug: bool):
    """The reference's model files through the loader its own `cfdon` calls."""
    fn = _reference.get("load_tables")
    if fn is None:
        mod = importlib.import_module("crisprhawk.scores.crisprhawk_scores")
        fn = mod.load_mismatch_pam_scores
    return cfd_tables(*fn(debug))


def cas9_systems():
    got = _reference.get("cas9_systems")
    if got is None:
        from . import pam as mirror

        got = (mirror.SPCAS9, mirror.XCAS9)
    return got


def cfdon_column(link, debug: bool = True):
    """Scores of every table row (emission order) while the table is still on the device, or None
    when this package has nothing to say: not a Cas9 PAM (scoring.py:845-857 never calls CFDon
    then), REF not the first haplotype, score tables not loadable here. A KeyError of the
    reference comes back as the exception to raise."""
    from . import _cabi

    pam, haps, res = link.get("pam"), link["haplotypes"], link.get("res")
    if pam is None or getattr(pam, "cas_system", None) not in cas9_systems() or res is None:
        return None
    is_ref = np.array([h.samples == "REF" for h in haps], dtype=np.uint8)
    if is_ref[1:].any():
        return None
    try:
        mm, pam2 = load_tables(debug)
    except Exception:
        return None  # no model files in this environment: the reference's own function will say so
    try:
        return res.cfdon(is_ref, mm, pam2)
    except _cabi.HawkLibraryError as e:
        if e.code == _cabi.HAWK_ECFD:
            return KeyError(str(e))
        raise


def cfdon_score(guides, verbosity: int, debug: bool):
    """scoring.py:352-387. On a list that came from this package's `search` (and was annotated
    through the seam) the scores are already computed; any other list goes to the reference."""
    link = getattr(guides, "hawk", None)
    col = link.get("cfdon") if link is not None else None
    if col is None:
        fn = _reference.get("cfdon_score")
        if fn is None:
            raise RuntimeError("crispr_hawk_b200.scoring.cfdon_score: not a crispr_hawk_b200 guide list and no reference "
                               "implementation installed (there is no CPU path here)")  # fmt: skip
        return fn(guides, verbosity, debug)
    if isinstance(col, Exception):
        exception_handler(error_class("CrisprHawkCfdScoreError"), "CFDon score calculation failed", 65, debug, col)
    order = link["order"]

    def apply(g, i):
        g.cfdon_score = float(col[i])  # the setter rounds to 4 places and prints, NaN -> "NA" (guide.py:468-488)

    from .annotation import _stage

    return _stage(guides, None, order, apply)


# --------------------------------------------------------------------------- scorer inputs (N4)
# scoring.py:50-84: `_extract_guide_sequences` / `_extract_guide_sequences_sgdesigner` slice and
# upper-case every guide's sequence in a Python loop before Azimuth, RS3, DeepCpf1, CRISPRon and
# sgDesigner are called. On a list that came from this package's `search` the whole batch was cut
# on the device (`hawk_result_featurize`) while the table was resident; the models stay the
# reference's host code.
def scorer_systems():
    """Cas systems whose guides `scoring.scoring_guides` feeds to a learned scorer
    (scoring.py:845-857: SpCas9 / xCas9 -> Azimuth, RS3, CRISPRon, sgDesigner; Cpf1 -> DeepCpf1)."""
    from . import pam as mirror

    return tuple(cas9_systems()) + (_reference.get("cpf1_system", mirror.CPF1),)


def kmer_columns(link):
    """{lead: (n, L) uint8} for lead 4 and 0, rows in emission order; None without a table or for
    a PAM no scorer takes (the strings would never be asked for)."""
    res, pam = link.get("res"), link.get("pam")
    if res is None or not hasattr(res, "featurize") or getattr(pam, "cas_system", None) not in scorer_systems():
        return None
    return {lead: res.featurize(lead=lead)[0] for lead in (4, 0)}


def _sequences(guides, lead: int, name: str):
    link = getattr(guides, "hawk", None)
    cols = link.get("kmers") if link is not None else None
    if cols is None or len(link["order"]) != len(guides):
        fn = _reference.get(name)
        if fn is None:
            raise RuntimeError(f"crispr_hawk_b200.scoring.{name}: not a crispr_hawk_b200 guide list and no reference "
                               "implementation installed (there is no CPU path here)")  # fmt: skip
        return fn(guides)
    k = cols[lead]
    n, L = len(guides), k.shape[1]
    flat = k[link["order"]].tobytes().decode("ascii")  # one conversion for the whole batch
    return [flat[i * L : (i + 1) * L] for i in range(n)]


def _extract_guide_sequences(guides):
    """scoring.py:50-67."""
    return _sequences(guides, 4, "_extract_guide_sequences")


def _extract_guide_sequences_sgdesigner(guides):
    """scoring.py:70-84."""
    return _sequences(guides, 0, "_extract_guide_sequences_sgdesigner")


def deepcpf1_input(res, device_ptr: int = 0):
    """scores/deepCpf1/seqdeepcpf1.py:71-92 (`preprocess`) for every row of a resident table
    (emission order): the float32 (n, 4, 34) one-hot tensor DeepCpf1 reads, as a numpy array, or
    written into `device_ptr` (device memory of n * 4 * L floats, e.g. a torch tensor's
    data_ptr()) when given. A letter other than A, C, G, T raises KeyError like the reference's
    NTENCODING lookup."""
    from . import _cabi

    try:
        return res.featurize(lead=4, kmers=False, onehot=True, onehot_device_ptr=device_ptr)[1]
    except _cabi.HawkLibraryError as e:
        if e.code == _cabi.HAWK_EFEATURE:
            raise KeyError(str(e)) from e
        raise


SEAM = ("cfdon_score", "_extract_guide_sequences", "_extract_guide_sequences_sgdesigner")
