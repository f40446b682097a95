"""Edit-list batches (N1) build their planes only around the edits (edits_kernels.cu,
hawk_ctx_set_edit_planes mode 1); mode 0 materialises every haplotype text and runs K1. Both
must give the same tables, hit lists, nibbles and annotations -- on edit lists meaner than the
bench's: adjacent edits, edits at the very ends, insertions longer than one and two chunks,
haplotypes with no edits that are not REF, a REF with edits."""

import numpy as np
import pytest

from crispr_hawk_b200 import _cabi
from crispr_hawk_b200.pam import pam_patterns

pytestmark = pytest.mark.gpu

BASES = np.frombuffer(b"ACGT", np.uint8)


def params_of(pam, G, right):
    fwd, rc = pam_patterns(pam)
    return _cabi.make_params(fwd, rc, G, right, False)


def random_edits(rng, ref, n_hap, mean_gap, long_ins=False, touch_ends=False):
    """CSR edit lists: sorted, non-overlapping SNVs / anchored insertions / anchored deletions."""
    L = len(ref)
    off, pos, rl, al, ao, pool = [0], [], [], [], [], []
    for h in range(n_hap):
        if h == 0 or (h == 3 and n_hap > 4):  # REF, and one more haplotype identical to it
            off.append(len(pos))
            continue
        p = 0 if touch_ends else int(rng.integers(1, 40))
        while p < L:
            kind = rng.random()
            if kind < 0.55:
                r, a = 1, 1
            elif kind < 0.8:
                r, a = 1, 1 + int(rng.integers(1, 90 if long_ins and rng.random() < 0.2 else 9))
            else:
                r, a = 1 + int(rng.integers(1, 9)), 1
            if p + r > L:
                break
            text = BASES[rng.integers(0, 4, a)].copy()
            if r == 1 and a == 1:
                text[0] = BASES[(np.searchsorted(BASES, ref[p]) + rng.integers(1, 4)) % 4]
            else:
                text[0] = ref[p]
            if rng.random() < 0.3:
                text |= 0x20  # the pool may hold either case
            pos.append(p), rl.append(r), al.append(a), ao.append(len(pool))
            pool.extend(text.tolist())
            p += r + (0 if rng.random() < 0.15 else int(rng.geometric(1.0 / mean_gap)))
        if touch_ends and pos and pos[-1] + rl[-1] < L and len(pos) > off[-1]:
            pos.append(L - 1), rl.append(1), al.append(1), ao.append(len(pool))
            pool.append(int(BASES[(np.searchsorted(BASES, ref[L - 1]) + 1) % 4]))
        off.append(len(pos))
    return (np.array(off, np.int64), np.array(pos, np.int32), np.array(rl, np.int32), np.array(al, np.int32),
            np.array(ao, np.int64), np.array(pool if pool else [65], np.uint8))  # fmt: skip


def run(ctx, mode, ref, edits, pam, G, right, is_ref, want=("table", "hits", "nibbles")):
    ctx.set_edit_planes(mode)
    b = _cabi.Batch.from_edits(ctx, ref, 5000, *edits)
    params = params_of(pam, G, right)
    a = np.zeros(b.n_hap, np.int32)
    e = (b.lens - len(pam)).astype(np.int32)
    out = {"lens": b.lens.copy()}
    res = _cabi.search(ctx, b, params, a, e, is_ref)
    out["table"] = res.table()
    out["hits"] = [res.hits(0), res.hits(1)]
    out["annot"] = res.annotate(b)
    res.close()
    if "raw" in want:
        raw = _cabi.pam_search(ctx, b, params, a, e)
        out["raw"] = [raw.hits(0), raw.hits(1)]
        raw.close()
    if "nibbles" in want:
        out["nibbles"] = [b.export_nibbles(h) for h in range(b.n_hap)]
    b.close()
    return out


def same(x, y, what):
    if isinstance(x, dict):
        assert x.keys() == y.keys(), what
        for k in x:
            same(x[k], y[k], f"{what}.{k}")
    elif isinstance(x, (list, tuple)):
        assert len(x) == len(y), what
        for i, (p, q) in enumerate(zip(x, y)):
            same(p, q, f"{what}[{i}]")
    else:
        assert np.array_equal(np.asarray(x), np.asarray(y)), f"{what} differs"


GEOMETRIES = [("NGG", 20, False), ("TTTV", 23, True), ("NNGRRT", 21, False), ("NGG", 32, False), ("NGG", 3, True)]


@pytest.mark.parametrize("seed,mean_gap,long_ins,touch_ends", [(1, 60, False, False), (2, 9, False, True), (3, 150, True, False),
                                                                (4, 25, True, True), (5, 700, False, False)])  # fmt: skip
def test_window_planes_equal_materialised_texts(seed, mean_gap, long_ins, touch_ends):
    rng = np.random.default_rng(seed)
    ref = BASES[rng.integers(0, 4, int(rng.integers(3000, 9000)))]
    n_hap = 9
    edits = random_edits(rng, ref, n_hap, mean_gap, long_ins, touch_ends)
    is_ref = np.zeros(n_hap, np.uint8)
    is_ref[0] = 1
    ctx = _cabi.Context.default()
    try:
        for pam, G, right in GEOMETRIES:
            got = run(ctx, 1, ref, edits, pam, G, right, is_ref)
            want = run(ctx, 0, ref, edits, pam, G, right, is_ref)
            same(got, want, f"{pam}/{G}")
            assert len(want["table"]["hap"]) > 0
    finally:
        ctx.set_edit_planes(1)


def test_whole_haplotype_readers_build_every_plane_first():
    """Long guides, a REF that carries edits, the raw PAM scan and the nibble export on a batch
    that has only window planes (or none yet): each builds all planes on its own."""
    rng = np.random.default_rng(11)
    ref = BASES[rng.integers(0, 4, 6000)]
    edits = random_edits(rng, ref, 6, 80, True, False)
    ctx = _cabi.Context.default()
    try:
        is_ref = np.zeros(6, np.uint8)
        is_ref[0] = 1
        for pam, G, right in (("NGG", 40, False), ("TTTV", 70, True)):  # beyond the scan's fast form
            same(run(ctx, 1, ref, edits, pam, G, right, is_ref, want=("raw",)), run(ctx, 0, ref, edits, pam, G, right, is_ref, want=("raw",)), pam)
        odd = np.zeros(6, np.uint8)
        odd[2] = 1  # a haplotype with edits scanned like REF
        same(run(ctx, 1, ref, edits, "NGG", 20, False, odd), run(ctx, 0, ref, edits, "NGG", 20, False, odd), "REF with edits")
        # nibbles / raw scan before any search, and a second search with a longer reach on the same batch
        for mode in (1, 0):
            ctx.set_edit_planes(mode)
            b = _cabi.Batch.from_edits(ctx, ref, 5000, *edits)
            nib = [b.export_nibbles(h) for h in range(b.n_hap)]
            b.close()
            b = _cabi.Batch.from_edits(ctx, ref, 5000, *edits)
            a, e = np.zeros(6, np.int32), (b.lens - 3).astype(np.int32)
            tabs = []
            for G in (5, 30, 20):
                r = _cabi.search(ctx, b, params_of("NGG", G, False), a, e, is_ref)
                tabs.append(r.table())
                r.close()
            b.close()
            if mode == 1:
                first = (nib, tabs)
            else:
                same(first, (nib, tabs), "reuse")
    finally:
        ctx.set_edit_planes(1)


def test_soft_masked_reference_takes_the_text_path():
    """Lower-case reference bases count as variant bases in every haplotype (search_guides.py
    :468-471 looks at the case only): such a batch is built from the texts in either mode."""
    rng = np.random.default_rng(12)
    ref = BASES[rng.integers(0, 4, 4000)].copy()
    ref[1000:1100] |= 0x20
    edits = random_edits(rng, ref, 5, 120)
    is_ref = np.array([1, 0, 0, 0, 0], np.uint8)
    ctx = _cabi.Context.default()
    try:
        got, want = run(ctx, 1, ref, edits, "NGG", 20, False, is_ref), run(ctx, 0, ref, edits, "NGG", 20, False, is_ref)
        same(got, want, "soft-masked")
        assert (got["table"]["hap"] > 0).sum() > 0
    finally:
        ctx.set_edit_planes(1)


def test_bad_alt_text_is_reported_like_the_text_path():
    rng = np.random.default_rng(13)
    ref = BASES[rng.integers(0, 4, 3000)]
    edits = list(random_edits(rng, ref, 4, 100))
    edits[5] = edits[5].copy()
    edits[5][len(edits[5]) // 2] = ord("!")
    ctx = _cabi.Context.default()
    errs = []
    for mode in (1, 0):
        ctx.set_edit_planes(mode)
        with pytest.raises(_cabi.HawkLibraryError) as ei:
            _cabi.Batch.from_edits(ctx, ref, 5000, *edits)
        errs.append((ei.value.code, ei.value.bad_slot))
    ctx.set_edit_planes(1)
    assert errs[0] == errs[1] and errs[0][0] == _cabi.HAWK_EIUPAC
