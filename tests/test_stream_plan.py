"""Host logic of the streamed search (no GPU): how hawk_search_stream cuts the haplotypes of a
batch into groups (csrc/stream_api.cu make_plan, exported as hawk_stream_plan)."""

import numpy as np
import pytest

from crispr_hawk_b200 import _cabi, marshal


def _layout(lens):
    off, _ = marshal.layout(np.asarray(lens, np.int32))
    return off


def _check_cover(groups, first, n_hap):
    assert groups[0][0] == first and groups[-1][1] == n_hap
    for (a, b), (c, d) in zip(groups, groups[1:]):
        assert b == c and a < b and c < d


def test_ref_first_groups_cover_the_other_haplotypes_in_order():
    lens = [1000] + [1000 + 7 * i for i in range(40)]
    off = _layout(lens)
    is_ref = [1] + [0] * 40
    for k in (1, 2, 3, 7, 40):
        g = _cabi.stream_plan(off, is_ref, k)
        assert k - 1 <= len(g) <= k and (k > 7 or len(g) == k)  # k is a target: blocks are whole haplotypes
        _check_cover(g, 1, 41)
    # more groups than haplotypes: one haplotype each
    g = _cabi.stream_plan(off, is_ref, 100)
    assert 39 <= len(g) <= 40 and all(1 <= b - a <= 2 for a, b in g)
    # balanced by slot count
    g = _cabi.stream_plan(off, is_ref, 4)
    sizes = [off[b] - off[a] for a, b in g]
    assert max(sizes) - min(sizes) <= 2 * (off[41] - off[40])


def test_automatic_group_count_follows_the_text_size():
    off = _layout([1_000_000] * 6)  # 5 MB of text: one group
    assert _cabi.stream_plan(off, [1, 0, 0, 0, 0, 0]) == [(1, 6)]
    big = np.concatenate(([128], 128 + np.cumsum(np.full(2000, 1_000_064 + 128, np.int64))))  # ~2 GB, layout rule
    g = _cabi.stream_plan(big, [1] + [0] * 1999)
    assert 8 <= len(g) <= 12
    _check_cover(g, 1, 2000)


def test_ref_elsewhere_or_absent_and_errors():
    off = _layout([500] * 5)
    assert _cabi.stream_plan(off, [0, 0, 1, 0, 0], 3) == [(0, 5)]  # REF in the middle: the batch as it stands
    g = _cabi.stream_plan(off, [0, 0, 0, 0, 0], 2)  # no REF: plain blocks
    _check_cover(g, 0, 5)
    assert _cabi.stream_plan(off[:2], [1], 4) == [(1, 1)]  # REF alone
    assert _cabi.stream_plan(_layout([]), [], 4) == []
    with pytest.raises(_cabi.HawkLibraryError) as ei:
        _cabi.stream_plan(off, [1, 0, 1, 0, 0], 2)
    assert ei.value.code == _cabi.HAWK_EDUPREF
