"""GPU parity of the streamed search (hawk_search_stream / hawk_search_stream_edits): host
texts (or edit lists) in, host guide table out, haplotypes processed group by group with the
PCIe copies overlapped. The table must be the one hawk_search gives -- the reference's golden
guides -- whatever the number of groups."""

import numpy as np
import pytest

from crispr_hawk_b200 import _cabi, marshal, synth
from crispr_hawk_b200.pam import pam_patterns
from crispr_hawk_b200.workload import Workload
from tests.helpers import all_golden_cases, fixture_objects, golden_guides, table_to_tuples

pytestmark = pytest.mark.gpu

PHASED = [c for c in all_golden_cases() if not (c["variants_present"] and not c["phased"])]
COLS = ("hap", "strand", "pos", "start", "stop", "bucket")


def _stream_case(case, haps, region, n_groups, pinned=False):
    fwd, rc = pam_patterns(case["pam"])
    params = _cabi.make_params(fwd, rc, case["guidelen"], case["right"], False)
    texts = [h.sequence.sequence for h in haps]
    buf, off, lens = marshal.stage_ascii(texts)
    seg = marshal.segment_table(haps)
    bounds = [marshal.scan_bounds(h, region.start, region.stop, len(fwd)) for h in haps]
    a = np.array([b[0] for b in bounds], np.int32)
    b = np.array([b[1] for b in bounds], np.int32)
    is_ref = np.array([h.samples == "REF" for h in haps], np.uint8)
    ctx = _cabi.Context.default()
    return _cabi.search_stream(ctx, buf, off, lens, seg, params, a, b, is_ref, n_groups=n_groups, pinned=pinned)


@pytest.mark.parametrize("case", PHASED, ids=[c["name"] for c in PHASED])
def test_streamed_search_matches_reference_golden(case):
    region, haps = fixture_objects(case)
    if not haps:
        pytest.skip("no haplotypes")
    P = len(case["pam"])
    want = golden_guides(case)
    for n_groups in (1, 2, 3, 0):
        res = _stream_case(case, haps, region, n_groups)
        got = table_to_tuples(res.table(), haps, case["guidelen"], P, case["right"])
        assert got == want, f"n_groups={n_groups}"


def test_streamed_search_ref_not_first_and_no_ref():
    """REF in the middle of the batch (served by one group) and batches without REF."""
    case = [c for c in PHASED if len(c["haps"]) >= 4 and c["haps"][0]["samples"] == "REF"][0]
    region, haps = fixture_objects(case)
    fwd, _ = pam_patterns(case["pam"])
    ctx = _cabi.Context.default()
    for order in ([1, 2, 0] + list(range(3, len(haps))), list(range(1, len(haps)))):
        sub = [haps[i] for i in order]
        for n_groups in (1, 3):
            res = _stream_case(case, sub, region, n_groups)
            # the two-call path on the same list
            buf, off, lens = marshal.stage_ascii([h.sequence.sequence for h in sub])
            batch = _cabi.Batch(ctx, buf, off, lens)
            batch.set_posmap(marshal.segment_table(sub))
            bounds = [marshal.scan_bounds(h, region.start, region.stop, len(fwd)) for h in sub]
            params = _cabi.make_params(*pam_patterns(case["pam"]), case["guidelen"], case["right"], False)
            ref = _cabi.search(ctx, batch, params, [b[0] for b in bounds], [b[1] for b in bounds],
                               [h.samples == "REF" for h in sub])  # fmt: skip
            want = ref.table()
            got = res.table()
            for k in COLS:
                assert np.array_equal(got[k], want[k]), (order[:3], n_groups, k)
            assert np.array_equal(got["text"], want["text"])
            assert res.n_hits == ref.n_hits and res.scanned_bp == ref.scanned_bp
            ref.close()
            batch.close()


def test_streamed_search_errors():
    case = PHASED[2]
    region, haps = fixture_objects(case)
    # a non-IUPAC character is reported in the caller's slot numbering
    texts = [h.sequence.sequence for h in haps]
    buf, off, lens = marshal.stage_ascii(texts)
    last = len(haps) - 1
    buf = buf.copy()
    buf[off[last] + 7] = ord("!")
    fwd, rc = pam_patterns(case["pam"])
    params = _cabi.make_params(fwd, rc, case["guidelen"], case["right"], False)
    seg = marshal.segment_table(haps)
    a = np.zeros(len(haps), np.int32)
    b = lens.astype(np.int32)
    is_ref = np.array([h.samples == "REF" for h in haps], np.uint8)
    ctx = _cabi.Context.default()
    with pytest.raises(_cabi.HawkLibraryError) as ei:
        _cabi.search_stream(ctx, buf, off, lens, seg, params, a, b, is_ref, n_groups=2)
    assert ei.value.code == _cabi.HAWK_EIUPAC and ei.value.bad_slot == off[last] + 7
    # unphased searches are refused
    up = _cabi.make_params(fwd, rc, case["guidelen"], case["right"], True)
    with pytest.raises(_cabi.HawkLibraryError) as ei:
        _cabi.search_stream(ctx, marshal.stage_ascii(texts)[0], off, lens, seg, up, a, b, is_ref)
    assert ei.value.code == _cabi.HAWK_EINVAL
    # too small an output: the wrapper grows the buffers and repeats
    small = _cabi.alloc_table(1, _cabi.text_stride(params))
    res = _cabi.search_stream(ctx, marshal.stage_ascii(texts)[0], off, lens, seg, params, a, b, is_ref, buffers=small)
    full = _cabi.search_stream(ctx, marshal.stage_ascii(texts)[0], off, lens, seg, params, a, b, is_ref)
    assert res.n_guides == full.n_guides > 1
    for k in COLS:
        assert np.array_equal(res.table()[k], full.table()[k])


@pytest.mark.parametrize("name,scale,n_alt", [("c2", 0.2, 23), ("c3", 0.1, 40), ("c5shard", 0.004, 7)])
def test_streamed_workload_equals_resident(name, scale, n_alt):
    k = synth.CONFIGS[name]
    c = synth.config_cohort(name, scale, n_alt_hap=n_alt)
    wl = Workload(c, k["pam"], k["guidelen"], k["right"])
    res = wl.step_resident()
    want = res.table()
    n_hits = res.n_hits
    res.close()
    stride = (wl.guidelen + len(wl.fwd) + 20 + 15) // 16 * 16
    for n_groups in (7, 2, 0, 1):
        for step in (wl.step_host, wl.step_edits):
            got, h2d, d2h = step(n_groups=n_groups)
            for col in COLS:
                assert np.array_equal(got[col], want[col]), (step.__name__, n_groups, col)
            assert np.array_equal(got["text"], want["text"]), (step.__name__, n_groups)
            assert d2h >= len(want["hap"]) * (21 + stride)
        assert wl.last_stream.n_hits == n_hits and wl.last_stream.scanned_bp == wl.scanned_bp
    assert h2d < wl.d.total_slots // 2  # edit lists (one group): a fraction of the texts crosses PCIe
    got, h2d, d2h = wl.step_host_twocall()
    assert h2d > wl.d.total_slots
    for col in COLS:
        assert np.array_equal(got[col], want[col])
    got, _, _ = wl.step_edits_twocall()
    assert np.array_equal(got["bucket"], want["bucket"]) and np.array_equal(got["text"], want["text"])


def test_device_resident_merge_single_rank():
    """shard.merge_tables_device with one rank: zero-copy views of the library's columns and
    hawk_first_seen_dev give the table's own bucket ids back."""
    import torch

    from crispr_hawk_b200 import shard

    k = synth.CONFIGS["c2"]
    c = synth.config_cohort("c2", 0.05, n_alt_hap=12)
    wl = Workload(c, k["pam"], k["guidelen"], k["right"])
    res = wl.step_resident()
    want = res.table()
    dev = f"cuda:{torch.cuda.current_device()}"
    m = shard.merge_tables_device(res, wl.ctx, 0, 0, 1, dev, c.region_start, c.region_stop - c.region_start + 1)
    for col in COLS:
        assert np.array_equal(m[col].cpu().numpy(), want[col]), col
    assert np.array_equal(m["text"].cpu().numpy()[:, : want["text"].shape[1]], want["text"])
    res.close()


def test_first_seen_rejects_out_of_range_start():
    """hawk_first_seen_dev with a start outside [key_min, key_min + key_span): HAWK_EINVAL, no
    write outside the key table (the words behind it keep their canary)."""
    import ctypes as C

    import torch

    lib = _cabi.load_library()
    dev = f"cuda:{torch.cuda.current_device()}"
    key_min, key_span = 1000, 64
    start = torch.tensor([1000, 1005, 1063, 1005], dtype=torch.int32, device=dev)
    strand = torch.tensor([0, 1, 0, 1], dtype=torch.uint8, device=dev)
    table = torch.full((2 * key_span + 1 + 64,), 0x5A5A5A5A, dtype=torch.int32, device=dev)
    bucket = torch.zeros(4, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def run(s):
        return lib.hawk_first_seen_dev(C.c_void_p(stream), C.c_void_p(s.data_ptr()), C.c_void_p(strand.data_ptr()), 4, key_min,
                                       key_span, C.c_void_p(table.data_ptr()), C.c_void_p(bucket.data_ptr()))  # fmt: skip

    assert run(start) == _cabi.HAWK_OK
    assert bucket.cpu().tolist() == [0, 1, 2, 1]
    for bad in (999, 1064, 1000 + (1 << 20), -5):
        s2 = start.clone()
        s2[2] = bad
        assert run(s2) == _cabi.HAWK_EINVAL
        assert table[2 * key_span + 1 :].cpu().eq(0x5A5A5A5A).all()
