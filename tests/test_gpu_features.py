"""GPU parity of N4's scorer inputs (hawk_result_featurize) against the reference's own strings and
one-hot tensors (tests/golden/features.json.gz); the one-hot written straight into device memory;
a workload-scale run checked against numpy on the fetched window texts."""

import numpy as np
import pytest
import torch

from crispr_hawk_b200 import _cabi, synth
from crispr_hawk_b200.workload import Workload
from tests.test_features import FEAT, check

pytestmark = pytest.mark.gpu

COMP = bytes.maketrans(b"ACGTacgtRYKMBDHVrykmbdhv", b"TGCAtgcaYRMKVHDByrmkvhdb")


@pytest.mark.parametrize("case", FEAT["cases"], ids=[c["name"] for c in FEAT["cases"]])
def test_features_match_reference(case):
    check(case)


@pytest.mark.parametrize("name", ["c2", "c3"])
def test_features_at_scale(name):
    k = synth.CONFIGS[name]
    c = synth.config_cohort(name, 0.1, n_alt_hap=1200)
    check_against_numpy(Workload(c, k["pam"], k["guidelen"], k["right"]), 50_000)


@pytest.mark.parametrize("pam,G,right", [("NGG", 1, False), ("TTTV", 70, True), ("NGG", 125, False), ("NNGRRT", 33, True)])
def test_features_other_geometries(pam, G, right):
    """One-base guides, windows of several 16-byte text words, the longest window the library
    takes (G + P = 128: 148 characters, 135-letter scorer strings)."""
    c = synth.make_cohort(bed_len=60_000, n_alt_hap=9, n_sites=900, mean_alts_per_hap=150, seed=12,
                          snv_frac=0.6, ins_frac=0.2, max_indel=8)  # fmt: skip
    check_against_numpy(Workload(c, pam, G, right), 300)


def check_against_numpy(wl, min_rows):
    res = wl.step_resident()
    table = res.table()
    n, W = len(table["hap"]), res.window
    assert n > min_rows
    k4, _ = res.featurize(lead=4)
    k0, _ = res.featurize(lead=0)
    L = W - 20 + 7
    assert k4.shape == (n, L) and k0.shape == (n, L - 4) and np.array_equal(k4[:, 4:], k0)
    # numpy restatement of scoring.py:50-67 on the fetched window texts
    text = table["text"][:, :W]
    fwd = text & np.uint8(0xDF)
    lut = np.arange(256, dtype=np.uint8)
    lut[list(b"ACGTRYKMBDHV")] = list(b"TGCAYRMKVHDB")
    rev = lut[fwd[:, ::-1]]
    full = np.where((table["strand"] == 1)[:, None], rev, fwd)
    assert np.array_equal(k4, full[:, 6 : W - 7])
    # one-hot straight into a torch tensor on the device == the host copy == one_hot of the strings
    dev = torch.empty((n, 4, L), dtype=torch.float32, device="cuda")
    _, none = res.featurize(lead=4, kmers=False, onehot=True, onehot_device_ptr=dev.data_ptr())
    assert none is None
    _, host = res.featurize(lead=4, kmers=False, onehot=True)
    assert np.array_equal(dev.cpu().numpy(), host)
    code = np.full(256, -1, np.int64)
    code[list(b"ACGT")] = range(4)
    want = (code[k4][:, None, :] == np.arange(4)[None, :, None]).astype(np.float32)
    assert np.array_equal(host, want) and host.sum() == n * L
    res.close()


def test_onehot_refuses_other_letters():
    c = synth.make_cohort(bed_len=20_000, n_alt_hap=5, n_sites=300, mean_alts_per_hap=40, seed=3)
    wl = Workload(c, "NGG", 20, False)
    buf = wl.ascii_dev.cpu().numpy().copy()
    off = int(wl.d.slot_off[0])
    buf[off + 5_000] = ord("N")
    wl.ascii_dev.copy_(torch.from_numpy(buf))
    res = wl.step_resident()
    table = res.table()
    k4, _ = res.featurize(lead=4)
    has_n = (k4 == ord("N")).any(axis=1)
    assert has_n.any()
    with pytest.raises(_cabi.HawkLibraryError) as ei:
        res.featurize(lead=4, kmers=False, onehot=True)
    assert ei.value.code == _cabi.HAWK_EFEATURE and ei.value.bad_row == int(np.flatnonzero(has_n)[0])
    assert len(table["hap"]) == len(k4)
    res.close()
