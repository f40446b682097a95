"""N4, second half (the learned scorers' inputs): hawk_result_featurize against what the unmodified
reference computed (tests/golden/features.json.gz, generator: tests/golden/make_golden_features.py)
-- `scoring._extract_guide_sequences[_sgdesigner]` strings guide by guide, DeepCpf1's `preprocess`
tensor by shape, SHA-256 of its float32 bytes and first row, and the KeyError it raises on a letter
other than A, C, G, T. CPU: the kernels' own feature_byte / onehot_channel compiled for the host
(tests/fake_backend.py); tests/test_gpu_features.py: the GPU."""

import hashlib

import numpy as np
import pytest

import crispr_hawk_b200 as hawk
from crispr_hawk_b200 import _cabi, scoring
from tests import fake_backend
from tests.helpers import fixture_objects, load_golden

FEAT = load_golden("features")


def feature_rows(case):
    region, haps = fixture_objects(case)
    packed = hawk.encode_region(haps, 0, True)
    pam = hawk.PAM(case["pam"], case["right"], True)
    pam.encode(0)
    table, res = hawk.search_table(pam, region, haps, packed, case["guidelen"], case["right"], True, case["phased"], 0, True)
    k4, _ = res.featurize(lead=4)
    k0, _ = res.featurize(lead=0)
    try:
        onehot = scoring.deepcpf1_input(res)
    except KeyError:
        onehot = "KeyError"
    res.close()
    order = np.argsort(table["bucket"], kind="stable")
    rows = [[int(table["start"][i]), int(table["strand"][i]), haps[int(table["hap"][i])].id, k4[i].tobytes().decode(),
             k0[i].tobytes().decode()] for i in order]  # fmt: skip
    return rows, (onehot if isinstance(onehot, str) else onehot[order])


def check(case):
    rows, onehot = feature_rows(case)
    want = case["rows"]
    G, P = case["guidelen"], len(case["pam"])
    assert len(rows) == len(want) > 50
    for k, (g, w) in enumerate(zip(rows, want)):
        assert g == w, f"guide {k}: {g} != {w}"
        assert len(g[3]) == G + P + 7 and len(g[4]) == G + P + 3 and g[3][4:] == g[4]
    if case["onehot"] == "KeyError":
        assert onehot == "KeyError"
    else:
        assert onehot.dtype == np.float32 and list(onehot.shape) == case["onehot"]["shape"]
        assert onehot[0].astype(int).tolist() == case["onehot"]["row0"]
        assert hashlib.sha256(np.ascontiguousarray(onehot).tobytes()).hexdigest() == case["onehot"]["sha256"]
    assert any(r[1] == 1 for r in rows) and any(r[1] == 0 for r in rows)


@pytest.mark.parametrize("case", FEAT["cases"], ids=[c["name"] for c in FEAT["cases"]])
def test_features_match_reference(case, monkeypatch):
    fake_backend.activate(monkeypatch)
    check(case)


def test_error_code_is_declared():
    assert _cabi.HAWK_EFEATURE == -10
