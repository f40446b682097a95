"""CPU check of the kernels' bit logic: the __host__ __device__ core the CUDA
kernels are built from (hawk_core.h), compiled for the host, against the golden
vectors. Covers K1's SWAR packing, the per-chunk scan with fused filters, the
run-length posmap, REF-core comparison and IUPAC resolution. The block-level
parts of the kernels need a GPU (tests/test_gpu_parity.py)."""

import numpy as np
import pytest

from oracle import hawk_oracle as O
from tests import hostcheck
from tests.helpers import all_golden_cases, fixture_objects, golden_guides, split_hits, table_to_tuples

CASES = all_golden_cases()


def test_pack_matches_oracle_encode():
    rng = np.random.default_rng(7)
    alphabet = "ACGTNRYSWKMBDHVacgtnryswkmbdhv"
    texts = ["".join(rng.choice(list(alphabet), size=n)) for n in (0, 1, 31, 32, 33, 127, 128, 129, 1000)]
    q, v, off, lens, bad = hostcheck.pack(texts)
    assert bad == -1
    for h, t in enumerate(texts):
        nib, low = hostcheck.unpack_nibbles(q, v, off, lens, h)
        assert nib.tolist() == O.encode(t)
        assert low.tolist() == [int(c.islower()) for c in t]
    # unused slots stay zero
    used = np.zeros(len(v) * 32, bool)
    for h, t in enumerate(texts):
        used[off[h] : off[h] + len(t)] = True
    planes = q.reshape(-1, 4)
    for k in range(4):
        bits = ((planes[:, k][:, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(-1).astype(bool)
        assert not bits[~used].any()


def test_pack_flags_first_invalid_slot():
    q, v, off, lens, bad = hostcheck.pack(["ACGT" * 50, "ACGTXACGT!"])
    assert bad == off[1] + 4


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_core_matches_golden(case):
    region, haps = fixture_objects(case)
    G, P = case["guidelen"], len(case["pam"])
    raw = hostcheck.search(case["pam"], region, haps, G, case["right"], case["variants_present"],
                           case["phased"], raw=True)  # fmt: skip
    got_hits = [[f, r] for f, r in zip(split_hits(raw[0], len(haps)), split_hits(raw[1], len(haps)))]
    assert got_hits == case["pam_hits"]
    tab = hostcheck.search(case["pam"], region, haps, G, case["right"], case["variants_present"], case["phased"])
    assert table_to_tuples(tab, haps, G, P, case["right"]) == golden_guides(case)


LONG = CASES[:4] + CASES[4:36:4] + CASES[36:55:5]


@pytest.mark.parametrize("case", LONG, ids=[c["name"] for c in LONG])
@pytest.mark.parametrize("guidelen", [33, 60])
def test_core_long_guides_match_oracle(case, guidelen):
    """Guides longer than 32 nt leave the kernels' 96-bit fast path (K.small == 0): the generic
    sliding-window code must agree with the oracle too. (Hits near the region ends disappear
    because the fixed 100-bp padding no longer covers guide + pad -- same in the reference.)"""
    region, haps = fixture_objects(case)
    P = len(case["pam"])
    tab = hostcheck.search(case["pam"], region, haps, guidelen, case["right"], case["variants_present"], case["phased"])
    want = O.search(case["pam"], region.start, region.stop, [O.OracleHap.from_object(h) for h in haps], guidelen,
                    case["right"], case["variants_present"], case["phased"])  # fmt: skip
    assert table_to_tuples(tab, haps, guidelen, P, case["right"]) == want


def test_pack_every_byte_value():
    """K1's boolean network against the table for all 256 byte values (NUL = unused slot)."""
    valid = "ACGTNRYSWKMBDHV"
    buf = np.zeros(256 + 128, np.uint8)
    buf[128:384] = np.arange(256, dtype=np.uint8)
    q = np.zeros((len(buf) // 32 + 8) * 4, np.uint32)
    v = np.zeros(len(buf) // 32 + 8, np.uint32)
    import ctypes as C

    bad = hostcheck.lib().hawkcheck_pack(buf.ctypes.data_as(C.c_void_p), C.c_int64(len(buf)),
                                         q.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))  # fmt: skip
    assert bad == 128 + 1  # first non-IUPAC, non-NUL byte
    planes = q.reshape(-1, 4)
    for b in range(256):
        slot = 128 + b
        nib = sum(((int(planes[slot // 32, k]) >> (slot % 32)) & 1) << k for k in range(4))
        low = (int(v[slot // 32]) >> (slot % 32)) & 1
        ch = chr(b)
        if ch.upper() in valid and ch.isalpha() and b < 128:
            assert nib == O.IUPAC_BITS[ch.upper()] and low == int(ch.islower()), b
        else:
            assert nib == 0 and low == 0, b


def _annot_cases():
    import gzip
    import json
    import os

    from tests.helpers import GOLDEN_DIR

    with gzip.open(os.path.join(GOLDEN_DIR, "annot.json.gz"), "rb") as fh:
        annot = json.loads(fh.read().decode())
    return [(c, annot[c["name"]]) for c in all_golden_cases() if c["name"] in annot]


@pytest.mark.parametrize("case,want", _annot_cases(), ids=[c["name"] for c, _ in _annot_cases()])
def test_n2_row_logic_on_cpu_matches_reference_annotation(case, want):
    """The per-row code the N2 kernels execute (annot_row_variants / annot_text_byte /
    annot_gc_counts in hawk_core.h), compiled for the CPU, against the reference's own
    annotation outputs."""
    from crispr_hawk_b200.annotation import format_af

    region, haps = fixture_objects(case)
    if not haps:
        pytest.skip("no haplotypes")
    tab = hostcheck.search(case["pam"], region, haps, case["guidelen"], case["right"], case["variants_present"],
                           case["phased"])  # fmt: skip
    ann = hostcheck.annotate(tab, haps, case["pam"], case["guidelen"], case["right"])
    order = np.argsort(tab["bucket"], kind="stable").tolist()
    assert len(order) == len(want)
    right = case["right"]
    for k, i in enumerate(order):
        h = haps[int(tab["hap"][i])]
        s = int(tab["strand"][i])
        if h.variants == "NA":
            v, afs = "NA", "NA"
        else:
            ids = sorted(ann["vt"].ids[int(tab["hap"][i])][j] for j in ann["gv_idx"][ann["gv_off"][i] : ann["gv_off"][i + 1]])
            v = ",".join(ids)
            vals = [format_af(h.afs[x]) if str(h.afs[x]) != "nan" else "NA" for x in v.split(",")]
            afs = "NA" if not vals or (len(set(vals)) == 1 and vals[0] == "NA") else ",".join(vals)
        rp = ((not right) if s == 1 else bool(right)) != (s == 1)
        got = [v, afs, ann["rc_text"][i].tobytes().decode("ascii"), rp, str(int(ann["gc_num"][i]) / int(ann["gc_den"][i]))]
        assert got == want[k], f"guide {k}"


def test_v3_pack_equals_pack_on_every_byte(fn="hawkcheck_pack_v3_diff"):
    """pack_chunk_v3 (the kernels' K1) vs pack_chunk: all 256 byte values in every
    lane position, random mixes of IUPAC letters / NUL / junk."""
    import ctypes as C

    import numpy as np

    from tests import hostcheck

    lib = hostcheck.lib()
    getattr(lib, fn).restype = C.c_int64
    rng = np.random.default_rng(5)
    letters = np.frombuffer(b"ACGTRYSWKMBDHVNacgtryswkmbdhvn", np.uint8)
    parts = []
    for lane in range(32):  # every byte value at every position of an otherwise valid chunk
        blk = letters[rng.integers(0, len(letters), (256, 32))]
        blk[:, lane] = np.arange(256)
        parts.append(blk)
    parts.append(letters[rng.integers(0, len(letters), (4096, 32))])
    mix = letters[rng.integers(0, len(letters), (4096, 32))]
    mix[rng.random(mix.shape) < 0.2] = 0
    parts.append(mix)
    parts.append(rng.integers(0, 256, (4096, 32)).astype(np.uint8))
    parts.append(np.zeros((4, 32), np.uint8))
    buf = np.ascontiguousarray(np.concatenate(parts).reshape(-1))
    assert getattr(lib, fn)(buf.ctypes.data_as(C.c_void_p), C.c_int64(len(buf) // 32)) == 0


def test_planes_to_chars_gives_the_letters_back():
    """planes_to_chars32 (the text column of gather_fast): every IUPAC letter of either case and
    the unused slot in every lane position, random mixes, every valid-prefix length."""
    import ctypes as C

    import numpy as np

    from tests import hostcheck

    lib = hostcheck.lib()
    lib.hawkcheck_chars32_diff.restype = C.c_int64
    rng = np.random.default_rng(9)
    letters = np.frombuffer(b"ACGTRYSWKMBDHVNacgtryswkmbdhvn\0", np.uint8)
    parts = []
    for lane in range(32):
        blk = letters[rng.integers(0, len(letters), (len(letters), 32))]
        blk[:, lane] = letters
        parts.append(blk)
    parts.append(letters[rng.integers(0, len(letters), (4096, 32))])
    buf = np.ascontiguousarray(np.concatenate(parts).reshape(-1))
    for keep in list(range(0, 33)):
        assert lib.hawkcheck_chars32_diff(buf.ctypes.data_as(C.c_void_p), C.c_int64(len(buf) // 32), C.c_int32(keep)) == 0, keep
