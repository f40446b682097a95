"""ctypes driver of libhawkcheck.so: the kernels' __host__ __device__ core compiled for
the CPU (crispr_hawk_b200/csrc/hostcheck.cpp). TEST SUPPORT ONLY."""

from __future__ import annotations

import ctypes as C

import numpy as np

from crispr_hawk_b200 import _cabi, build, marshal
from crispr_hawk_b200.pam import pam_patterns

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build.build_hostcheck()
        _lib = C.CDLL(path)
        _lib.hawkcheck_pack.restype = C.c_int64
        _lib.hawkcheck_search.restype = C.c_void_p
        _lib.hawkcheck_n.restype = C.c_int64
        _lib.hawkcheck_nhits.restype = C.c_int64
        for f in ("hawkcheck_n", "hawkcheck_err", "hawkcheck_window", "hawkcheck_free"):
            getattr(_lib, f).argtypes = [C.c_void_p]
        _lib.hawkcheck_nhits.argtypes = [C.c_void_p, C.c_int]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def pack(texts):
    buf, off, lens = marshal.stage_ascii(texts)
    n_chunks = len(buf) // 32 + 8
    q = np.zeros(n_chunks * 4, np.uint32)
    v = np.zeros(n_chunks, np.uint32)
    bad = lib().hawkcheck_pack(_p(buf), C.c_int64(len(buf)), _p(q), _p(v))
    return q, v, off, lens, bad


def unpack_nibbles(q, v, off, lens, h):
    n = int(lens[h])
    idx = np.arange(n) + int(off[h])
    chunk, bit = idx // 32, (idx % 32).astype(np.uint32)
    planes = q.reshape(-1, 4)
    nib = np.zeros(n, np.uint8)
    for k in range(4):
        nib |= (((planes[chunk, k] >> bit) & 1) << k).astype(np.uint8)
    low = ((v[chunk] >> bit) & 1).astype(np.uint8)
    return nib, low


def search(pam, region, haps, guidelen, right, variants_present, phased, raw=False):
    texts = [marshal.hap_text(h) for h in haps]
    fwd, rc = pam_patterns(pam)
    unphased = bool(variants_present and not phased)
    params = _cabi.make_params(fwd, rc, guidelen, right, unphased)
    bounds = [marshal.scan_bounds(h, region.start, region.stop, len(fwd)) for h in haps]
    a = np.array([b[0] for b in bounds], np.int32)
    b = np.array([b[1] for b in bounds], np.int32)
    is_ref = np.array([h.samples == "REF" for h in haps], np.uint8)
    return search_flat(texts, params, a, b, is_ref, marshal.segment_table(haps), marshal.allele_table(haps), raw)


def search_flat(texts, params, a, b, is_ref, seg, va, raw=False):
    """The same from the flat arrays the C-ABI takes (tests/fake_backend.py feeds these)."""
    q, v, off, lens, bad = pack(texts)
    assert bad == -1
    a, b = np.ascontiguousarray(a, np.int32), np.ascontiguousarray(b, np.int32)
    is_ref = np.ascontiguousarray(is_ref, np.uint8)
    haps = texts
    L = lib()
    t = L.hawkcheck_search(
        _p(q), _p(v), _p(off), _p(lens), _p(a), _p(b), _p(is_ref), C.c_int32(len(haps)),
        _p(seg.seg_off), _p(seg.seg_rel), _p(seg.seg_gen), _p(seg.seg_step),
        _p(va.va_off), _p(va.va_idx), _p(va.va_ent_off), _p(va.va_ref),
        C.byref(params), C.c_int(1 if raw else 0),
    )  # fmt: skip
    t = C.c_void_p(t)
    try:
        hits = []
        for s in (0, 1):
            n = L.hawkcheck_nhits(t, s)
            arr = np.empty(n, np.uint64)
            L.hawkcheck_fetch_hits(t, C.c_int(s), _p(arr))
            hits.append(arr)
        if raw:
            return hits
        err = L.hawkcheck_err(t)
        if err:
            raise KeyError(f"hostcheck error {err}")
        n, w = L.hawkcheck_n(t), L.hawkcheck_window(t)
        tab = {
            "hap": np.empty(n, np.int32), "strand": np.empty(n, np.uint8), "pos": np.empty(n, np.int32),
            "start": np.empty(n, np.int32), "stop": np.empty(n, np.int32),
            "bucket": np.empty(n, np.int64), "text": np.empty((n, w), np.uint8),
        }  # fmt: skip
        L.hawkcheck_fetch(t, _p(tab["hap"]), _p(tab["strand"]), _p(tab["pos"]), _p(tab["start"]),
                          _p(tab["stop"]), _p(tab["bucket"]), _p(tab["text"]))  # fmt: skip
        return tab
    finally:
        L.hawkcheck_free(t)


def annotate(table, haps, pam, guidelen, right):
    """N2 through the kernels' own per-row logic compiled for the CPU (hawkcheck_annotate):
    same columns as _cabi.Result.annotate, for a table in emission order."""
    fwd, rc = pam_patterns(pam)
    params = _cabi.make_params(fwd, rc, guidelen, right, False)
    return annotate_flat(table, params, marshal.segment_table(haps), marshal.variant_table(haps))


def annotate_flat(table, params, seg, vt):
    L = lib()
    L.hawkcheck_annotate.restype = C.c_int64
    guidelen = params.guide_len
    fwd = [0] * params.pam_len
    n, w = len(table["hap"]), table["text"].shape[1] if len(table["hap"]) else guidelen + len(fwd) + 20
    stride = (w + 15) // 16 * 16
    text = np.zeros((n, stride), np.uint8)
    text[:, :w] = table["text"]
    rc_text = np.zeros((n, stride), np.uint8)
    num, den = np.zeros(n, np.int32), np.zeros(n, np.int32)
    off = np.zeros(n + 1, np.int64)
    idx = np.zeros(max(1, n * (guidelen + len(fwd))), np.int32)
    pool = vt.alt_pool if len(vt.alt_pool) else np.zeros(1, np.uint8)
    total = L.hawkcheck_annotate(
        _p(seg.seg_off), _p(seg.seg_rel), _p(seg.seg_gen), _p(seg.seg_step), _p(vt.var_off),
        _p(vt.var_pos if len(vt.var_pos) else np.zeros(1, np.int32)),
        _p(vt.var_reflen if len(vt.var_reflen) else np.zeros(1, np.int32)),
        _p(vt.var_altlen if len(vt.var_altlen) else np.zeros(1, np.int32)),
        _p(vt.var_altoff if len(vt.var_altoff) else np.zeros(1, np.int64)), _p(pool), C.byref(params), C.c_int64(n),
        _p(np.ascontiguousarray(table["hap"], np.int32)), _p(np.ascontiguousarray(table["strand"], np.uint8)),
        _p(np.ascontiguousarray(table["pos"], np.int32)), _p(np.ascontiguousarray(table["stop"], np.int32)), _p(text),
        C.c_int32(stride), _p(rc_text), _p(num), _p(den), _p(off), _p(idx),
    )  # fmt: skip
    if total < 0:
        raise AssertionError("the reference's _find_insertion_stop assert would fire")
    return {"rc_text": rc_text[:, :w], "gc_num": num, "gc_den": den, "gv_off": off, "gv_idx": idx[:total], "vt": vt}


def cfdon_flat(table, params, is_ref, mm, pam2):
    """hawk_result_cfdon on the CPU: the kernels' own cfdon_row over the table."""
    from crispr_hawk_b200 import _cabi

    L = lib()
    L.hawkcheck_cfdon.restype = C.c_int64
    n = len(table["hap"])
    w = params.guide_len + params.pam_len + 20
    stride = (w + 15) // 16 * 16
    text = np.zeros((max(n, 1), stride), np.uint8)
    if n:
        text[:n, :w] = table["text"][:, :w]
    out = np.zeros(max(n, 1), np.float64)
    bad = L.hawkcheck_cfdon(
        _p(np.ascontiguousarray(table["hap"], np.int32)), _p(np.ascontiguousarray(table["strand"], np.uint8)),
        _p(np.ascontiguousarray(table["bucket"], np.uint32)), _p(text), C.c_int32(stride), C.c_int32(w),
        C.c_int32(params.guide_len), C.c_int32(params.pam_len), C.c_int32(params.right),
        _p(np.ascontiguousarray(is_ref, np.uint8)), _p(np.ascontiguousarray(mm, np.float64).reshape(320)),
        _p(np.ascontiguousarray(pam2, np.float64).reshape(16)), C.c_int64(n), _p(out),
    )  # fmt: skip
    if bad >= 0:
        err = _cabi.HawkLibraryError(f"hawk_result_cfdon: row {bad}", _cabi.HAWK_ECFD)
        err.bad_row = bad
        raise err
    return out[:n]


def featurize_flat(table, params, lead, kmers=True, onehot=False):
    """hawk_result_featurize on the CPU: the kernels' own feature_byte / onehot_channel."""
    from crispr_hawk_b200 import _cabi

    L = lib()
    L.hawkcheck_featurize.restype = C.c_int64
    n = len(table["hap"])
    w = params.guide_len + params.pam_len + 20
    stride = (w + 15) // 16 * 16
    text = np.zeros((max(n, 1), stride), np.uint8)
    if n:
        text[:n, :w] = table["text"][:, :w]
    fl = w - 20 + lead + 3
    k = np.zeros((max(n, 1), fl), np.uint8) if kmers else None
    o = np.zeros((max(n, 1), 4, fl), np.float32) if onehot else None
    bad = L.hawkcheck_featurize(_p(np.ascontiguousarray(table["strand"], np.uint8)), _p(text), C.c_int32(stride), C.c_int32(w),
                                C.c_int32(lead), C.c_int64(n), _p(k) if kmers else None, _p(o) if onehot else None)  # fmt: skip
    if onehot and bad >= 0:
        err = _cabi.HawkLibraryError(f"hawk_result_featurize: row {bad}", _cabi.HAWK_EFEATURE)
        err.bad_row = bad
        raise err
    return (k[:n] if kmers else None), (o[:n] if onehot else None)
