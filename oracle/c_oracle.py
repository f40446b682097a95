"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/scan_oracle.c (the plain-C
restatement of the reference scan used for large parity checks and as the CPU
baseline of bench.py). Never imported by the product package."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libscan_oracle.so")


class _Table(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int64),
        ("hap", C.POINTER(C.c_int32)), ("pos", C.POINTER(C.c_int32)),
        ("start", C.POINTER(C.c_int32)), ("stop", C.POINTER(C.c_int32)),
        ("strand", C.POINTER(C.c_uint8)), ("text", C.POINTER(C.c_uint8)),
        ("window", C.c_int),
        ("n_hits", C.c_int64 * 2), ("hits", C.POINTER(C.c_uint64) * 2),
        ("scanned_bp", C.c_int64), ("status", C.c_int),
    ]  # fmt: skip


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "scan_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(LIB)
        _lib.oracle_search.restype = C.POINTER(_Table)
        _lib.oracle_search2.restype = C.POINTER(_Table)
        _lib.oracle_table_free.argtypes = [C.POINTER(_Table)]
        _lib.oracle_encode.restype = C.c_int64
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    """Host threads the CPU arm may use: the cores this process is allowed to run on.
    (torchrun exports OMP_NUM_THREADS=1 to every rank, so OpenMP's own default is not it.)"""
    import os

    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or int(lib().oracle_max_threads()))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def encode(text: bytes) -> np.ndarray:
    buf = np.frombuffer(text, np.uint8)
    out = np.empty(len(buf), np.uint8)
    bad = lib().oracle_encode(_p(buf), C.c_int64(len(buf)), _p(out))
    if bad >= 0:
        raise ValueError(f"non-IUPAC character at {bad}")
    return out


_scratch = np.empty(0, np.uint8)


def encode_into(buf: np.ndarray) -> None:
    """oracle_encode of one haplotype text into a reusable scratch buffer (timing only)."""
    global _scratch
    if len(_scratch) < len(buf):
        _scratch = np.empty(len(buf), np.uint8)
    bad = lib().oracle_encode(_p(buf), C.c_int64(len(buf)), _p(_scratch))
    if bad >= 0:
        raise ValueError(f"non-IUPAC character at {bad}")


class OracleKeyError(KeyError):
    """The reference's bare KeyError: ambiguity code without variant_alleles entry (:207-213)."""


def search(ascii_slots, slot_off, lens, scan_start, scan_stop, is_ref, seg, pam_fwd, pam_rc, G, right,
           threads=1, raw_only=False, unphased=False, alleles=None):  # fmt: skip
    """Returns dict(hap,strand,pos,start,stop,text) in FINAL order + raw hit lists.
    `unphased` (variants_present and not phased, search_guides.py:473-479) needs `alleles`, the
    flattened variant_alleles tables (crispr_hawk_b200.marshal.AlleleTable)."""
    f = np.asarray(pam_fwd, np.uint8)
    r = np.asarray(pam_rc, np.uint8)
    a = np.ascontiguousarray(scan_start, np.int32)
    b = np.ascontiguousarray(scan_stop, np.int32)
    ir = np.ascontiguousarray(is_ref, np.uint8)
    lens = np.ascontiguousarray(lens, np.int32)
    slot_off = np.ascontiguousarray(slot_off, np.int64)
    if unphased:
        va = [np.ascontiguousarray(x, dt) for x, dt in ((alleles.va_off, np.int64), (alleles.va_idx, np.int32),
                                                        (alleles.va_ent_off, np.int64), (alleles.va_ref, np.uint8))]  # fmt: skip
        vap = [_p(x) for x in va]
    else:
        vap = [None] * 4
    tp = lib().oracle_search2(
        _p(ascii_slots), _p(slot_off), _p(lens), _p(a), _p(b), _p(ir), C.c_int32(len(lens)),
        _p(seg.seg_off), _p(seg.seg_rel), _p(seg.seg_gen), _p(seg.seg_step), _p(f), _p(r),
        C.c_int(len(f)), C.c_int(G), C.c_int(1 if right else 0), C.c_int(threads),
        C.c_int(1 if raw_only else 0), C.c_int(1 if unphased else 0), *vap,
    )  # fmt: skip
    if not tp:
        raise ValueError("oracle_search2: window too long")
    t = tp.contents
    if t.status:
        status = int(t.status)
        lib().oracle_table_free(tp)
        if status == 1:
            raise OracleKeyError("ambiguity code without variant_alleles entry")
        raise ValueError({2: "duplicate REF guide", 3: "expansion above 2^24 strings"}.get(status, f"status {status}"))
    n, w = t.n_rows, t.window

    def arr(ptr, count, dtype):
        if count == 0:
            return np.empty(0, dtype)
        return np.ctypeslib.as_array(ptr, shape=(count,)).astype(dtype, copy=True)

    out = {
        "hap": arr(t.hap, n, np.int32), "strand": arr(t.strand, n, np.uint8),
        "pos": arr(t.pos, n, np.int32), "start": arr(t.start, n, np.int32),
        "stop": arr(t.stop, n, np.int32), "text": arr(t.text, n * w, np.uint8).reshape(n, w),
        "hits": [arr(t.hits[0], t.n_hits[0], np.uint64), arr(t.hits[1], t.n_hits[1], np.uint64)],
        "scanned_bp": int(t.scanned_bp), "window": w,
    }  # fmt: skip
    lib().oracle_table_free(tp)
    return out
