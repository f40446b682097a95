"""ctypes binding of libhawkscan.so (include/hawkscan.h). No CPU fallback: if the
library or a CUDA device is missing every entry point raises `HawkLibraryError`."""

from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HAWKSCAN_LIB", os.path.join(PKG_DIR, "libhawkscan.so"))

ABI_VERSION = 3
HAWK_MAX_PAM = 16
HAWK_F_UNPHASED = 1

HAWK_OK = 0
HAWK_EINVAL = -1
HAWK_ECUDA = -2
HAWK_ENOMEM = -3
HAWK_EIUPAC = -4
HAWK_ECAPACITY = -5
HAWK_EALLELES = -6
HAWK_EDUPREF = -7
HAWK_EASSERT = -8
HAWK_ECFD = -9
HAWK_EFEATURE = -10


class HawkLibraryError(RuntimeError):
    """libhawkscan.so is missing, cannot be loaded, or reported a failure."""

    def __init__(self, message: str, code: int = 0):
        super().__init__(message)
        self.code = code


class HawkParams(C.Structure):
    _fields_ = [
        ("pam_len", C.c_int32),
        ("guide_len", C.c_int32),
        ("right", C.c_int32),
        ("flags", C.c_uint32),
        ("pam_fwd", C.c_uint8 * HAWK_MAX_PAM),
        ("pam_rc", C.c_uint8 * HAWK_MAX_PAM),
    ]


class HawkTableOut(C.Structure):
    """hawk_table_out: caller-owned host columns of the streamed search."""

    _fields_ = [
        ("hap", C.POINTER(C.c_int32)),
        ("strand", C.POINTER(C.c_uint8)),
        ("pos", C.POINTER(C.c_int32)),
        ("start", C.POINTER(C.c_int32)),
        ("stop", C.POINTER(C.c_int32)),
        ("bucket", C.POINTER(C.c_uint32)),
        ("text", C.POINTER(C.c_uint8)),
        ("capacity", C.c_int64),
        ("text_stride", C.c_int32),
    ]


_P = C.c_void_p
_I32P = C.POINTER(C.c_int32)
_I64P = C.POINTER(C.c_int64)
_U8P = C.POINTER(C.c_uint8)
_U32P = C.POINTER(C.c_uint32)
_U64P = C.POINTER(C.c_uint64)

# name -> (restype, argtypes); every symbol include/hawkscan.h declares
SIGNATURES = {
    "hawk_abi_version": (C.c_int, []),
    "hawk_last_error": (C.c_char_p, []),
    "hawk_strerror": (C.c_char_p, [C.c_int]),
    "hawk_launch_count": (C.c_int64, []),
    "hawk_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "hawk_ctx_destroy": (C.c_int, [_P]),
    "hawk_ctx_info": (C.c_int, [_P, _I32P, _I64P, _I64P]),
    "hawk_ctx_traffic": (C.c_int, [_P, _I64P, _I64P]),
    "hawk_layout": (C.c_int, [_I32P, C.c_int32, _I64P, _I64P]),
    "hawk_batch_create": (C.c_int, [_P, _U8P, _I64P, _I32P, C.c_int32, C.POINTER(_P), _I64P]),
    "hawk_batch_create_dev": (C.c_int, [_P, _P, _I64P, _I32P, C.c_int32, C.POINTER(_P), _I64P]),
    "hawk_batch_repack_dev": (C.c_int, [_P, _P, _I64P]),
    "hawk_encode_search_dev": (
        C.c_int, [_P, _P, _P, C.POINTER(HawkParams), _I32P, _I32P, _U8P, C.POINTER(_P), _I64P],
    ),
    "hawk_batch_create_from_edits": (
        C.c_int,
        [_P, _U8P, C.c_int64, C.c_int32, C.c_int32, _I64P, _I32P, _I32P, _I32P, _I64P, _U8P, C.c_int64,
         C.POINTER(_P), _I64P],
    ),  # fmt: skip
    "hawk_batch_layout": (C.c_int, [_P, _I64P, _I32P]),
    "hawk_ctx_stream": (C.c_void_p, [_P]),
    "hawk_ctx_set_profiling": (C.c_int, [_P, C.c_int32]),
    "hawk_ctx_set_fused": (C.c_int, [_P, C.c_int32]),
    "hawk_ctx_set_edit_planes": (C.c_int, [_P, C.c_int32]),
    "hawk_ctx_sync": (C.c_int, [_P]),
    "hawk_ctx_profile": (C.c_int, [_P, C.POINTER(C.c_double), _I64P]),
    "hawk_materialize_dev": (
        C.c_int, [_P, _P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _P],
    ),
    "hawk_batch_destroy": (C.c_int, [_P]),
    "hawk_batch_export_nibbles": (C.c_int, [_P, C.c_int32, _U8P, _U8P]),
    "hawk_batch_set_posmap": (C.c_int, [_P, _I64P, _I32P, _I32P, _U8P]),
    "hawk_batch_set_alleles": (C.c_int, [_P, _I64P, _I32P, _I64P, _U8P]),
    "hawk_batch_set_scan": (C.c_int, [_P, _I32P, _I32P, _U8P]),
    "hawk_search": (C.c_int, [_P, _P, C.POINTER(HawkParams), _I32P, _I32P, _U8P, C.POINTER(_P)]),
    "hawk_pam_search": (C.c_int, [_P, _P, C.POINTER(HawkParams), _I32P, _I32P, C.POINTER(_P)]),
    "hawk_result_destroy": (C.c_int, [_P]),
    "hawk_result_info": (C.c_int, [_P, _I64P, _I64P, _I32P, _I32P, _I64P]),
    "hawk_result_fetch": (C.c_int, [_P, _I32P, _U8P, _I32P, _I32P, _I32P, _U32P, _U8P]),
    "hawk_result_fetch_hits": (C.c_int, [_P, C.c_int32, _U64P]),
    "hawk_result_device_columns": (C.c_int, [_P, C.POINTER(_P)]),
    "hawk_first_seen_dev": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int64, _P, _P]),
    "hawk_merge_layout": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, _I64P, _I64P]),
    "hawk_peer_alloc": (C.c_int, [_P, C.c_int64, C.POINTER(_P), _U8P]),
    "hawk_peer_free": (C.c_int, [_P, _P]),
    "hawk_peer_open": (C.c_int, [_P, _U8P, C.POINTER(_P)]),
    "hawk_peer_close": (C.c_int, [_P, _P]),
    "hawk_result_push": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, C.c_int64, C.c_int64, C.c_int32, _I64P]),
    "hawk_pack_dev": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, _P]),
    "hawk_scan_plan": (C.c_int64, [_I32P, _I32P, C.c_int32, _I64P]),
    "hawk_scan_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int64]),
    "hawk_scan_match_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "hawk_scan_totals": (C.c_void_p, [_P]),
    "hawk_scan_count_dev": (
        C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int64, C.POINTER(HawkParams), C.c_int32, _P],
    ),
    "hawk_scan_match_dev": (
        C.c_int,
        [_P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int64, C.POINTER(HawkParams), C.c_int32,
         C.c_int64, _P, _P, _P],
    ),  # fmt: skip
    "hawk_scan_expand_dev": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int64, _P, _P, _P, _P, _P]),
    "hawk_batch_set_variants": (C.c_int, [_P, _I64P, _I32P, _I32P, _I32P, _I64P, _U8P, C.c_int64]),
    "hawk_result_annotate": (C.c_int, [_P, _P, _U8P, _I32P, _I32P, _I64P, _I64P]),
    "hawk_result_fetch_variants": (C.c_int, [_P, _I32P]),
    "hawk_result_collapse": (C.c_int, [_P, _U8P, C.c_int32, _U32P, _U8P, _I32P]),
    "hawk_result_cfdon": (C.c_int, [_P, _U8P, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), _I64P]),
    "hawk_result_featurize": (C.c_int, [_P, C.c_int32, _U8P, _P, C.c_int32, _I64P]),
    "hawk_table_text_stride": (C.c_int32, [C.c_int32, C.c_int32]),
    "hawk_stream_plan": (C.c_int32, [_I64P, C.c_int32, _U8P, C.c_int32, _I32P, _I32P, C.c_int32]),
    "hawk_search_stream": (
        C.c_int,
        [_P, _U8P, _I64P, _I32P, C.c_int32, _I64P, _I32P, _I32P, _U8P, C.POINTER(HawkParams), _I32P, _I32P, _U8P,
         C.c_int32, C.POINTER(HawkTableOut), _I64P, _I64P, _I64P, _I64P],
    ),  # fmt: skip
    "hawk_search_stream_edits": (
        C.c_int,
        [_P, _U8P, C.c_int64, C.c_int32, C.c_int32, _I64P, _I32P, _I32P, _I32P, _I64P, _U8P, C.c_int64,
         C.POINTER(HawkParams), _I32P, _I32P, _U8P, C.c_int32, C.POINTER(HawkTableOut), _I64P, _I64P, _I64P],
    ),  # fmt: skip
}

_lib = None
_lock = threading.Lock()


def load_library(path: Optional[str] = None):
    """dlopen libhawkscan.so and bind every declared symbol (works without a GPU)."""
    global _lib
    with _lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise HawkLibraryError(
                f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(crispr_hawk_b200 has no CPU fallback)"
            )
        try:
            lib = C.CDLL(p)
        except OSError as e:
            raise HawkLibraryError(f"cannot load {p}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise HawkLibraryError(f"{p} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        if lib.hawk_abi_version() != ABI_VERSION:
            raise HawkLibraryError(f"{p}: unexpected ABI version {lib.hawk_abi_version()}")
        if path is None:
            _lib = lib
        return lib


def check(rc: int, what: str = "") -> None:
    if rc != HAWK_OK:
        lib = load_library()
        msg = lib.hawk_last_error().decode("utf-8", "replace")
        raise HawkLibraryError(f"{what or 'libhawkscan'}: {msg} [{lib.hawk_strerror(rc).decode()}]", rc)


def ptr(arr: Optional[np.ndarray], ctype):
    if arr is None:
        return None
    assert arr.flags["C_CONTIGUOUS"], "array must be contiguous"
    return arr.ctypes.data_as(C.POINTER(ctype))


def make_params(pam_fwd, pam_rc, guide_len: int, right: bool, unphased: bool) -> HawkParams:
    p = HawkParams()
    if len(pam_fwd) != len(pam_rc) or not (1 <= len(pam_fwd) <= HAWK_MAX_PAM):
        raise HawkLibraryError(f"PAM length must be 1..{HAWK_MAX_PAM} (got {len(pam_fwd)})", HAWK_EINVAL)
    p.pam_len = len(pam_fwd)
    p.guide_len = int(guide_len)
    p.right = 1 if right else 0
    p.flags = HAWK_F_UNPHASED if unphased else 0
    for i, (a, b) in enumerate(zip(pam_fwd, pam_rc)):
        p.pam_fwd[i] = int(a)
        p.pam_rc[i] = int(b)
    return p


class Context:
    """One device, one stream (hawk_ctx)."""

    _default = {}

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = _P()
        check(self.lib.hawk_ctx_create(int(device), C.byref(h)), "hawk_ctx_create")
        self.handle = h
        self.device = device

    @classmethod
    def default(cls, device: Optional[int] = None) -> "Context":
        if device is None:
            device = int(os.environ.get("HAWKSCAN_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def info(self):
        sm, tot, free = C.c_int32(), C.c_int64(), C.c_int64()
        check(self.lib.hawk_ctx_info(self.handle, C.byref(sm), C.byref(tot), C.byref(free)))
        return {"sm_count": sm.value, "total_mem": tot.value, "free_mem": free.value}

    def traffic(self):
        """(host-to-device, device-to-host) bytes moved by the host layer so far."""
        a, b = C.c_int64(), C.c_int64()
        check(self.lib.hawk_ctx_traffic(self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def stream(self) -> int:
        """cudaStream_t of the context as an integer (for torch.cuda.ExternalStream)."""
        return int(self.lib.hawk_ctx_stream(self.handle) or 0)

    def set_fused(self, mode: int) -> None:
        """hawk_ctx_set_fused: 0 staged, 1 fused kernel, 2 by haplotype shape (default)."""
        check(self.lib.hawk_ctx_set_fused(self.handle, int(mode)))

    def set_edit_planes(self, mode: int) -> None:
        """hawk_ctx_set_edit_planes: 1 planes around the edits only (default), 0 texts + K1."""
        check(self.lib.hawk_ctx_set_edit_planes(self.handle, int(mode)))

    def set_profiling(self, enabled: bool) -> None:
        check(self.lib.hawk_ctx_set_profiling(self.handle, 1 if enabled else 0))

    def profile(self):
        """{'pack'|'scan'|'post': (total ms, launches)} since the last call."""
        ms, n = (C.c_double * 5)(), (C.c_int64 * 5)()
        check(self.lib.hawk_ctx_profile(self.handle, ms, n))
        return {k: (ms[i], n[i]) for i, k in enumerate(("pack", "cand", "post", "expand", "match"))}

    def close(self):
        if self.handle:
            self.lib.hawk_ctx_destroy(self.handle)
            self.handle = None


class Batch:
    """Packed haplotypes of one region on the device (hawk_batch)."""

    def __init__(self, ctx: Context, ascii_slots, slot_off: np.ndarray, lens: np.ndarray,
                 device_ptr: Optional[int] = None):  # fmt: skip
        """`ascii_slots`: host uint8 array in the slot layout, or None with `device_ptr`
        = address of the same bytes in device memory (hawk_batch_create_dev)."""
        self.ctx, self.lib = ctx, ctx.lib
        self.slot_off = np.ascontiguousarray(slot_off, dtype=np.int64)
        self.lens = np.ascontiguousarray(lens, dtype=np.int32)
        self.n_hap = len(self.lens)
        h, bad = _P(), C.c_int64(-1)
        if device_ptr is not None:
            rc = self.lib.hawk_batch_create_dev(
                ctx.handle, C.c_void_p(device_ptr), ptr(self.slot_off, C.c_int64),
                ptr(self.lens, C.c_int32), self.n_hap, C.byref(h), C.byref(bad),
            )  # fmt: skip
        else:
            ascii_slots = np.ascontiguousarray(ascii_slots, dtype=np.uint8)
            rc = self.lib.hawk_batch_create(
                ctx.handle, ptr(ascii_slots, C.c_uint8), ptr(self.slot_off, C.c_int64),
                ptr(self.lens, C.c_int32), self.n_hap, C.byref(h), C.byref(bad),
            )  # fmt: skip
        self.bad_slot = bad.value
        if rc != HAWK_OK:
            err = HawkLibraryError(self.lib.hawk_last_error().decode(), rc)
            err.bad_slot = bad.value
            raise err
        self.handle = h
        self.has_posmap = False
        self.has_alleles = False
        self.has_variants = False

    @classmethod
    def from_edits(cls, ctx: "Context", ref_ascii: np.ndarray, region_start: int, edit_off, edit_pos, edit_reflen,
                   edit_altlen, edit_altoff, alt_pool) -> "Batch":  # fmt: skip
        """hawk_batch_create_from_edits: haplotypes materialised on the device from edit lists."""
        self = cls.__new__(cls)
        self.ctx, self.lib = ctx, ctx.lib
        ref = np.ascontiguousarray(ref_ascii, dtype=np.uint8)
        eo = np.ascontiguousarray(edit_off, dtype=np.int64)
        ep = np.ascontiguousarray(edit_pos, dtype=np.int32)
        rl = np.ascontiguousarray(edit_reflen, dtype=np.int32)
        al = np.ascontiguousarray(edit_altlen, dtype=np.int32)
        ao = np.ascontiguousarray(edit_altoff, dtype=np.int64)
        pool = np.ascontiguousarray(alt_pool, dtype=np.uint8)
        self.n_hap = len(eo) - 1
        h, bad = _P(), C.c_int64(-1)
        rc = self.lib.hawk_batch_create_from_edits(
            ctx.handle, ptr(ref, C.c_uint8), len(ref), int(region_start), self.n_hap, ptr(eo, C.c_int64),
            ptr(ep, C.c_int32), ptr(rl, C.c_int32), ptr(al, C.c_int32), ptr(ao, C.c_int64), ptr(pool, C.c_uint8),
            len(pool), C.byref(h), C.byref(bad),
        )  # fmt: skip
        self.bad_slot = bad.value
        if rc != HAWK_OK:
            err = HawkLibraryError(self.lib.hawk_last_error().decode(), rc)
            err.bad_slot = bad.value
            raise err
        self.handle = h
        self.slot_off = np.zeros(self.n_hap + 1, np.int64)
        self.lens = np.zeros(self.n_hap, np.int32)
        check(self.lib.hawk_batch_layout(h, ptr(self.slot_off, C.c_int64), ptr(self.lens, C.c_int32)))
        self.has_posmap, self.has_alleles = True, False
        self.has_variants = True  # the edit lists stay on the device as the variant table (N2)
        return self

    def export_text(self, hap: int) -> str:
        """Haplotype text rebuilt from the planes and the case plane."""
        nib, low = self.export_nibbles(hap, want_lower=True)
        lut = np.frombuffer(b"?ACMGRSVTWYHKDBN", np.uint8)
        return (lut[nib] | (low << 5).astype(np.uint8)).tobytes().decode("ascii")

    def repack(self, device_ptr: int) -> None:
        """Re-run K1 from device-resident texts of the same layout (hawk_batch_repack_dev)."""
        bad = C.c_int64(-1)
        check(self.lib.hawk_batch_repack_dev(self.handle, C.c_void_p(device_ptr), C.byref(bad)),
              "hawk_batch_repack_dev")  # fmt: skip

    def export_nibbles(self, hap: int, want_lower: bool = False):
        n = int(self.lens[hap])
        nib = np.empty(n, np.uint8)
        low = np.empty(n, np.uint8) if want_lower else None
        check(self.lib.hawk_batch_export_nibbles(self.handle, hap, ptr(nib, C.c_uint8), ptr(low, C.c_uint8)))
        return (nib, low) if want_lower else nib

    def set_posmap(self, seg):
        check(
            self.lib.hawk_batch_set_posmap(
                self.handle, ptr(seg.seg_off, C.c_int64), ptr(seg.seg_rel, C.c_int32),
                ptr(seg.seg_gen, C.c_int32), ptr(seg.seg_step, C.c_uint8),
            ),  # fmt: skip
            "hawk_batch_set_posmap",
        )
        self.has_posmap = True

    def set_alleles(self, va):
        check(
            self.lib.hawk_batch_set_alleles(
                self.handle, ptr(va.va_off, C.c_int64), ptr(va.va_idx, C.c_int32),
                ptr(va.va_ent_off, C.c_int64), ptr(va.va_ref, C.c_uint8),
            ),  # fmt: skip
            "hawk_batch_set_alleles",
        )
        self.has_alleles = True

    def set_scan(self, scan_start, scan_stop, is_ref):
        """hawk_batch_set_scan: attach the scan bounds; searches may then pass None for them."""
        a = np.ascontiguousarray(scan_start, dtype=np.int32)
        b = np.ascontiguousarray(scan_stop, dtype=np.int32)
        r = np.ascontiguousarray(is_ref, dtype=np.uint8)
        check(self.lib.hawk_batch_set_scan(self.handle, ptr(a, C.c_int32), ptr(b, C.c_int32), ptr(r, C.c_uint8)),
              "hawk_batch_set_scan")  # fmt: skip
        self.has_scan = True

    def set_variants(self, vt):
        """N2: the haplotypes' variant tables (marshal.VariantTable) for Result.annotate."""
        check(
            self.lib.hawk_batch_set_variants(
                self.handle, ptr(vt.var_off, C.c_int64), ptr(vt.var_pos, C.c_int32), ptr(vt.var_reflen, C.c_int32),
                ptr(vt.var_altlen, C.c_int32), ptr(vt.var_altoff, C.c_int64), ptr(vt.alt_pool, C.c_uint8),
                len(vt.alt_pool),
            ),  # fmt: skip
            "hawk_batch_set_variants",
        )
        self.has_variants = True

    def device_bytes(self) -> int:
        """Device memory of the packed planes (0.625 B per slot + the chunk slack)."""
        total = int(self.slot_off[-1]) if len(self.slot_off) else 0
        return total * 5 // 8 + 1024

    def close(self):
        if getattr(self, "handle", None):
            self.lib.hawk_batch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Result:
    """Guide table / hit lists of one search (hawk_result)."""

    def __init__(self, lib, handle):
        self.lib, self.handle = lib, handle
        n, hits, w, ts, bp = C.c_int64(), (C.c_int64 * 2)(), C.c_int32(), C.c_int32(), C.c_int64()
        check(lib.hawk_result_info(handle, C.byref(n), hits, C.byref(w), C.byref(ts), C.byref(bp)))
        self.n_guides, self.n_hits, self.window, self.scanned_bp = n.value, (hits[0], hits[1]), w.value, bp.value
        self.text_stride = ts.value or w.value

    def device_bytes(self) -> int:
        """Device memory the table's columns and hit lists hold (rows x 21 B + text + 8 B per hit)."""
        rows = max(int(self.n_guides), int(self.n_hits[0] + self.n_hits[1]))  # columns are sized for every hit kept
        return rows * (21 + self.text_stride + 8)

    def table(self, buffers=None, want_text: bool = True):
        """Download the guide table. `buffers`: optional dict of preallocated (e.g. pinned)
        1-D numpy arrays with at least n rows each; views of them are returned.
        `want_text=False`: the window-text column stays on the device (no "text" key)."""
        n, w, ts = self.n_guides, self.window, self.text_stride
        if not want_text:
            out = ({k: buffers[k][:n] for k in ("hap", "strand", "pos", "start", "stop", "bucket")} if buffers is not None else
                   {"hap": np.empty(n, np.int32), "strand": np.empty(n, np.uint8), "pos": np.empty(n, np.int32),
                    "start": np.empty(n, np.int32), "stop": np.empty(n, np.int32), "bucket": np.empty(n, np.uint32)})  # fmt: skip
            check(
                self.lib.hawk_result_fetch(
                    self.handle, ptr(out["hap"], C.c_int32), ptr(out["strand"], C.c_uint8), ptr(out["pos"], C.c_int32),
                    ptr(out["start"], C.c_int32), ptr(out["stop"], C.c_int32), ptr(out["bucket"], C.c_uint32), None,
                ),  # fmt: skip
                "hawk_result_fetch",
            )
            return out
        if buffers is not None:
            out = {k: buffers[k][:n] for k in ("hap", "strand", "pos", "start", "stop", "bucket")}
            out["text"] = buffers["text"][: n * ts].reshape(n, ts)
        else:
            out = {
                "hap": np.empty(n, np.int32), "strand": np.empty(n, np.uint8), "pos": np.empty(n, np.int32),
                "start": np.empty(n, np.int32), "stop": np.empty(n, np.int32),
                "bucket": np.empty(n, np.uint32), "text": np.empty((n, ts), np.uint8),
            }  # fmt: skip
        check(
            self.lib.hawk_result_fetch(
                self.handle, ptr(out["hap"], C.c_int32), ptr(out["strand"], C.c_uint8),
                ptr(out["pos"], C.c_int32), ptr(out["start"], C.c_int32), ptr(out["stop"], C.c_int32),
                ptr(out["bucket"], C.c_uint32), ptr(out["text"], C.c_uint8),
            ),  # fmt: skip
            "hawk_result_fetch",
        )
        out["text"] = out["text"][:, :w]  # rows are padded to text_stride bytes on the device
        return out

    def annotate(self, batch: "Batch", want_variants: bool = True, want_text: bool = True, buffers=None):
        """hawk_result_annotate (N2): reverse-complemented text of the strand-1 rows, GC counts
        of the guides, and per-row CSR lists of the haplotype-local variant indices that
        polish_guide_variants keeps. Row order = table order. `buffers`: optional dict of
        preallocated (e.g. pinned) arrays `rc_text` (n * text_stride), `gc_num`, `gc_den` (n),
        `gv_off` (n + 1), `gv_idx` (used when large enough for the variant references)."""
        n, ts, w = self.n_guides, self.text_stride, self.window
        b = buffers or {}
        rc = num = den = off = None
        if want_text:
            rc = b["rc_text"][: n * ts] if "rc_text" in b else np.zeros(n * ts, np.uint8)
            num = b["gc_num"][:n] if "gc_num" in b else np.zeros(n, np.int32)
            den = b["gc_den"][:n] if "gc_den" in b else np.zeros(n, np.int32)
        if want_variants:
            off = b["gv_off"][: n + 1] if "gv_off" in b else np.zeros(n + 1, np.int64)
        total = C.c_int64(0)
        check(
            self.lib.hawk_result_annotate(self.handle, batch.handle, ptr(rc, C.c_uint8), ptr(num, C.c_int32),
                                          ptr(den, C.c_int32), ptr(off, C.c_int64), C.byref(total)),
            "hawk_result_annotate",
        )  # fmt: skip
        if "gv_idx" in b and len(b["gv_idx"]) >= total.value:
            idx = b["gv_idx"][: total.value]
        else:
            idx = np.zeros(total.value, np.int32)
        if total.value:
            check(self.lib.hawk_result_fetch_variants(self.handle, ptr(idx, C.c_int32)), "hawk_result_fetch_variants")
        return {"rc_text": rc.reshape(n, ts)[:, :w] if want_text else None, "gc_num": num, "gc_den": den,
                "gv_off": off, "gv_idx": idx}  # fmt: skip

    def collapse(self, is_ref, buffers=None):
        """hawk_result_collapse: (perm, head, collision) -- rows ordered by (start, stop, group),
        head[k] = 1 where a report row (group) starts. `buffers`: optional dict of preallocated
        (e.g. pinned) arrays `perm` (uint32) and `head` (uint8) with at least n rows."""
        r = np.ascontiguousarray(is_ref, dtype=np.uint8)
        n, col = self.n_guides, C.c_int32(0)
        perm = buffers["perm"][:n] if buffers else np.empty(n, np.uint32)
        head = buffers["head"][:n] if buffers else np.empty(n, np.uint8)
        check(self.lib.hawk_result_collapse(self.handle, ptr(r, C.c_uint8), len(r), ptr(perm, C.c_uint32), ptr(head, C.c_uint8),
                                            C.byref(col)), "hawk_result_collapse")  # fmt: skip
        return perm, head, bool(col.value)

    def cfdon(self, is_ref, mm, pam2, out=None):
        """hawk_result_cfdon: float64 CFDon score per row (NaN: no REF guide at the row's key).
        `out`: optional preallocated (e.g. pinned) float64 array with at least n rows."""
        r = np.ascontiguousarray(is_ref, dtype=np.uint8)
        mm = np.ascontiguousarray(mm, dtype=np.float64).reshape(320)
        p2 = np.ascontiguousarray(pam2, dtype=np.float64).reshape(16)
        out, bad = (out[: self.n_guides] if out is not None else np.empty(self.n_guides, np.float64)), C.c_int64(-1)
        rc = self.lib.hawk_result_cfdon(self.handle, ptr(r, C.c_uint8), len(r), ptr(mm, C.c_double), ptr(p2, C.c_double),
                                        ptr(out, C.c_double), C.byref(bad))  # fmt: skip
        if rc != HAWK_OK:
            err = HawkLibraryError(self.lib.hawk_last_error().decode(), rc)
            err.bad_row = bad.value
            raise err
        return out

    def featurize(self, lead: int = 4, kmers: bool = True, onehot: bool = False, onehot_device_ptr: int = 0,
                  kmers_out=None):
        """hawk_result_featurize: the learned scorers' inputs of every row (emission order).
        Returns (kmers, onehot): `kmers` an (n, L) uint8 array of upper-case letters
        (scoring.py:50-84; lead 4, or 0 for sgDesigner) or None; `onehot` the float32 (n, 4, L)
        tensor of DeepCpf1's preprocess (seqdeepcpf1.py:71-92) as a numpy array, or None when not
        requested / when `onehot_device_ptr` names device memory of n * 4 * L floats to fill."""
        n, L = self.n_guides, self.window - 20 + lead + 3
        k = None
        if kmers:
            k = kmers_out[: n * L] if kmers_out is not None else np.empty(n * L, np.uint8)
        o = np.empty((n, 4, L), np.float32) if onehot and not onehot_device_ptr else None
        optr = C.c_void_p(onehot_device_ptr) if onehot_device_ptr else (o.ctypes.data_as(_P) if o is not None else None)
        bad = C.c_int64(-1)
        rc = self.lib.hawk_result_featurize(self.handle, lead, ptr(k, C.c_uint8) if k is not None else None, optr,
                                            1 if onehot_device_ptr else 0, C.byref(bad))  # fmt: skip
        if rc != HAWK_OK:
            err = HawkLibraryError(self.lib.hawk_last_error().decode(), rc)
            err.bad_row = bad.value
            raise err
        return (k.reshape(n, L) if k is not None else None), o

    def device_columns(self):
        """Borrowed device addresses {column: int} of the table (hawk_result_device_columns)."""
        cols = (_P * 7)()
        check(self.lib.hawk_result_device_columns(self.handle, cols), "hawk_result_device_columns")
        return {k: int(cols[i] or 0) for i, k in enumerate(("hap", "strand", "pos", "start", "stop", "bucket", "text"))}

    def hits(self, strand: int) -> np.ndarray:
        out = np.empty(self.n_hits[strand], np.uint64)
        check(self.lib.hawk_result_fetch_hits(self.handle, strand, ptr(out, C.c_uint64)))
        return out

    def close(self):
        if getattr(self, "handle", None):
            self.lib.hawk_result_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def search(ctx: Context, batch: Batch, params: HawkParams, scan_start=None, scan_stop=None, is_ref=None) -> Result:
    a = None if scan_start is None else np.ascontiguousarray(scan_start, dtype=np.int32)
    b = None if scan_stop is None else np.ascontiguousarray(scan_stop, dtype=np.int32)
    r = None if is_ref is None else np.ascontiguousarray(is_ref, dtype=np.uint8)
    h = _P()
    check(
        ctx.lib.hawk_search(ctx.handle, batch.handle, C.byref(params), ptr(a, C.c_int32),
                            ptr(b, C.c_int32), ptr(r, C.c_uint8), C.byref(h)),
        "hawk_search",
    )  # fmt: skip
    return Result(ctx.lib, h)


def encode_search(ctx: Context, batch: Batch, device_ptr: int, params: HawkParams, scan_start=None, scan_stop=None,
                  is_ref=None) -> Result:  # fmt: skip
    """hawk_encode_search_dev: re-encode `batch` from device-resident texts and search it in one
    pass. Bounds None: the ones attached with Batch.set_scan."""
    a = None if scan_start is None else np.ascontiguousarray(scan_start, dtype=np.int32)
    b = None if scan_stop is None else np.ascontiguousarray(scan_stop, dtype=np.int32)
    r = None if is_ref is None else np.ascontiguousarray(is_ref, dtype=np.uint8)
    h, bad = _P(), C.c_int64(-1)
    rc = ctx.lib.hawk_encode_search_dev(ctx.handle, batch.handle, C.c_void_p(device_ptr), C.byref(params),
                                        ptr(a, C.c_int32), ptr(b, C.c_int32), ptr(r, C.c_uint8), C.byref(h), C.byref(bad))  # fmt: skip
    if rc != HAWK_OK:
        err = HawkLibraryError(ctx.lib.hawk_last_error().decode(), rc)
        err.bad_slot = bad.value
        raise err
    return Result(ctx.lib, h)


def pam_search(ctx: Context, batch: Batch, params: HawkParams, scan_start, scan_stop) -> Result:
    a = np.ascontiguousarray(scan_start, dtype=np.int32)
    b = np.ascontiguousarray(scan_stop, dtype=np.int32)
    h = _P()
    check(
        ctx.lib.hawk_pam_search(ctx.handle, batch.handle, C.byref(params), ptr(a, C.c_int32),
                                ptr(b, C.c_int32), C.byref(h)),
        "hawk_pam_search",
    )  # fmt: skip
    return Result(ctx.lib, h)


# --------------------------------------------------------------------------- streamed search
TABLE_COLUMNS = (("hap", np.int32), ("strand", np.uint8), ("pos", np.int32), ("start", np.int32),
                 ("stop", np.int32), ("bucket", np.uint32))  # fmt: skip


def text_stride(params: HawkParams) -> int:
    return int(load_library().hawk_table_text_stride(params.pam_len, params.guide_len))


def alloc_table(capacity: int, stride: int, pinned: bool = False, want_text: bool = True):
    """Host columns for `capacity` guide rows (pinned through torch when asked: the streamed
    search overlaps its copies only with page-locked memory). `want_text=False`: no text column
    (hawk_table_out.text = NULL: the rows come down at 21 bytes each instead of 21 + stride)."""
    if pinned:
        import torch

        mk = lambda dt, k=1: torch.empty(max(capacity * k, 1), dtype=dt, pin_memory=True).numpy()  # noqa: E731
        tdt = {np.int32: torch.int32, np.uint8: torch.uint8, np.int64: torch.int64, np.uint32: torch.int32}
        out = {name: mk(tdt[dt]).view(dt) for name, dt in TABLE_COLUMNS}
        out["text"] = mk(torch.uint8, stride) if want_text else None
    else:
        out = {name: np.empty(max(capacity, 1), dt) for name, dt in TABLE_COLUMNS}
        out["text"] = np.empty(max(capacity * stride, 1), np.uint8) if want_text else None
    return out


def _table_out(buffers, stride: int) -> HawkTableOut:
    t = HawkTableOut()
    cap = min(len(buffers[name]) for name, _ in TABLE_COLUMNS)
    text = buffers.get("text")
    if text is not None:
        cap = min(cap, len(text.reshape(-1)) // stride)
    t.hap, t.strand, t.pos = ptr(buffers["hap"], C.c_int32), ptr(buffers["strand"], C.c_uint8), ptr(buffers["pos"], C.c_int32)
    t.start, t.stop = ptr(buffers["start"], C.c_int32), ptr(buffers["stop"], C.c_int32)
    t.bucket, t.text = ptr(buffers["bucket"], C.c_uint32), (ptr(text.reshape(-1), C.c_uint8) if text is not None else None)
    t.capacity, t.text_stride = cap, stride
    return t


def _table_views(buffers, n: int, stride: int, window: int):
    out = {name: buffers[name][:n] for name, _ in TABLE_COLUMNS}
    if buffers.get("text") is not None:
        out["text"] = buffers["text"].reshape(-1)[: n * stride].reshape(n, stride)[:, :window]
    return out


class StreamResult:
    """Guide table of a streamed search, in host memory."""

    def __init__(self, table, n_guides, n_hits, scanned_bp, window, stride, buffers):
        self.table_, self.n_guides, self.n_hits, self.scanned_bp = table, n_guides, n_hits, scanned_bp
        self.window, self.text_stride, self.buffers = window, stride, buffers

    def table(self):
        return self.table_


def _run_stream(call, params: HawkParams, buffers, pinned: bool, what: str, want_text: bool = True) -> StreamResult:
    """Run `call(table_out_ptr, n, hits, bp)`; on HAWK_ECAPACITY grow the buffers and repeat."""
    stride = text_stride(params)
    window = params.pam_len + params.guide_len + 20
    if buffers is None:
        # count first: one pass without output sizes the table exactly
        n, hits, bp = C.c_int64(), (C.c_int64 * 2)(), C.c_int64()
        check(call(None, n, hits, bp), what)
        buffers = alloc_table(n.value, stride, pinned, want_text)
    for _ in range(2):
        t = _table_out(buffers, stride)
        n, hits, bp = C.c_int64(), (C.c_int64 * 2)(), C.c_int64()
        rc = call(C.byref(t), n, hits, bp)
        if rc == HAWK_ECAPACITY and n.value > t.capacity:
            buffers = alloc_table(int(n.value * 1.05) + 1024, stride, pinned, want_text)
            continue
        check(rc, what)
        return StreamResult(_table_views(buffers, n.value, stride, window), n.value, (hits[0], hits[1]), bp.value,
                            window, stride, buffers)  # fmt: skip
    raise HawkLibraryError(f"{what}: output capacity still too small after growing", HAWK_ECAPACITY)


def search_stream(ctx: Context, ascii_slots: np.ndarray, slot_off, lens, seg, params: HawkParams, scan_start,
                  scan_stop, is_ref, n_groups: int = 0, buffers=None, pinned: bool = False) -> StreamResult:  # fmt: skip
    """hawk_search_stream: host texts in, host guide table out, PCIe copies overlapped.
    `buffers` (see alloc_table) are reused when large enough."""
    asc = np.ascontiguousarray(ascii_slots, dtype=np.uint8)
    so = np.ascontiguousarray(slot_off, dtype=np.int64)
    ln = np.ascontiguousarray(lens, dtype=np.int32)
    a = np.ascontiguousarray(scan_start, dtype=np.int32)
    b = np.ascontiguousarray(scan_stop, dtype=np.int32)
    r = np.ascontiguousarray(is_ref, dtype=np.uint8)
    bad = C.c_int64(-1)

    def call(t, n, hits, bp):
        return ctx.lib.hawk_search_stream(
            ctx.handle, ptr(asc, C.c_uint8), ptr(so, C.c_int64), ptr(ln, C.c_int32), len(ln),
            ptr(seg.seg_off, C.c_int64), ptr(seg.seg_rel, C.c_int32), ptr(seg.seg_gen, C.c_int32),
            ptr(seg.seg_step, C.c_uint8), C.byref(params), ptr(a, C.c_int32), ptr(b, C.c_int32), ptr(r, C.c_uint8),
            int(n_groups), t, C.byref(n), hits, C.byref(bp), C.byref(bad),
        )  # fmt: skip

    try:
        return _run_stream(call, params, buffers, pinned, "hawk_search_stream")
    except HawkLibraryError as e:
        e.bad_slot = bad.value
        raise


def search_stream_edits(ctx: Context, ref_ascii, region_start: int, edit_off, edit_pos, edit_reflen, edit_altlen,
                        edit_altoff, alt_pool, params: HawkParams, scan_start, scan_stop, is_ref,
                        n_groups: int = 0, buffers=None, pinned: bool = False, want_text: bool = True) -> StreamResult:  # fmt: skip
    """hawk_search_stream_edits: reference text + edit lists in, host guide table out."""
    ref = np.ascontiguousarray(ref_ascii, dtype=np.uint8)
    eo = np.ascontiguousarray(edit_off, dtype=np.int64)
    ep = np.ascontiguousarray(edit_pos, dtype=np.int32)
    rl = np.ascontiguousarray(edit_reflen, dtype=np.int32)
    al = np.ascontiguousarray(edit_altlen, dtype=np.int32)
    ao = np.ascontiguousarray(edit_altoff, dtype=np.int64)
    pool = np.ascontiguousarray(alt_pool, dtype=np.uint8)
    a = np.ascontiguousarray(scan_start, dtype=np.int32)
    b = np.ascontiguousarray(scan_stop, dtype=np.int32)
    r = np.ascontiguousarray(is_ref, dtype=np.uint8)

    def call(t, n, hits, bp):
        return ctx.lib.hawk_search_stream_edits(
            ctx.handle, ptr(ref, C.c_uint8), len(ref), int(region_start), len(eo) - 1, ptr(eo, C.c_int64),
            ptr(ep, C.c_int32), ptr(rl, C.c_int32), ptr(al, C.c_int32), ptr(ao, C.c_int64), ptr(pool, C.c_uint8),
            len(pool), C.byref(params), ptr(a, C.c_int32), ptr(b, C.c_int32), ptr(r, C.c_uint8), int(n_groups), t,
            C.byref(n), hits, C.byref(bp),
        )  # fmt: skip

    return _run_stream(call, params, buffers, pinned, "hawk_search_stream_edits", want_text)


def stream_plan(slot_off, is_ref, n_groups: int = 0):
    """hawk_stream_plan: [(lo, hi), ...] haplotype groups of the streamed search (host only)."""
    so = np.ascontiguousarray(slot_off, dtype=np.int64)
    r = np.ascontiguousarray(is_ref, dtype=np.uint8)
    lo, hi = np.zeros(256, np.int32), np.zeros(256, np.int32)
    n = load_library().hawk_stream_plan(ptr(so, C.c_int64), len(r), ptr(r, C.c_uint8), int(n_groups), ptr(lo, C.c_int32),
                                        ptr(hi, C.c_int32), 256)  # fmt: skip
    if n < 0:
        check(n, "hawk_stream_plan")
    return [(int(lo[g]), int(hi[g])) for g in range(n)]
