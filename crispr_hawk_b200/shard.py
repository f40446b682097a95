"""Multi-GPU sharding of one region's search (one process per GPU, torch.distributed).

Every (region, haplotype) pair is an independent unit of the scan; only the redundancy
filter (search_guides.py:340-369) needs the REF guides of the same region, and REF is one
haplotype, so every rank scans REF + a contiguous block of the other haplotypes and no
collective is needed on the data path. The per-rank guide tables are gathered once, at the
end, for the final merge: rank blocks are contiguous haplotype ranges, so the reference's
emission order (haplotype-major, search_guides.py:530-547) is the concatenation of the rank
tables in rank order, and the first-seen bucket ids (:306-337) are recomputed over the
merged table.

Works with the `nccl` backend (tables stay on the GPUs, NVLink send/recv) and with `gloo`
(CPU tensors; used by the tests)."""

from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

COLUMNS = ("hap", "strand", "pos", "start", "stop")


def partition(lengths: Sequence[int], is_ref: Sequence[bool], world: int) -> List[np.ndarray]:
    """Haplotype indices per rank: the REF haplotype(s) on every rank, the others dealt as
    contiguous blocks of roughly equal total length (order preserved)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    is_ref = np.asarray(is_ref, dtype=bool)
    ref = np.flatnonzero(is_ref)
    alt = np.flatnonzero(~is_ref)
    csum = np.cumsum(lengths[alt]) if len(alt) else np.zeros(0, np.int64)
    total = int(csum[-1]) if len(alt) else 0
    out = []
    for r in range(world):
        lo = int(np.searchsorted(csum, total * r / world, side="right")) if r else 0
        hi = int(np.searchsorted(csum, total * (r + 1) / world, side="right")) if r + 1 < world else len(alt)
        out.append(np.sort(np.concatenate((ref, alt[lo:hi]))))
    return out


def _pack_rows(table: Dict[str, np.ndarray], keep: np.ndarray, global_hap: np.ndarray):
    """Rows `keep` of a local table as one contiguous uint8 matrix (fixed bytes per row)."""
    n = int(keep.sum())
    w = table["text"].shape[1] if table["text"].ndim == 2 else 0
    row = 4 + 1 + 4 + 4 + 4 + w
    buf = np.zeros((n, row), np.uint8)
    buf[:, 0:4] = global_hap[table["hap"][keep]].astype("<i4").view(np.uint8).reshape(n, 4)
    buf[:, 4] = table["strand"][keep]
    for k, name in enumerate(("pos", "start", "stop")):
        buf[:, 5 + 4 * k : 9 + 4 * k] = table[name][keep].astype("<i4").view(np.uint8).reshape(n, 4)
    if w:
        buf[:, 17:] = table["text"][keep]
    return buf, w


def _unpack_rows(buf: np.ndarray, w: int) -> Dict[str, np.ndarray]:
    n = len(buf)
    out = {"hap": np.ascontiguousarray(buf[:, 0:4]).view("<i4").reshape(n),
           "strand": np.ascontiguousarray(buf[:, 4])}  # fmt: skip
    for k, name in enumerate(("pos", "start", "stop")):
        out[name] = np.ascontiguousarray(buf[:, 5 + 4 * k : 9 + 4 * k]).view("<i4").reshape(n)
    out["text"] = np.ascontiguousarray(buf[:, 17 : 17 + w])
    return out


def first_seen_buckets(start: np.ndarray, strand: np.ndarray) -> np.ndarray:
    """bucket[i] = smallest row index sharing row i's (start, strand) key
    (group_guides_position, search_guides.py:306-337: dict insertion order)."""
    key = start.astype(np.int64) * 2 + strand
    order = np.argsort(key, kind="stable")
    sk = key[order]
    head = np.ones(len(sk), bool)
    head[1:] = sk[1:] != sk[:-1]
    first = order[head]  # stable sort: first element of every run is the smallest index
    bucket = np.empty(len(key), np.uint32)
    bucket[order] = first[np.cumsum(head) - 1]
    return bucket


def merge_tables(table: Dict[str, np.ndarray], local_haps: np.ndarray, is_ref_local: np.ndarray,
                 rank: int, world: int, group=None, device: Optional[str] = None):  # fmt: skip
    """Final merge. `table`: this rank's guide table in emission order (hawk_result_fetch
    columns); `local_haps[i]` = global index of local haplotype i. Ranks > 0 drop their REF
    rows (rank 0 owns them). Returns the merged table on rank 0 (with `bucket`), None
    elsewhere. `device`: where the exchanged tensors live ("cuda:N" for nccl, None for gloo)."""
    import torch
    import torch.distributed as dist

    keep = np.ones(len(table["hap"]), bool)
    if rank > 0:
        keep &= ~np.asarray(is_ref_local, bool)[table["hap"]]
    buf, w = _pack_rows(table, keep, np.asarray(local_haps, np.int64))
    dev = torch.device(device) if device else torch.device("cpu")
    if world == 1:
        merged = _unpack_rows(buf, w)
    else:
        meta = torch.tensor([buf.shape[0], buf.shape[1]], dtype=torch.int64, device=dev)
        metas = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(metas, meta, group=group)
        t = torch.from_numpy(buf).to(dev)
        if rank == 0:
            parts = [buf]
            for r in range(1, world):
                n, row = int(metas[r][0]), int(metas[r][1])
                recv = torch.empty((n, row), dtype=torch.uint8, device=dev)
                if n:
                    dist.recv(recv, src=r, group=group)
                parts.append(recv.cpu().numpy())
            merged = _unpack_rows(np.concatenate([p.reshape(-1, buf.shape[1]) for p in parts]), w)
        else:
            if t.shape[0]:
                dist.send(t, dst=0, group=group)
            return None
    merged["bucket"] = first_seen_buckets(merged["start"], merged["strand"])
    return merged


# --------------------------------------------------------------------------- device-resident merge
class _DevArray:
    """A borrowed device buffer as a `__cuda_array_interface__` object (zero-copy into torch)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"data": (ptr, False), "shape": tuple(shape), "typestr": typestr, "version": 2}


_COLS_T = (("hap", "int32"), ("strand", "uint8"), ("pos", "int32"), ("start", "int32"), ("stop", "int32"))
_TYPESTR = {"int32": "<i4", "uint8": "|u1", "uint32": "<u4"}


def result_tensors(res, device: str):
    """The columns of a `_cabi.Result` as torch tensors over the library's own device memory."""
    import torch

    n, ts = res.n_guides, res.text_stride
    ptrs = res.device_columns()
    out = {}
    for name, dt in _COLS_T:
        out[name] = (torch.as_tensor(_DevArray(ptrs[name], (n,), _TYPESTR[dt]), device=device) if n
                     else torch.empty(0, dtype=getattr(torch, dt), device=device))  # fmt: skip
    out["text"] = (torch.as_tensor(_DevArray(ptrs["text"], (n, ts), "|u1"), device=device) if n
                   else torch.empty((0, ts), dtype=torch.uint8, device=device))  # fmt: skip
    return out


class MergeSession:
    """The gathering rank's buffer for merged tables, allocated once and mapped by every other
    rank through CUDA IPC (hawk_peer_alloc / hawk_peer_open). Mapping a peer's memory costs
    ~10 ms and freeing exported memory ~20 ms, so the buffer is kept across merges and only
    re-made (a collective) when a merge needs more room. A merged table is a view of this
    buffer: valid until the session's next merge or close()."""

    def __init__(self, ctx, rank: int, world: int, device: str, group=None):
        self.ctx, self.rank, self.world, self.device, self.group = ctx, rank, world, device, group
        self.capacity, self.base = 0, None

    def ensure(self, nbytes: int):
        """Collective: afterwards every rank holds a pointer to >= nbytes of rank 0's buffer."""
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _cabi

        if nbytes <= self.capacity:
            return self.base
        self.close()
        lib = self.ctx.lib
        want = int(nbytes * 1.25) + (1 << 20)
        base = C.c_void_p()
        handle = torch.zeros(64, dtype=torch.uint8)
        if self.rank == 0:
            hbuf = (C.c_uint8 * 64)()
            _cabi.check(lib.hawk_peer_alloc(self.ctx.handle, want, C.byref(base), hbuf), "hawk_peer_alloc")
            handle = torch.tensor(list(hbuf), dtype=torch.uint8)
        if self.world > 1:
            hdev = handle.to(self.device)
            dist.broadcast(hdev, src=0, group=self.group)
            handle = hdev.cpu()
        if self.rank > 0:
            hbuf = (C.c_uint8 * 64)(*handle.tolist())
            _cabi.check(lib.hawk_peer_open(self.ctx.handle, hbuf, C.byref(base)), "hawk_peer_open")
        self.base, self.capacity = base, want
        return base

    def close(self):
        """Collective when the session holds a buffer (peers unmap before the owner frees)."""
        if self.base is None:
            return
        import torch.distributed as dist

        lib = self.ctx.lib
        if self.rank > 0:
            lib.hawk_peer_close(self.ctx.handle, self.base)
        if self.world > 1:
            dist.barrier(group=self.group)
        if self.rank == 0:
            lib.hawk_peer_free(self.ctx.handle, self.base)
        self.base, self.capacity = None, 0


class MergedTable(dict):
    """The merged guide table on the gathering rank: {column: torch tensor} over the session's
    device buffer (valid until the session's next merge / close)."""

    def __init__(self, session, cols):
        super().__init__(cols)
        self.session = session

    def close(self):
        """Drop the views; a table that made its own (temporary) session also frees the buffer."""
        self.clear()
        if self.session is not None and self.session.world == 1:
            self.session.close()
        self.session = None


_MERGE_COLS = (("hap", "<i4", 4), ("strand", "|u1", 1), ("pos", "<i4", 4), ("start", "<i4", 4), ("stop", "<i4", 4),
               ("bucket", "<u4", 4))  # fmt: skip


def merge_tables_device(res, ctx, hap_offset: int, rank: int, world: int, device: str, key_min: int, key_span: int,
                        group=None, with_text: bool = True, session: Optional[MergeSession] = None):  # fmt: skip
    """Final merge with the tables resident on the GPUs, as a one-sided push over NVLink: rank 0
    owns ONE buffer for the merged table, mapped by every other rank (MergeSession: CUDA IPC,
    set up once and reused); every rank writes its rows (ranks > 0 without the REF rows, which
    every rank emits first and rank 0 owns; haplotype indices shifted by `hap_offset`) straight
    into its slice -- all ranks at once, no receive calls (hawk_result_push). Rank blocks are
    contiguous haplotype ranges, so the concatenation in rank order is the reference's emission
    order; rank 0 then recomputes the first-seen bucket ids in place (hawk_first_seen_dev).
    `with_text=False` leaves the window text behind (the host that owns the haplotype texts can
    slice it, as search_guides.py:134-160 does). Returns a MergedTable on rank 0, None elsewhere
    (`.session` keeps the buffer; pass it to the next merge, close() it at the end -- both
    collectives). Every rank returns only after its own rows have landed (stream synchronised)
    and a barrier, so `res` may be closed right away."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from . import _cabi

    lib = ctx.lib
    torch.cuda.synchronize(device)  # the table was written on the library's stream
    t = result_tensors(res, device)
    n, ts = res.n_guides, res.text_stride
    n_ref = int((t["hap"] == 0).sum().item()) if (rank > 0 and n) else 0  # REF rows: a prefix
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    mine = torch.tensor([n - n_ref], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_gather(counts, mine, group=group)
    else:
        counts = [mine]
    sizes = [int(c.item()) for c in counts]
    total = sum(sizes)
    at = sum(sizes[:rank])
    off = (C.c_int64 * 7)()
    nbytes = C.c_int64()
    _cabi.check(lib.hawk_merge_layout(total, ts, 1 if with_text else 0, off, C.byref(nbytes)), "hawk_merge_layout")
    temporary = session is None
    session = session or MergeSession(ctx, rank, world, device, group)
    base = session.ensure(nbytes.value)
    import time

    t_push = time.perf_counter()
    pushed = C.c_int64()
    _cabi.check(lib.hawk_result_push(res.handle, n_ref, int(hap_offset) if rank > 0 else 0, 0, base, total, at,
                                     1 if with_text else 0, C.byref(pushed)), "hawk_result_push")  # fmt: skip
    torch.cuda.synchronize(device)  # hawk_result_push runs on the library's stream
    _cabi.check(lib.hawk_ctx_sync(ctx.handle), "hawk_ctx_sync")
    if world > 1:
        dist.barrier(group=group)  # every rank's rows are in rank 0's memory
    push_ms = 1e3 * (time.perf_counter() - t_push)
    if temporary and world > 1:
        # no session to keep: the peers unmap now, rank 0's table owns the buffer from here on
        if rank > 0:
            lib.hawk_peer_close(ctx.handle, base)
            session.base, session.capacity = None, 0
        dist.barrier(group=group)
        session.world = 1  # closing it later is no longer a collective
    if rank > 0:
        return None
    cols = {}
    for k, (name, typestr, width) in enumerate(_MERGE_COLS):
        cols[name] = (torch.as_tensor(_DevArray(base.value + off[k], (total,), typestr), device=device) if total
                      else torch.empty(0, dtype=torch.int32, device=device))  # fmt: skip
    if with_text:
        cols["text"] = (torch.as_tensor(_DevArray(base.value + off[6], (total, ts), "|u1"), device=device) if total
                        else torch.empty((0, ts), dtype=torch.uint8, device=device))  # fmt: skip
    merged = MergedTable(session, cols)
    merged.received_bytes = (total - sizes[0]) * (17 + (ts if with_text else 0))
    merged.push_ms = push_ms  # hawk_result_push on every rank + the barrier, as rank 0 saw it
    merged.first_seen_ms = 0.0
    t_fs = time.perf_counter()
    if total:
        if getattr(session, "_key_table", None) is None or session._key_table.numel() < 2 * int(key_span) + 1:
            session._key_table = torch.empty(2 * int(key_span) + 1, dtype=torch.int32, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        _cabi.check(
            lib.hawk_first_seen_dev(C.c_void_p(stream), C.c_void_p(base.value + off[3]), C.c_void_p(base.value + off[1]),
                                    total, int(key_min), int(key_span), C.c_void_p(session._key_table.data_ptr()),
                                    C.c_void_p(base.value + off[5])),
            "hawk_first_seen_dev",
        )  # fmt: skip
        merged.first_seen_ms = 1e3 * (time.perf_counter() - t_fs)  # hawk_first_seen_dev synchronises its stream
    return merged
