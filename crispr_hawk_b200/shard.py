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
    bucket = np.empty(len(key), np.int64)
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
_TYPESTR = {"int32": "<i4", "uint8": "|u1"}


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


def merge_tables_device(res, ctx, hap_offset: int, rank: int, world: int, device: str, key_min: int, key_span: int,
                        group=None):  # fmt: skip
    """Final merge with the tables resident on the GPUs: every rank's guide table stays in the
    library's device memory, ranks > 0 send their rows (without the REF rows, which every rank
    emits first and rank 0 owns) to rank 0 over NCCL (NVLink), column by column; rank 0
    concatenates in rank order -- the reference's emission order, rank blocks being contiguous
    haplotype ranges -- and recomputes the first-seen bucket ids on the device
    (hawk_first_seen_dev). `hap_offset`: added to the local index of this rank's non-REF
    haplotypes (local 0 = REF stays 0). Returns {column: torch tensor} on rank 0, None elsewhere.
    The caller keeps `res` alive until the merge is done."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    torch.cuda.synchronize(device)  # the table was written on the library's stream
    t = result_tensors(res, device)
    n, ts = res.n_guides, res.text_stride
    n_ref = int((t["hap"] == 0).sum().item()) if (rank > 0 and n) else 0  # REF rows: a prefix
    lo = n_ref
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    mine = torch.tensor([n - lo], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_gather(counts, mine, group=group)
    else:
        counts = [mine]
    sizes = [int(c.item()) for c in counts]
    if rank > 0:
        if n - lo > 0:
            hap = t["hap"][lo:] + int(hap_offset)
            # plain blocking sends, one column after the other: every transfer gets the whole
            # NVLink path to rank 0 (a batched isend/irecv group of the same columns measured
            # 4x slower: 23 ms vs 5.3 ms for 0.95 GB between two B200s)
            for name, _ in _COLS_T:
                dist.send((hap if name == "hap" else t[name][lo:]).contiguous(), dst=0, group=group)
            dist.send(t["text"][lo:].contiguous(), dst=0, group=group)
        return None
    total = sum(sizes)
    merged = {name: torch.empty(total, dtype=getattr(torch, dt), device=device) for name, dt in _COLS_T}
    merged["text"] = torch.empty((total, ts), dtype=torch.uint8, device=device)
    at = 0
    for r in range(world):
        m = sizes[r]
        if r == 0:
            for name, _ in _COLS_T:
                merged[name][:m] = t[name]  # rank 0: local indices are global already (hap_offset 0)
            merged["text"][:m] = t["text"]
        elif m:
            for name, _ in _COLS_T:
                dist.recv(merged[name][at : at + m], src=r, group=group)
            dist.recv(merged["text"][at : at + m], src=r, group=group)
        at += m
    merged["bucket"] = torch.empty(total, dtype=torch.int64, device=device)
    if total:
        table = torch.empty(2 * int(key_span), dtype=torch.int32, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        from . import _cabi

        _cabi.check(
            ctx.lib.hawk_first_seen_dev(C.c_void_p(stream), C.c_void_p(merged["start"].data_ptr()),
                                        C.c_void_p(merged["strand"].data_ptr()), total, int(key_min), int(key_span),
                                        C.c_void_p(table.data_ptr()), C.c_void_p(merged["bucket"].data_ptr())),
            "hawk_first_seen_dev",
        )  # fmt: skip
        torch.cuda.synchronize(device)
    return merged
