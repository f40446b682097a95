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
