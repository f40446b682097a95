"""world_size-2 (gloo, CPU) test of the multi-GPU sharding and final merge
(crispr_hawk_b200/shard.py). The per-rank guide tables are produced by the C oracle here
(the GPU search needs a device); the merge does not care where a table came from. Rank 0's
merged table must equal the single-process search over the whole cohort, order included."""

import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _emission_table(out):
    key = (out["hap"].astype(np.int64) << 33) | (out["strand"].astype(np.int64) << 32) | out["pos"].astype(np.int64)
    o = np.argsort(key, kind="stable")
    return {k: out[k][o] for k in ("hap", "strand", "pos", "start", "stop", "text")}


def _search(c, d, texts, idx, pam, G, right):
    from crispr_hawk_b200 import marshal, synth
    from crispr_hawk_b200.pam import pam_patterns
    from oracle import c_oracle

    fwd, rc = pam_patterns(pam)
    a, b = synth.scan_bounds(c, len(fwd))
    buf, off, lens = marshal.stage_ascii([texts[i] for i in idx])
    so = d.seg.seg_off
    take = np.concatenate([np.arange(so[h], so[h + 1]) for h in idx])
    seg_off = np.concatenate(([0], np.cumsum(so[idx + 1] - so[idx]))).astype(np.int64)
    seg = marshal.SegmentTable(seg_off, d.seg.seg_rel[take], d.seg.seg_gen[take], d.seg.seg_step[take])
    return c_oracle.search(buf, off, lens, a[idx], b[idx], d.is_ref[idx], seg, fwd, rc, G, right, threads=1)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from crispr_hawk_b200 import shard, synth

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = synth.make_cohort(bed_len=6000, n_alt_hap=11, n_sites=240, mean_alts_per_hap=40, seed=17,
                              snv_frac=0.6, ins_frac=0.2, max_indel=6)  # fmt: skip
        d = synth.derive(c)
        texts = synth.materialize_host(c)
        parts = shard.partition(d.lens, d.is_ref.astype(bool), world)
        mine = parts[rank]
        local = _emission_table(_search(c, d, texts, mine, "NGG", 20, False))
        merged = shard.merge_tables(local, mine, d.is_ref[mine], rank, world)
        if rank == 0:
            full = _search(c, d, texts, np.arange(c.n_hap), "NGG", 20, False)  # final order
            order = np.argsort(merged["bucket"], kind="stable")
            ok = all(np.array_equal(merged[k][order], full[k]) for k in ("hap", "strand", "pos", "start", "stop"))
            ok = ok and np.array_equal(merged["text"][order], full["text"])
            # every non-REF haplotype on exactly one rank, REF on all of them
            cover = np.concatenate([p[~d.is_ref[p].astype(bool)] for p in parts])
            ok = ok and sorted(cover.tolist()) == list(range(1, c.n_hap)) and all(0 in p for p in parts)
            q.put(("ok" if ok else "mismatch", len(full["hap"])))
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


def test_partition_is_balanced_and_ordered():
    from crispr_hawk_b200 import shard

    lens = np.array([1000] + [1000 + (i % 7) for i in range(40)])
    is_ref = np.array([True] + [False] * 40)
    for world in (1, 2, 3, 8):
        parts = shard.partition(lens, is_ref, world)
        assert len(parts) == world and all(p[0] == 0 for p in parts)
        alt = np.concatenate([p[1:] for p in parts])
        assert alt.tolist() == list(range(1, 41))
        sizes = [lens[p[1:]].sum() for p in parts]
        assert max(sizes) - min(sizes) <= 2 * lens.max()


def test_first_seen_buckets():
    from crispr_hawk_b200 import shard

    start = np.array([5, 3, 5, 5, 3, 9], np.int32)
    strand = np.array([0, 0, 1, 0, 0, 1], np.uint8)
    assert shard.first_seen_buckets(start, strand).tolist() == [0, 1, 2, 0, 1, 5]


def test_two_rank_merge_equals_single_process_search():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    status, n = q.get(timeout=5)
    assert status == "ok" and n > 200
