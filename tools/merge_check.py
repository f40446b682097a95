"""Multi-GPU parity + timing of the final merge (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/merge_check.py [--scale 0.05 --haplotypes 40] [--full]

Rank r searches REF + block r of the cohort on its own GPU; the tables are gathered to rank 0
over NCCL with the rows resident in device memory (crispr_hawk_b200.shard.merge_tables_device)
and the first-seen bucket ids recomputed there. Check: the merged table equals ONE search over
the union of all blocks on rank 0's GPU -- every column, bucket ids included. --full times the
gather at the bench workload's full size (no union check: it would not fit the time budget)."""
import argparse
import dataclasses
import os

_OUT = os.dup(1)
os.dup2(2, 1)  # NCCL prints its banner on stdout: keep this tool's stdout to the one JSON line

import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crispr_hawk_b200 import _cabi, shard, synth  # noqa: E402
from crispr_hawk_b200.workload import Workload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--scale", type=float, default=0.05)
ap.add_argument("--haplotypes", type=int, default=40)
ap.add_argument("--full", action="store_true")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
k = synth.CONFIGS[args.workload]
scale = 1.0 if args.full else args.scale
n_alt = k["n_alt_hap"] if args.full else args.haplotypes
cohort = synth.config_cohort(args.workload, scale, n_alt_hap=n_alt, hap_block=rank)
ctx = _cabi.Context.default(local)
wl = Workload(cohort, k["pam"], k["guidelen"], k["right"], ctx, local)
res = wl.step_resident()
key_min, key_span = cohort.region_start, cohort.region_stop - cohort.region_start + 1


session = shard.MergeSession(ctx, rank, world, dev)


def merge():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    m = shard.merge_tables_device(res, ctx, rank * n_alt, rank, world, dev, key_min, key_span, session=session)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    return m, 1e3 * (time.perf_counter() - t0)


merged, _ = merge()  # warm-up (NCCL channels, allocator)
merged, ms = merge()
out = {"world": world, "rows_rank0_local": int(res.n_guides), "merge_ms": ms}
if rank == 0:
    n = int(merged["hap"].shape[0])
    row_bytes = 4 + 1 + 4 + 4 + 4 + res.text_stride
    out.update(rows=n, received_bytes=int((n - res.n_guides) * row_bytes),
               gather_gbs=(n - res.n_guides) * row_bytes / (ms / 1e3) / 1e9 if ms else None)
    if not args.full:
        # ONE search over the union of all blocks
        blocks = [synth.config_cohort(args.workload, scale, n_alt_hap=n_alt, hap_block=r) for r in range(world)]
        hap_sites = np.concatenate([b.hap_sites for b in blocks])
        counts = np.concatenate([[0]] + [np.diff(b.hap_off)[1:] for b in blocks])
        hap_off = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)
        union = dataclasses.replace(blocks[0], hap_off=hap_off, hap_sites=hap_sites.astype(np.int32), _derived={})
        wu = Workload(union, k["pam"], k["guidelen"], k["right"], ctx, local)
        ru = wu.step_resident()
        want = ru.table()
        ru.close()
        for col in ("hap", "strand", "pos", "start", "stop", "bucket"):
            got = merged[col].cpu().numpy()
            assert np.array_equal(got, want[col]), f"column {col} differs"
        assert np.array_equal(merged["text"].cpu().numpy()[:, : want["text"].shape[1]], want["text"])
        out["union_check"] = f"merged table == single search over {union.n_hap} haplotypes ({n} rows, all columns)"
    os.write(_OUT, (json.dumps(out) + "\n").encode())
res.close()
if world > 1:
    dist.destroy_process_group()
