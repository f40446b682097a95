#!/usr/bin/env python
"""bench.py -- haplotype-bp scanned per second for the guide-discovery hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2]

One *step* = one pass of the hot path (encode -> PAM scan on both strands -> in-range /
REF-core filters -> genomic coordinates -> redundancy removal -> guide table) over one
batch of synthetic haplotypes of the BASELINE.json shape (default: config 2, SpCas9
NGG / 20 nt, 1 Mb region, 2,504 phased samples -> 5,008 haplotypes + REF).

* `value`  : whole-job haplotype-bp/s, haplotype texts already resident in HBM.
* `e2e`    : the same through host buffers (pinned ASCII in, guide table out, PCIe copies
             inside the timed region).
* `roofline`: the dominant kernel's algorithmic bytes / its CUDA-event duration against the
             measured HBM copy bandwidth (MEASURED_PEAKS.json).
* `cpu_baseline`: the C restatement of the reference scan (oracle/, a checker -- never on
             the product path) on a bounded sample of the same workload, on the host cores.

`--impl reference` times only that CPU arm (all host threads). Under torchrun every rank
scans its own block of haplotypes of the same region (weak scaling, no data-path
collective; one NCCL all-reduce of the whole job's guide / hit counts at the end of the timed
region).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Rank 0 prints exactly ONE line on stdout, the JSON result. Libraries that write to the
# process's stdout behind Python's back (NCCL prints its version banner there) are kept out of
# it: file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to a
# duplicate of the original stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


import numpy as np  # noqa: E402

METRIC = "haplotype_bp_scanned_per_s"
UNIT = "hap-bp/s"
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampling of one GPU during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )  # fmt: skip
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        """Summary of the samples taken inside [t_begin, t_end] (the timed region); when the
        region was too short to catch one, of all samples since start() (warm-up + timed: the
        same kernels back to back) -- `window` says which."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if t_begin is not None and t_begin <= t <= t_end]
        window = "timed region" if inside else "warm-up + timed region"
        for r in inside or [r for _, r in self.rows]:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "samples": len(sm),
            "window": window,
            "reasons": sorted(reasons),
        }


# --------------------------------------------------------------------------- CPU arm
def cpu_sample(workload: str, n_alt: int, scale: float):
    """A bounded sample of the workload for the CPU arm: the same region and sites, the
    first `n_alt` non-reference haplotypes (+ REF), materialised on the host."""
    from crispr_hawk_b200 import marshal, synth

    k = synth.CONFIGS[workload]
    if k.get("unphased"):
        return cpu_sample_unphased(workload, scale)
    c = synth.config_cohort(workload, scale, n_alt_hap=n_alt)
    d = synth.derive(c)
    texts = synth.materialize_host(c)
    buf, off, lens = marshal.stage_ascii(texts)
    from crispr_hawk_b200.pam import pam_patterns

    fwd, rc = pam_patterns(k["pam"])
    a, b = synth.scan_bounds(c, len(fwd))
    bp = int((b.astype(np.int64) - a).clip(min=0).sum())
    return dict(cohort=c, d=d, texts=texts, buf=buf, off=off, lens=lens, fwd=fwd, rc=rc, a=a, b=b, bp=bp,
                G=k["guidelen"], right=k["right"], pam=k["pam"])  # fmt: skip


def cpu_sample_unphased(workload: str, scale: float):
    """Config 4's CPU sample: the same generator on a region `CPU_C4_FRACTION` of the size (all
    of its haplotypes: REF, the IUPAC SNV haplotypes, every indel-window haplotype)."""
    from crispr_hawk_b200 import synth, synth_unphased
    from crispr_hawk_b200.pam import pam_patterns

    k = synth.CONFIGS[workload]
    c = synth.config_cohort(workload, scale * CPU_C4_FRACTION)
    u = synth_unphased.derive_unphased(c)
    fwd, rc = pam_patterns(k["pam"])
    a, b = u.scan_bounds(c, len(fwd))
    bp = int((b.astype(np.int64) - a).clip(min=0).sum())
    return dict(cohort=c, u=u, buf=u.ascii, off=u.slot_off, lens=u.lens, fwd=fwd, rc=rc, a=a, b=b, bp=bp,
                G=k["guidelen"], right=k["right"], pam=k["pam"], unphased=True, n_hap=u.n_hap)  # fmt: skip


CPU_C4_FRACTION = 0.1


def cpu_step(s, threads):
    from oracle import c_oracle

    if s.get("unphased"):
        # encode happens inside the search (per haplotype, in the OpenMP loop)
        t0 = time.perf_counter()
        u = s["u"]
        out = c_oracle.search(s["buf"], s["off"], s["lens"], s["a"], s["b"], u.is_ref, u.seg, s["fwd"], s["rc"], s["G"],
                              s["right"], threads=threads, unphased=True, alleles=u.alleles)  # fmt: skip
        return time.perf_counter() - t0, len(out["hap"]), 0
    t0 = time.perf_counter()
    enc_bad = 0
    # encoder.encode over every haplotype (scalar table lookup), then search
    for h in range(len(s["lens"])):
        o, n = int(s["off"][h]), int(s["lens"][h])
        c_oracle.encode_into(s["buf"][o : o + n])
    out = c_oracle.search(s["buf"], s["off"], s["lens"], s["a"], s["b"], s["d"].is_ref, s["d"].seg,
                          s["fwd"], s["rc"], s["G"], s["right"], threads=threads)  # fmt: skip
    dt = time.perf_counter() - t0
    return dt, len(out["hap"]), enc_bad


def python_port_rate(workload: str, scale: float, budget_s: float = 6.0):
    """The pure-Python restatement (same interpreter-bound cost model as the reference,
    which is pure Python) on a tiny slice -- context for the C port's number."""
    from crispr_hawk_b200 import synth
    from oracle import hawk_oracle as O

    k = synth.CONFIGS[workload]
    if k.get("unphased"):
        from crispr_hawk_b200 import synth_unphased as SU

        c = SU.make_unphased_cohort(3000, 77, k["pitch"], k["snv"], k["multi"], k["max_indel"], k["n_samples"])
        haps = SU.unphased_haplotypes(c)
        a, b = SU.derive_unphased(c).scan_bounds(c, len(k["pam"]))
    else:
        c = synth.make_cohort(20_000, 3, 200, 20, seed=77, snv_frac=k["snv"], ins_frac=k["ins"], max_indel=k["max_indel"])
        haps = synth.synth_haplotypes(c)
        a, b = synth.scan_bounds(c, len(k["pam"]))
    bp = int((b.astype(np.int64) - a).sum())
    t0 = time.perf_counter()
    ohaps = [O.OracleHap.from_object(h) for h in haps]
    n = len(O.search(k["pam"], c.region_start, c.region_stop, ohaps, k["guidelen"], k["right"], True, not k.get("unphased")))
    dt = time.perf_counter() - t0
    return bp / dt, bp, n


def run_reference_arm(args, rank, world):
    if rank != 0:
        return 0
    from oracle import c_oracle

    threads = c_oracle.max_threads()
    # size the sample so one step is a few seconds of CPU work on all threads
    n_alt = max(threads * 2, 16) - 1  # REF + n_alt = a multiple of the thread count: no idle tail in the OpenMP loop
    s = cpu_sample(args.workload, n_alt, args.scale)
    if s.get("unphased"):
        n_alt = s["n_hap"] - 1
    times = []
    for i in range(args.warmup + args.steps):
        dt, n, _ = cpu_step(s, threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = s["bp"] / (ms / 1e3)
    sample = f"{n_alt + 1} haplotypes (REF + first {n_alt}) of workload {args.workload} x {s['bp'] // (n_alt + 1)} bp, encode + search, per step"
    if s.get("unphased"):
        sample = (f"workload {args.workload} on a region {CPU_C4_FRACTION:g} of the size: all {s['n_hap']} haplotypes, "
                  f"{s['bp']:,} hap-bp, encode + unphased search, per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, s["bp"], n_alt + 1, sample=True),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU arm: oracle/scan_oracle.c (C restatement of encoder.py + search_guides.py, OpenMP over haplotypes). "
                "The reference itself is single-threaded pure Python and cannot travel to this box.",
    }  # fmt: skip
    emit(line)
    return 0


def workload_config(args, scanned_bp, n_hap, sample=False):
    from crispr_hawk_b200 import synth

    k = synth.CONFIGS[args.workload]
    bed = int(k['bed_len'] * args.scale * (CPU_C4_FRACTION if sample and k.get("unphased") else 1.0))
    kind = ("unphased (10 population pseudo-samples, 1 site / %d bp): REF + IUPAC SNV haplotypes + indel-window haplotypes,"
            % k["pitch"]) if k.get("unphased") else "phased,"
    return {
        "workload": f"{args.workload}: {k['pam']} / {k['guidelen']} nt / {'right' if k['right'] else 'left'}, "
                    f"{bed:,} bp region, {kind} {n_hap} haplotypes per rank"
                    + (" (CPU sample)" if sample else ""),
        "haplotypes_per_rank": n_hap, "scanned_bp_per_rank_per_step": scanned_bp,
        "l2": "n/a (CPU arm)" if sample else ("inputs exceed L2 (no flush needed)" if scanned_bp > 400e6 else "L2 flushed between steps"),
        "seed": k["seed"],
    }  # fmt: skip



# --------------------------------------------------------------------------- Python seam legs
def api_search_leg():
    """`crispr_hawk_b200.search()` -- Python haplotype objects in, the reference's List[Guide] out
    -- beside the pure-Python port of the reference's search (oracle/hawk_oracle.py, the
    reference's own cost model) on the same objects; guides compared field by field."""
    import crispr_hawk_b200 as hawk
    from crispr_hawk_b200 import synth
    from oracle import hawk_oracle as O

    out = {}
    for label, name, scale, n_alt in (("config1_5kb_x21", "c1", 1.0, 20), ("c2_100kb_x65", "c2", 0.1, 64)):
        k = synth.CONFIGS[name]
        c = synth.config_cohort(name, scale, n_alt_hap=n_alt)
        haps = synth.synth_haplotypes(c)
        region = synth.SynthRegion(c)
        a, b = synth.scan_bounds(c, len(k["pam"]))
        pam = hawk.PAM(k["pam"], k["right"], True)
        pam.encode(0)

        def run():
            return hawk.search(pam, region, haps, hawk.encode_region(haps, 0, True), k["guidelen"], k["right"], True, True, 0, True)

        run()
        t0 = time.perf_counter()
        guides = run()
        t_list = time.perf_counter() - t0
        guides[0]
        t_first = time.perf_counter() - t0
        guides.realise()
        t_all = time.perf_counter() - t0
        t0 = time.perf_counter()
        want = O.search(k["pam"], c.region_start, c.region_stop, [O.OracleHap.from_object(h) for h in haps], k["guidelen"],
                        k["right"], True, True)  # fmt: skip
        t_py = time.perf_counter() - t0
        same = [(g.start, g.stop, g.strand, g.sequence, g.hapid) for g in guides] == [(g.start, g.stop, g.strand, g.sequence, g.hapid) for g in want]
        if not same:
            raise AssertionError(f"api_search {label}: guides differ from the Python port")
        out[label] = {"hap_bp": int((b.astype(np.int64) - a).clip(min=0).sum()), "guides": len(guides),
                      "search_returns_ms": 1e3 * t_list, "first_guide_ms": 1e3 * t_first, "all_guides_built_ms": 1e3 * t_all,
                      "python_port_ms": 1e3 * t_py, "speedup_all_guides_built": t_py / t_all}  # fmt: skip
    out["what"] = ("encode_region + search through the public Python API (host texts up, table without the text column "
                   "down, lazy GuideList); all_guides_built forces every Guide object like a report writer would")
    return out


def variant_records_leg(n_samples: int = 200):
    """N1 through the seam install() binds (crispr_hawk_b200.haplotypes.add_variants_phased): the
    reference's VariantRecord lists of config 2's region for `n_samples` phased samples ->
    haplotypes built on the device from edit lists -> search()."""
    import types

    import crispr_hawk_b200 as hawk
    from crispr_hawk_b200 import haplotypes as HN
    from crispr_hawk_b200 import synth

    k = synth.CONFIGS["c2"]
    c = synth.config_cohort("c2", 1.0, n_alt_hap=2 * n_samples)
    ref = c.ref.tobytes().decode()
    pool = c.alt_pool.tobytes().decode()
    names = [f"S{i + 1}" for i in range(n_samples)]
    carriers = [(set(), set()) for _ in range(len(c.site_pos))]
    for h in range(1, c.n_hap):
        smp, copy = names[(h - 1) // 2], (h - 1) % 2
        for site in c.hap_sites[c.hap_off[h] : c.hap_off[h + 1]].tolist():
            carriers[site][copy].add(smp)
    records = []
    for site in range(len(c.site_pos)):
        if not (carriers[site][0] or carriers[site][1]):
            continue
        p, rl, al, ao = int(c.site_pos[site]), int(c.site_reflen[site]), int(c.site_altlen[site]), int(c.site_altoff[site])
        r, a = ref[p : p + rl], pool[ao : ao + al]
        records.append(types.SimpleNamespace(position=c.region_start + p, ref=r, alt=[a], afs=[0.1], samples=[carriers[site]],
                                             vtype=["snp" if rl == 1 and al == 1 else "indel"],
                                             id=[f"chr1-{c.region_start + p}-{r}/{a}"]))  # fmt: skip
    region = synth.SynthRegion(c)
    region.sequence = types.SimpleNamespace(sequence=ref)
    vcfs = {"chr1": types.SimpleNamespace(samples=names)}
    pam = hawk.PAM(k["pam"], k["right"], True)
    pam.encode(0)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        haps = HN.add_variants_phased([types.SimpleNamespace(afs={})], region, vcfs, records, True, True)
        t_build = time.perf_counter() - t0
        guides = hawk.search(pam, region, haps, hawk.encode_region(haps, 0, True), k["guidelen"], k["right"], True, True, 0, True)
        t_all = time.perf_counter() - t0
        guides[len(guides) // 2]
        if best is None or t_all < best[1]:
            best = (t_build, t_all, len(haps), len(guides))
    bp = sum(len(h) - 200 - len(k["pam"]) for h in haps)
    return {"samples": n_samples, "records": len(records), "haplotypes": best[2], "guides": best[3], "hap_bp": int(bp),
            "haplotypes_built_ms": 1e3 * best[0], "build_plus_search_ms": 1e3 * best[1],
            "value": bp / best[1], "unit": UNIT,
            "what": "VariantRecord lists -> edit lists (host metadata) -> hawk_batch_create_from_edits -> search(); the "
                    "reference's own add_variants_phased rewrites the sequence and two 1,000,200-entry dicts per variant "
                    "and haplotype for the same input"}  # fmt: skip


# --------------------------------------------------------------------------- product arm
def run_product_arm(args, rank, world, local_rank):
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        try:  # run (and first-touch the pinned staging buffers) on the CPUs next to this rank's GPU
            import pynvml

            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        except Exception:
            pass
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from crispr_hawk_b200 import _cabi, synth
    from crispr_hawk_b200.workload import UnphasedWorkload, Workload

    lib = _cabi.load_library()
    k = synth.CONFIGS[args.workload]
    unphased = bool(k.get("unphased"))
    ctx = _cabi.Context.default(local_rank)
    if unphased:
        # every rank searches a region of its own (same shape, its own seed): regions are the
        # reference's unit of work (crisprhawk.py:84-115 loops over them)
        cohort = synth.config_cohort(args.workload, args.scale, seed_offset=1000 * rank)
        wl = UnphasedWorkload(cohort, k["pam"], k["guidelen"], k["right"], ctx, local_rank)
        n_hap_rank = wl.d.n_hap
    else:
        n_alt = args.haplotypes or k["n_alt_hap"]
        cohort = synth.config_cohort(args.workload, args.scale, n_alt_hap=n_alt, hap_block=rank)
        wl = Workload(cohort, k["pam"], k["guidelen"], k["right"], ctx, local_rank)
        n_hap_rank = cohort.n_hap
    wl.prepare_resident()
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    flush = None
    if wl.scanned_bp <= 400e6:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")

    def barrier():
        torch.cuda.synchronize(local_rank)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(local_rank)

    tally_host = [0, 0]

    def one_step():
        # the path shards with no exchange step: ranks never wait for each other inside a step
        res = wl.step_resident()
        n = res.n_guides
        h = res.n_hits
        res.close()
        tally_host[0] += n
        tally_host[1] += h[0] + h[1]
        if flush is not None:
            flush.fill_(1)
        return n, h

    def exchange_counts():
        # the one collective of the timed region: whole-job guide / hit counts, once, at its end
        t = torch.tensor(tally_host, dtype=torch.int64, device=f"cuda:{local_rank}")
        if dist is not None:
            dist.all_reduce(t)
        return t

    torch.cuda.set_stream(stream)  # NCCL / flush work is ordered with the library's stream
    sampler = None if args.no_clocks else ClockSampler(local_rank).start()
    for _ in range(max(args.warmup, 3)):
        n_guides, n_hits = one_step()
    exchange_counts()  # the collective / first device tensor of this process is set up outside the timed region
    barrier()
    launches0 = lib.hawk_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    w0 = time.time()
    ev0.record(stream)
    tally_host[0] = tally_host[1] = 0
    step_wall = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        n_guides, n_hits = one_step()
        step_wall.append(1e3 * (time.perf_counter() - ts))
    ts = time.perf_counter()
    job_counts = exchange_counts()
    exchange_ms = 1e3 * (time.perf_counter() - ts)
    ev1.record(stream)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    w1 = time.time()
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(w0, w1) if sampler else {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["not sampled"]}
    launches = lib.hawk_launch_count() - launches0
    # per-kernel times: a separate pass with the library's CUDA-event brackets switched on
    ctx.set_profiling(True)
    ctx.profile()
    prof_steps = max(3, min(args.steps, 10))
    for _ in range(prof_steps):
        one_step()
    prof = ctx.profile()
    ctx.set_profiling(False)
    # CUDA events on the context's stream (the launching stream); the host layer synchronises
    # inside a step, so this is ~= the wall time, which is reported beside it
    step_ms = dev_ms / args.steps
    t = torch.tensor([step_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    scanned_bp_rank = int(wl.scanned_bp)  # of the headline workload (`wl` is the config-5 block later on)
    tot = torch.tensor([wl.scanned_bp, n_guides], dtype=torch.int64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    step_ms = float(t.item())
    total_bp, total_guides = int(tot[0].item()), int(tot[1].item())
    value = total_bp / (step_ms / 1e3)

    # ---- per-kernel roofline (rank 0's kernels) ----
    peak, peak_src = measured_peak()
    hits_total = n_hits[0] + n_hits[1]
    kernels = {}
    per_step = lambda k: prof[k][0] / prof_steps  # noqa: E731  (sum of the kind's brackets per step)
    pack_ms, cand_ms, match_ms, expand_ms, post_ms = (per_step(k) for k in ("pack", "cand", "match", "expand", "post"))
    scan_ms = cand_ms + match_ms + expand_ms  # all of K2
    pack_bytes = wl.pack_algorithmic_bytes()
    scan_bytes = wl.scan_algorithmic_bytes(hits_total)
    profiled = args.workload in TRAFFIC and args.scale == 1.0 and not args.haplotypes
    traffic = lambda k: TRAFFIC.get(args.workload, {}).get(k) if profiled else None  # noqa: E731
    gbs = lambda nbytes, ms: nbytes / ms / 1e6 if ms and nbytes else None  # noqa: E731
    fused = wl.fused_auto()
    k1_name = "fused_scan_kernel" if fused else "pack_kernel"
    survey_bytes = pack_bytes + scan_bytes  # pack (1.625 B/slot) + scan, whatever kernels do them
    if fused:
        # K1 + K2 in one kernel: 1 B/slot of text read, planes written only where later stages read
        # them, 16 B per chunk with a hit -- the bytes THIS kernel has to move; the SURVEY formula
        # (which also charges the 0.625 B/slot of planes the fusion no longer writes) beside it
        pack_bytes = wl.fused_algorithmic_bytes(hits_total)
    kernels["pack_kernel"] = {"name": k1_name, "ms": pack_ms, "algorithmic_bytes": pack_bytes, "gbs": gbs(pack_bytes, pack_ms),
                              "dram_traffic": traffic(k1_name), "launches_per_step": 1}  # fmt: skip
    if fused:
        kernels["pack_kernel"].update({
            "survey_formula_bytes": survey_bytes, "survey_formula_gbs": gbs(survey_bytes, pack_ms),
            "limiter": "instruction issue / integer pipe (ncu: IPC 2.8 of 4, ALU pipe 65 % busy, DRAM 57 %), not HBM",
            "note": "K1 + K2 fused: texts read once, planes kept only around variant bases and for REF, PAM match in the same "
                    "pass. algorithmic_bytes = what this kernel must move; survey_formula_bytes = SURVEY 8(d) pack + scan "
                    "(the work it replaces: pack_kernel + cand_count + match_kernel)"})
    if fused:
        scan_bytes_k2 = 16.0 * hits_total / 2 + 8.0 * hits_total  # what is left outside the fused kernel: entries read, records written
    else:
        scan_bytes_k2 = scan_bytes
    kernels["scan_k2_total"] = {
        "ms": scan_ms, "algorithmic_bytes": scan_bytes_k2, "gbs": gbs(scan_bytes_k2, scan_ms),
        "note": ("fused path: the PAM match runs inside fused_scan_kernel (see pack_kernel); this entry is the segment "
                 "bookkeeping + fused_expand_kernel only" if fused else
                 "all K2 kernels (hapscan, block table, cand_count, match_kernel, expand_kernel, prefix sums); "
                 "algorithmic bytes by the SURVEY 8(d) formula, which still charges the whole 0.125 B/bp case plane "
                 "although the nz summary lets K2 skip it"),
    }  # fmt: skip
    kernels["match_kernel"] = {"ms": match_ms, "dram_traffic": traffic("match_kernel"),
                               "dram_gbs": gbs(traffic("match_kernel"), match_ms),
                               "note": "the PAM match proper: sparse sector reads around variants, bound by DRAM traffic"}  # fmt: skip
    kernels["candidate_kernels"] = {"ms": cand_ms}
    kernels["expand_kernel"] = {"ms": expand_ms}
    table_bytes = wl.table_algorithmic_bytes(n_guides, hits_total)
    kernels["table_pipeline"] = {
        "ms": post_ms, "algorithmic_bytes": table_bytes, "gbs": gbs(table_bytes, post_ms),
        "note": "everything downstream of the hit lists (coordinates, unphased resolution, redundancy filter, rows, "
                "bucket ids); algorithmic bytes = guide rows written + hit records read + window planes read",
    }  # fmt: skip
    dom = max((("pack_kernel", pack_ms), ("scan_k2_total", scan_ms), ("table_pipeline", post_ms)), key=lambda t: t[1])[0]
    ach = kernels[dom]["gbs"] or 0.0
    roofline = {
        "kernel": kernels[dom].get("name", dom), "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
        "traffic": traffic(kernels[dom].get("name", dom)), "traffic_source": "profiles/traffic.json (ncu --set full capture of this workload)" if traffic(kernels[dom].get("name", dom)) else None,
        "peak_source": peak_src,
        "step_survey_formula_frac": (survey_bytes / step_ms / 1e6) / peak,
        "step_with_table_frac": ((survey_bytes + table_bytes) / step_ms / 1e6) / peak,
        "scan_kernel_frac": (kernels["scan_k2_total"]["gbs"] or 0.0) / peak,
        "pack_kernel_frac": (kernels["pack_kernel"]["gbs"] or 0.0) / peak,
        "table_pipeline_frac": (kernels["table_pipeline"]["gbs"] or 0.0) / peak,
        "match_kernel_dram_frac": (kernels["match_kernel"]["dram_gbs"] or 0.0) / peak if profiled and not fused else None,
    }  # fmt: skip

    # ---- e2e: host buffers in, host table out ----
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        wl.batch.close()
        wl.batch = None
        wl.host_buffers()
        wl.step_host()  # warm-up (allocates the pinned output buffers)
        wl.step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            table, h2d, d2h = wl.step_host()
        torch.cuda.synchronize(local_rank)
        e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
        te = torch.tensor([e_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_ms = float(te.item())
        e2e = {"value": total_bp / (e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e_ms, "steps": e2e_steps, "rows_per_step": int(len(table["hap"])),
               "call": "hawk_search_stream: one C-ABI call, haplotype groups pipelined, H2D / compute / D2H overlapped",
               "input": "haplotype texts (pinned ASCII slot space), the reference's own input to this path"}  # fmt: skip
        if unphased:
            e2e["call"] = ("hawk_batch_create + hawk_batch_set_posmap + hawk_batch_set_alleles + hawk_search + "
                           "hawk_result_fetch (the streamed call is phased / variant-free only)")

        def timed(fn, reps):
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                out = fn()
            torch.cuda.synchronize(local_rank)
            ms = 1e3 * (time.perf_counter() - t0) / reps
            tt = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item()), out

    if e2e is not None and not unphased:
        # the same as two calls with nothing overlapped (encode_haplotypes, then search + fetch)
        two_ms, (_, h2d_b, d2h_b) = timed(wl.step_host_twocall, max(1, e2e_steps - 1))
        e2e["two_calls_no_overlap"] = {"value": total_bp / (two_ms / 1e3), "unit": UNIT, "ms_per_step": two_ms,
                                       "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_b)}  # fmt: skip
        # N1 (next row): the same search when the host holds the reference text and per-haplotype
        # edit lists (what its VCF reader produced) and the texts are materialised on the device
        n1_ms, (table2, h2d2, d2h2) = timed(wl.step_edits, e2e_steps)
        e2e["from_edit_lists"] = {"value": total_bp / (n1_ms / 1e3), "unit": UNIT, "ms_per_step": n1_ms,
                                  "h2d_bytes_per_step": int(h2d2), "d2h_bytes_per_step": int(d2h2),
                                  "rows_per_step": int(len(table2["hap"])), "call": "hawk_search_stream_edits"}  # fmt: skip
        n1b_ms, _ = timed(wl.step_edits_twocall, max(1, e2e_steps - 1))
        e2e["from_edit_lists"]["two_calls_no_overlap_ms"] = n1b_ms
        # ... and with the window-text column left on the device (hawk_table_out.text = NULL): 21 B per row
        # come down instead of 69; the text is a slice of the haplotype the host can rebuild lazily
        n1c_ms, (table3, h2d3, d2h3) = timed(lambda: wl.step_edits(want_text=False), e2e_steps)
        e2e["from_edit_lists"]["rows_only"] = {"value": total_bp / (n1c_ms / 1e3), "unit": UNIT, "ms_per_step": n1c_ms,
                                               "h2d_bytes_per_step": int(h2d3), "d2h_bytes_per_step": int(d2h3),
                                               "rows_per_step": int(len(table3["hap"]))}  # fmt: skip

    # ---- final merge (N > 1): every rank pushes its rows into rank 0's buffer over NVLink ----
    def measure_merge(w, result, hap_add, with_text):
        """Three merges (the first one warms the IPC mapping); best of the other two."""
        from crispr_hawk_b200 import shard

        dev = f"cuda:{local_rank}"
        key_min, key_span = w.cohort.region_start, w.cohort.region_stop - w.cohort.region_start + 1
        m_ms, out = [], None
        session = shard.MergeSession(ctx, rank, world, dev)
        for _ in range(3):  # the first merge sets the session up (allocation + CUDA IPC mapping, ~40 ms once)
            barrier()
            t0 = time.perf_counter()
            merged = shard.merge_tables_device(result, ctx, hap_add, rank, world, dev, key_min, key_span,
                                               with_text=with_text, session=session)
            barrier()
            m_ms.append(1e3 * (time.perf_counter() - t0))
            if rank == 0:
                if out is None or m_ms[-1] <= min(m_ms[1:] or m_ms):
                    phases = {"push_ms": merged.push_ms, "first_seen_ms": merged.first_seen_ms}
                out = {"rows": int(merged["hap"].shape[0]), "received_bytes": int(merged.received_bytes),
                       "first_merge_ms_with_session_setup": m_ms[0]}
            del merged
        session.close()
        if rank != 0:
            return None
        best = min(m_ms[1:])
        out.update(ms=best, gbs=out["received_bytes"] / (best / 1e3) / 1e9, with_text=with_text,
                   push_ms=phases["push_ms"], push_gbs=out["received_bytes"] / (phases["push_ms"] / 1e3) / 1e9,
                   first_seen_ms=phases["first_seen_ms"])
        return out

    final_merge = None
    if world > 1 and not args.no_e2e and not unphased:
        wl.prepare_resident()
        res = wl.step_resident()
        full = measure_merge(wl, res, rank * n_alt, True)
        slim = measure_merge(wl, res, rank * n_alt, False)
        if rank == 0:
            final_merge = dict(full)
            final_merge["rows_only"] = slim
            final_merge["what"] = (
                "one-sided push over NVLink: rank 0 owns ONE buffer for the merged table (CUDA IPC), every other rank "
                "writes its guide rows (REF rows dropped, haplotype indices shifted) straight into its slice, all ranks "
                "at once (hawk_result_push); rank 0 then computes the first-seen bucket ids in place "
                "(hawk_first_seen_dev). rows_only: the same without the window text (17 B per row; the host that owns "
                "the haplotype texts slices it). Outside the timed step")
        res.close()
        wl.batch.close()
        wl.batch = None

    # ---- the chromosome-scale configuration (BASELINE config 5) on the same ranks: one 626-haplotype
    # block of a 50 Mb region per GPU (N x 31.3 G haplotype-bp per step), steps + the merge ----
    c5 = None
    if world > 1 and not args.no_e2e and not args.no_c5 and args.workload == "c2" and args.scale == 1.0:
        del wl
        torch.cuda.empty_cache()
        k5 = synth.CONFIGS["c5shard"]
        cohort5 = synth.config_cohort("c5shard", 1.0, hap_block=rank)
        wl5 = Workload(cohort5, k5["pam"], k5["guidelen"], k5["right"], ctx, local_rank)
        wl5.prepare_resident()
        for _ in range(2):
            wl5.step_resident().close()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n5 = 5
        e0.record(stream)
        for _ in range(n5):
            r5 = wl5.step_resident()
            g5 = r5.n_guides
            r5.close()
        e1.record(stream)
        barrier()
        t5 = torch.tensor([e0.elapsed_time(e1) / n5], dtype=torch.float64, device=f"cuda:{local_rank}")
        b5 = torch.tensor([wl5.scanned_bp, g5], dtype=torch.int64, device=f"cuda:{local_rank}")
        dist.all_reduce(t5, op=dist.ReduceOp.MAX)
        dist.all_reduce(b5)
        r5 = wl5.step_resident()
        wl5.ascii_dev = None  # the texts are not needed for the merge: make room for the merged table
        torch.cuda.empty_cache()
        m5 = measure_merge(wl5, r5, rank * k5["n_alt_hap"], False)
        r5.close()
        if rank == 0:
            ms5 = float(t5.item())
            c5 = {"workload": f"c5: NGG / 20 nt, 50,000,000 bp region, {k5['n_alt_hap'] + 1} haplotypes per rank "
                              f"({world * k5['n_alt_hap']} + REF over the job)",
                  "scanned_bp_per_step": int(b5[0].item()), "guides_per_step": int(b5[1].item()), "ms_per_step": ms5,
                  "value": int(b5[0].item()) / (ms5 / 1e3), "unit": UNIT, "steps": n5, "final_merge_rows_only": m5}  # fmt: skip
        wl = wl5

    # ---- next rows of the scope table on the same workload (rank 0, N = 1 only) ----
    next_rows = None
    if rank == 0 and world == 1 and not args.no_e2e and not unphased:
        from oracle import annot_oracle

        m = wl.annotate_measure(oracle=annot_oracle)
        ms = m["variants_ms"] + m["text_gc_ms"]
        next_rows = {"N2_annotate": {
            "what": "hawk_result_annotate over the step's whole guide table: polish_guide_variants (CSR lists of the "
                    "haplotype's variants visible in each guide), reverse complement of strand-1 rows, GC counts; "
                    "outputs copied to pinned host memory inside the timed calls",
            "rows": m["rows"], "variant_refs": m["variant_refs"], "variants_ms": m["variants_ms"],
            "text_gc_ms": m["text_gc_ms"], "rows_per_s": m["rows"] / (ms / 1e3) if ms else None,
            "cpu_port_rows_per_s": m.get("oracle_rows_per_s"), "cpu_port_rows_checked": m.get("oracle_rows_checked"),
            "cpu_port": "oracle/annot_oracle.py (pure Python, like the reference) on a random sample of non-REF rows, "
                        "each also compared with the device result",
        }}  # fmt: skip
        next_rows["N2_collapse"] = {
            "what": "hawk_result_collapse: the groups of reports._collapse_report_entries over the whole table (row keys, two "
                    "radix sorts, head flags; permutation + flags copied to the host inside the timed call)",
            "rows": m["rows"], "groups": m["collapse_groups"], "ms": m["collapse_ms"], "hash_collision": m["collapse_collision"]}
        if "kmers_ms" in m:
            next_rows["N4_features"] = {
                "what": "hawk_result_featurize: the learned scorers' input strings of every row (scoring.py:50-84, lead 4) copied to "
                        "pinned host memory inside the timed call; DeepCpf1's float32 one-hot tensor (seqdeepcpf1.py:71-92) written "
                        "into a device tensor",
                "rows": m["rows"], "kmers_ms": m["kmers_ms"], "kmers_bytes": m["kmers_bytes"], "kmers_all_acgt": m["kmers_all_acgt"],
                "onehot_ms": m.get("onehot_ms"), "onehot_bytes": m.get("onehot_bytes"), "onehot_sum_ok": m.get("onehot_sum_ok")}
        if "cfdon_ms" in m:
            next_rows["N4_cfdon"] = {
                "what": "hawk_result_cfdon: CFDon of every row against the REF guide of its (start, strand) key, float64 scores "
                        "copied to the host inside the timed call; stand-in factor tables",
                "rows": m["rows"], "rows_with_ref_guide": m["cfdon_scored"], "ms": m["cfdon_ms"]}

    # ---- the Python side of the seam (rank 0, N = 1 only): what a user of the drop-in sees ----
    if rank == 0 and world == 1 and not args.no_e2e and not unphased:
        next_rows = dict(next_rows or {})
        next_rows["api_search"] = api_search_leg()
        next_rows["from_variant_records"] = variant_records_leg()

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import c_oracle

        threads = c_oracle.max_threads()
        n_cpu = max(threads * 2, 16) - 1  # REF + n_cpu = a multiple of the thread count
        s = cpu_sample(args.workload, n_cpu, args.scale)
        if s.get("unphased"):
            n_cpu = s["n_hap"] - 1
        cpu_step(s, threads)
        reps, acc = 0, 0.0
        while acc < 8.0 and reps < 20:
            dt, _, _ = cpu_step(s, threads)
            acc += dt
            reps += 1
        one_dt, _, _ = cpu_step(s, 1) if s["bp"] < 200e6 else (None, None, None)
        py_rate, py_bp, _ = python_port_rate(args.workload, args.scale)
        cpu = {
            "value": s["bp"] * reps / acc, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": (f"the same workload on a region {CPU_C4_FRACTION:g} of the size, all {n_cpu + 1} haplotypes, " if unphased
                       else f"{n_cpu + 1} haplotypes (REF + first {n_cpu}) of the same workload, ")
                      + f"{s['bp']:,} hap-bp per pass, {reps} passes, encode + search (oracle/scan_oracle.c, OpenMP)",
            "single_thread_value": (s["bp"] / one_dt) if one_dt else None,
            "python_port_value": py_rate, "python_port_sample_bp": py_bp,
        }  # fmt: skip

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, scanned_bp_rank, n_hap_rank),
            "guides_per_step": total_guides, "hits_per_step_rank0": hits_total,
            "job_guides_in_timed_region": int(job_counts[0].item()),
            "device_ms_per_step": dev_ms / args.steps, "wall_ms_per_step": wall_ms / args.steps,
            "step_wall_ms": {"min": min(step_wall), "median": statistics.median(step_wall), "max": max(step_wall),
                             "argmax": step_wall.index(max(step_wall)), "count_exchange_ms": exchange_ms},
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e, "final_merge": final_merge, "c5": c5, "next_rows": next_rows,
            "gpu_launches": int(launches), "clocks": clocks, "profile_steps": prof_steps,
        }  # fmt: skip
        emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
# (profiles/); filled in once a capture of this workload exists
TRAFFIC = {}
_traffic = os.path.join(ROOT, "profiles", "traffic.json")
if os.path.exists(_traffic):
    try:
        TRAFFIC = json.load(open(_traffic))
    except Exception:
        TRAFFIC = {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the region (debugging only)")
    ap.add_argument("--haplotypes", type=int, default=0, help="non-REF haplotypes per rank (default: the config's)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="N > 1: skip the chromosome-scale (config 5) block")
    ap.add_argument("--no-clocks", action="store_true", help="do not run nvidia-smi beside the timed region")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        # the CPU arm builds and loads the checker only: libhawkscan.so is never opened here
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        return run_reference_arm(args, rank, world)
    import __graft_entry__ as entry

    entry.build()  # every rank: serialised by a file lock, a no-op when the libraries are up to date
    return run_product_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
