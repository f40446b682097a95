"""Device-resident synthetic workloads for bench.py and the large GPU tests.

One `Workload` = one region: the cohort (host), its haplotype texts materialised in
HBM in the slot layout, the derived flat arrays, and the two ways of running the hot
path over it: `step_resident` (inputs already in HBM) and `step_host` (everything
starts and ends in host memory, through the host layer of the C-ABI)."""

from __future__ import annotations

from typing import Optional

import numpy as np

from . import _cabi, marshal, synth, synth_unphased
from .pam import pam_patterns


class Workload:
    def __init__(self, cohort: synth.Cohort, pam: str, guidelen: int, right: bool,
                 ctx: Optional[_cabi.Context] = None, device: Optional[int] = None):  # fmt: skip
        import torch

        self.torch = torch
        self.cohort = cohort
        self.pam, self.guidelen, self.right = pam, guidelen, right
        self.device = torch.cuda.current_device() if device is None else device
        self.ctx = ctx or _cabi.Context.default(self.device)
        self.d = synth.derive(cohort)
        self.fwd, self.rc = pam_patterns(pam)
        self.params = _cabi.make_params(self.fwd, self.rc, guidelen, right, False)
        self.a, self.b = synth.scan_bounds(cohort, len(self.fwd))
        self.scanned_bp = int((self.b.astype(np.int64) - self.a).clip(min=0).sum())
        self.ascii_dev = synth.materialize_device(cohort, self.ctx, self.device)
        self.batch = None

    # ---- sizes for the roofline (algorithmic bytes) ----
    def ref_scanned_bp(self) -> int:
        ref = self.d.is_ref.astype(bool)
        return int((self.b.astype(np.int64) - self.a)[ref].clip(min=0).sum())

    def scan_algorithmic_bytes(self, n_hits: int) -> float:
        """0.5 B/bp of planes for REF haplotypes; for the others 0.125 B/bp of case plane
        plus 0.5 B for every base inside a variant window (2(G+P)-1 bases around each
        variant base, capped by the scanned length); 8 B per emitted hit record."""
        ref_bp = self.ref_scanned_bp()
        alt_bp = self.scanned_bp - ref_bp
        span = 2 * (self.guidelen + len(self.fwd)) - 1
        window_bp = min(alt_bp, self.d.variant_bases * span)
        return 0.5 * ref_bp + 0.125 * alt_bp + 0.5 * window_bp + 8.0 * n_hits

    def pack_algorithmic_bytes(self) -> float:
        """1 B ASCII read + 0.5 B planes + 0.125 B case bits written per slot."""
        return 1.625 * self.d.total_slots

    def fused_auto(self) -> bool:
        """Whether hawk_encode_search_dev picks the fused K1 + K2 kernel for this workload
        (api.cu: whenever the guide geometry has the fast form, G <= 32 and G + P <= 33)."""
        return self.guidelen <= 32 and self.guidelen + len(self.fwd) <= 33

    def fused_algorithmic_bytes(self, n_hits: int, reach: int = 2) -> float:
        """What the fused kernel has to move: 1 B of text read per slot; 0.625 B of planes written
        per slot it keeps -- every REF slot and the chunks within `reach` chunks of a variant base
        (an upper bound: windows of neighbouring variants overlap), never more than all slots; one
        16-byte entry per chunk with a hit (about one per two hit records)."""
        if self.params.flags & _cabi.HAWK_F_UNPHASED:
            kept = float(self.d.total_slots)  # IUPAC codes every few bases: every chunk is kept
        else:
            ref_slots = float(self.d.lens[self.d.is_ref.astype(bool)].sum())
            kept = min(float(self.d.total_slots), ref_slots + float(self.d.variant_bases) * (2 * reach + 1) * 32)
        return 1.0 * self.d.total_slots + 0.625 * kept + 16.0 * n_hits / 2

    def table_algorithmic_bytes(self, n_rows: int, n_hits: int) -> float:
        """Guide-table pipeline: rows written (25 B of columns + the padded text row), hit records
        read (8 B), window planes read per hit (0.625 B per window base)."""
        w = self.guidelen + len(self.fwd) + 2 * marshal.GUIDESEQPAD
        return n_rows * (25.0 + (w + 15) // 16 * 16) + n_hits * (8.0 + 0.625 * w)

    # ---- the hot path, inputs resident in HBM ----
    def prepare_resident(self) -> None:
        self.batch = _cabi.Batch(self.ctx, None, self.d.slot_off, self.d.lens, device_ptr=self.ascii_dev.data_ptr())
        self.batch.set_posmap(self.d.seg)
        self.batch.set_scan(self.a, self.b, self.d.is_ref)  # bounds are batch metadata: uploaded once

    def step_resident(self, fused=None) -> _cabi.Result:
        """encode + search over the texts resident in HBM; the table stays on the device.
        `fused`: True = hawk_encode_search_dev with the fused K1 + K2 kernel forced, None = the
        same call choosing by haplotype shape (the default of the library), False = the two
        calls hawk_batch_repack_dev (K1) + hawk_search (staged K2)."""
        if self.batch is None:
            self.prepare_resident()
        if fused is None or fused:
            self.ctx.set_fused(2 if fused is None else 1)
            return _cabi.encode_search(self.ctx, self.batch, self.ascii_dev.data_ptr(), self.params)
        self.batch.repack(self.ascii_dev.data_ptr())
        return _cabi.search(self.ctx, self.batch, self.params)

    # ---- the hot path through host buffers ----
    def host_buffers(self):
        """Pinned host copies of the inputs and pinned output buffers (allocated once)."""
        torch = self.torch
        if not hasattr(self, "_host"):
            ascii_host = torch.empty(self.d.total_slots, dtype=torch.uint8, pin_memory=True)
            ascii_host.copy_(self.ascii_dev)
            torch.cuda.synchronize(self.device)
            self._host = {"ascii": ascii_host, "out": None}
        return self._host

    def step_host(self, n_groups: int = 0):
        """Host ASCII in, host guide table out through ONE call (hawk_search_stream): the H2D
        copies of the texts, K1, K2, the table pipeline and the D2H copies of the rows all
        inside, group by group, PCIe traffic of both directions overlapped. Returns
        (table, h2d bytes, d2h bytes) -- the byte counts are the library's own counters."""
        hb = self.host_buffers()
        t0 = self.ctx.traffic()
        res = _cabi.search_stream(self.ctx, hb["ascii"].numpy(), self.d.slot_off, self.d.lens, self.d.seg, self.params,
                                  self.a, self.b, self.d.is_ref, n_groups=n_groups, buffers=hb["out"], pinned=True)  # fmt: skip
        hb["out"] = res.buffers
        t1 = self.ctx.traffic()
        self.last_stream = res
        return res.table(), t1[0] - t0[0], t1[1] - t0[1]

    def step_host_twocall(self):
        """The same as two calls, the way the reference's driver is written (encode_haplotypes,
        then search): the whole batch goes up, is searched, and the table comes down, nothing
        overlapped."""
        torch = self.torch
        hb = self.host_buffers()
        t0 = self.ctx.traffic()
        batch = _cabi.Batch(self.ctx, hb["ascii"].numpy(), self.d.slot_off, self.d.lens)
        batch.set_posmap(self.d.seg)
        res = _cabi.search(self.ctx, batch, self.params, self.a, self.b, self.d.is_ref)
        n, w = res.n_guides, res.text_stride
        if hb.get("out2") is None or len(hb["out2"]["hap"]) < n:
            hb["out2"] = _cabi.alloc_table(int(n * 1.05) + 1024, w, pinned=True)
        table = res.table(hb["out2"])
        res.close()
        batch.close()
        t1 = self.ctx.traffic()
        return table, t1[0] - t0[0], t1[1] - t0[1]

    def _edit_buffers(self):
        torch = self.torch
        c = self.cohort
        if not hasattr(self, "_edits"):
            sites = c.hap_sites
            pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
            self._edits = dict(ref=pin(c.ref), off=pin(c.hap_off), pos=pin(c.site_pos[sites]), rl=pin(c.site_reflen[sites]),
                               al=pin(c.site_altlen[sites]), ao=pin(c.site_altoff[sites]),
                               pool=pin(c.alt_pool if len(c.alt_pool) else np.zeros(1, np.uint8)))  # fmt: skip
            self._edits_out = None
        return self._edits

    def step_edits(self, n_groups: int = 0, want_text: bool = True):
        """N1 path through ONE call (hawk_search_stream_edits): only the reference text and the
        per-haplotype edit lists start on the host; texts are materialised on the device group
        by group, then K1, K2 and the table pipeline as usual, while the previous group's guide
        rows leave for pinned host memory."""
        c = self.cohort
        e = self._edit_buffers()
        t0 = self.ctx.traffic()
        key = "_edits_out" if want_text else "_edits_out_slim"
        res = _cabi.search_stream_edits(self.ctx, e["ref"], c.region_start, e["off"], e["pos"], e["rl"], e["al"], e["ao"],
                                        e["pool"], self.params, self.a, self.b, self.d.is_ref, n_groups=n_groups,
                                        buffers=getattr(self, key, None), pinned=True, want_text=want_text)  # fmt: skip
        setattr(self, key, res.buffers)
        t1 = self.ctx.traffic()
        return res.table(), t1[0] - t0[0], t1[1] - t0[1]

    def step_edits_twocall(self):
        """hawk_batch_create_from_edits + hawk_search + fetch, nothing overlapped."""
        c = self.cohort
        e = self._edit_buffers()
        t0 = self.ctx.traffic()
        batch = _cabi.Batch.from_edits(self.ctx, e["ref"], c.region_start, e["off"], e["pos"], e["rl"], e["al"], e["ao"], e["pool"])
        res = _cabi.search(self.ctx, batch, self.params, self.a, self.b, self.d.is_ref)
        n, w = res.n_guides, res.text_stride
        if getattr(self, "_edits_out2", None) is None or len(self._edits_out2["hap"]) < n:
            self._edits_out2 = _cabi.alloc_table(int(n * 1.05) + 1024, w, pinned=True)
        table = res.table(self._edits_out2)
        res.close()
        batch.close()
        t1 = self.ctx.traffic()
        return table, t1[0] - t0[0], t1[1] - t0[1]

    # ---- N2: post-search pure functions on the resident table ----
    def annotate_measure(self, reps: int = 3, n_sample: int = 192, oracle=None):
        """hawk_result_annotate over the whole guide table of this workload (batch built from
        the edit lists, which stay on the device as the variant table). Returns timings and,
        when `oracle` (oracle.annot_oracle, tests / bench only) is given, checks a random
        sample of non-REF rows against it and times it."""
        import time

        torch = self.torch
        c = self.cohort
        e = self._edit_buffers()
        batch = _cabi.Batch.from_edits(self.ctx, e["ref"], c.region_start, e["off"], e["pos"], e["rl"], e["al"], e["ao"], e["pool"])
        res = _cabi.search(self.ctx, batch, self.params, self.a, self.b, self.d.is_ref)
        n, ts = res.n_guides, res.text_stride
        pin = lambda m, dt: torch.empty(max(m, 1), dtype=dt, pin_memory=True).numpy()  # noqa: E731
        bufs = {"rc_text": pin(n * ts, torch.uint8), "gc_num": pin(n, torch.int32), "gc_den": pin(n, torch.int32),
                "gv_off": pin(n + 1, torch.int64), "gv_idx": pin(2 * n + 16, torch.int32)}  # fmt: skip
        ann = res.annotate(batch, buffers=bufs)
        out = {"rows": int(n), "variant_refs": int(len(ann["gv_idx"]))}
        for key, kw in (("variants_ms", dict(want_text=False)), ("text_gc_ms", dict(want_variants=False))):
            best = None
            for _ in range(max(1, reps)):
                torch.cuda.synchronize(self.device)
                t0 = time.perf_counter()
                res.annotate(batch, buffers=bufs, **kw)
                dt = 1e3 * (time.perf_counter() - t0)
                best = dt if best is None else min(best, dt)
            out[key] = best  # best of `reps` calls: the 0.8 GB going down is sensitive to the host's state
        # N2 row collapse (groups of the report) and N4 CFDon (Cas9 PAMs) on the same table
        def best_of(fn):
            best, val = None, None
            for _ in range(max(1, reps)):
                torch.cuda.synchronize(self.device)
                t0 = time.perf_counter()
                val = fn()
                dt = 1e3 * (time.perf_counter() - t0)
                best = dt if best is None else min(best, dt)
            return best, val

        cbufs = {"perm": pin(n, torch.int32).view(np.uint32), "head": pin(n, torch.uint8), "cfd": pin(n, torch.float64)}
        out["collapse_ms"], (perm, head, collision) = best_of(lambda: res.collapse(self.d.is_ref, cbufs))
        out["collapse_groups"], out["collapse_collision"] = int(head.sum()), bool(collision)
        if len(self.fwd) >= 2 and not self.right:
            import random

            from . import scoring

            rnd = random.Random(7)  # stand-in factor tables (the reference's are model files)
            mm = {f"r{w}:d{g},{i + 1}": rnd.random() for i in range(20) for w in "ACGU" for g in "ACGT"}
            pam = {a + b: rnd.random() for a in "ACGT" for b in "ACGT"}
            mm_t, pam_t = scoring.cfd_tables(mm, pam)
            out["cfdon_ms"], col = best_of(lambda: res.cfdon(self.d.is_ref, mm_t, pam_t, cbufs["cfd"]))
            out["cfdon_scored"] = int((~np.isnan(col)).sum())
        # N4, second half: the learned scorers' input strings (host, pinned) and DeepCpf1's one-hot
        # tensor written straight into a device tensor
        L = res.window - 20 + 7
        kbuf = pin(n * L, torch.uint8)
        out["kmers_ms"], (k4, _) = best_of(lambda: res.featurize(lead=4, kmers_out=kbuf))
        out["kmers_bytes"] = int(n * L)
        acgt = np.zeros(256, bool)
        acgt[list(b"ACGT")] = True
        out["kmers_all_acgt"] = bool(acgt[k4].all()) if n else True
        if n * 4 * L * 4 <= 24 << 30:
            dev = torch.empty((max(n, 1), 4, L), dtype=torch.float32, device=self.device)
            out["onehot_ms"], _ = best_of(lambda: res.featurize(lead=4, kmers=False, onehot=True, onehot_device_ptr=dev.data_ptr()))
            out["onehot_bytes"] = int(n * 4 * L * 4)
            out["onehot_sum_ok"] = bool(int(dev[:n].sum(dtype=torch.float64).item()) == n * L)
            del dev
        if oracle is not None and n:
            if getattr(self, "_edits_out2", None) is None or len(self._edits_out2["hap"]) < n:
                self._edits_out2 = _cabi.alloc_table(int(n * 1.05) + 1024, ts, pinned=True)
            table = res.table(self._edits_out2)
            rng = np.random.default_rng(7)
            alt_rows = np.flatnonzero(table["hap"] != 0)
            rows = rng.choice(alt_rows, size=min(n_sample, len(alt_rows)), replace=False) if len(alt_rows) else []
            G, P = self.guidelen, len(self.fwd)
            seg = self.d.seg
            pool = c.alt_pool.tobytes().decode("ascii")
            reft = c.ref.tobytes().decode("ascii")
            ids_cache = {}
            cpu_s = 0.0
            for r in rows:
                h, s_ = int(table["hap"][r]), int(table["strand"][r])
                if h not in ids_cache:
                    sites = c.hap_sites[c.hap_off[h] : c.hap_off[h + 1]]
                    ids_cache[h] = [
                        f"chr1-{c.region_start + int(c.site_pos[k])}-{reft[int(c.site_pos[k]) : int(c.site_pos[k]) + int(c.site_reflen[k])]}/"
                        f"{pool[int(c.site_altoff[k]) : int(c.site_altoff[k]) + int(c.site_altlen[k])]}" for k in sites
                    ]  # fmt: skip
                ids = ids_cache[h]
                rp = (not self.right) if s_ == 1 else bool(self.right)
                pivot = int(table["pos"][r]) - (0 if rp else G)
                s0, s1 = int(seg.seg_off[h]), int(seg.seg_off[h + 1])
                pm = marshal.eval_segments(seg.seg_rel[s0:s1], seg.seg_gen[s0:s1], seg.seg_step[s0:s1],
                                           np.arange(pivot, pivot + G + P)).tolist()  # fmt: skip
                text = table["text"][r].tobytes().decode("ascii")
                t0 = time.perf_counter()
                v, _, seq2, _, gc = oracle.annotate_guide(text, G, P, s_, rp, int(table["stop"][r]), pm, ",".join(ids),
                                                          {x: 0.5 for x in ids})  # fmt: skip
                cpu_s += time.perf_counter() - t0
                got_v = ",".join(sorted(ids[j] for j in ann["gv_idx"][ann["gv_off"][r] : ann["gv_off"][r + 1]]))
                got_seq = ann["rc_text"][r].tobytes().decode("ascii")
                got_gc = str(int(ann["gc_num"][r]) / int(ann["gc_den"][r]))
                if (got_v, got_seq, got_gc) != (v, seq2, gc):
                    raise AssertionError(f"N2 mismatch at row {int(r)}: {(got_v, got_seq, got_gc)} != {(v, seq2, gc)}")
            out["oracle_rows_checked"] = int(len(rows))
            out["oracle_rows_per_s"] = (len(rows) / cpu_s) if cpu_s > 0 else None
        res.close()
        batch.close()
        return out

    def host_arrays_for_oracle(self, hap_indices):
        """(ascii slots, slot_off, lens, a, b, is_ref, segments) of a subset of haplotypes,
        copied to the host, in the form oracle/c_oracle.search takes."""
        idx = np.asarray(hap_indices)
        lens = self.d.lens[idx]
        off, total = marshal.layout(lens)
        buf = np.zeros(total, np.uint8)
        for k, h in enumerate(idx):
            s = int(self.d.slot_off[h])
            n = int(lens[k])
            buf[off[k] : off[k] + n] = self.ascii_dev[s : s + n].cpu().numpy()
        so = self.d.seg.seg_off
        cnt = (so[idx + 1] - so[idx]).astype(np.int64)
        seg_off = np.concatenate(([0], np.cumsum(cnt)))
        take = np.concatenate([np.arange(so[h], so[h + 1]) for h in idx]) if len(idx) else np.zeros(0, np.int64)
        seg = marshal.SegmentTable(seg_off.astype(np.int64), self.d.seg.seg_rel[take], self.d.seg.seg_gen[take],
                                   self.d.seg.seg_step[take])  # fmt: skip
        return buf, off, lens, self.a[idx], self.b[idx], self.d.is_ref[idx], seg


class UnphasedWorkload:
    """BASELINE config 4: an unphased cohort (synth_unphased) -- REF, the per-sample IUPAC SNV
    haplotypes and the per-indel window haplotypes -- searched with `variants_present and not
    phased` semantics (is_pamhit_valid + resolve_guide, search_guides.py:473-479). Same two ways
    of running the path as `Workload`; the texts (a few hundred MB) are built on the host by the
    generator and uploaded once."""

    def __init__(self, cohort: synth_unphased.UnphasedCohort, pam: str, guidelen: int, right: bool,
                 ctx: Optional[_cabi.Context] = None, device: Optional[int] = None,
                 uset: Optional[synth_unphased.UnphasedSet] = None):  # fmt: skip
        """`uset`: search these haplotypes (a subset / reordering of the cohort's) instead of all."""
        import torch

        self.torch = torch
        self.cohort = cohort
        self.pam, self.guidelen, self.right = pam, guidelen, right
        self.device = torch.cuda.current_device() if device is None else device
        self.ctx = ctx or _cabi.Context.default(self.device)
        self.d = uset if uset is not None else synth_unphased.derive_unphased(cohort)
        self.fwd, self.rc = pam_patterns(pam)
        self.params = _cabi.make_params(self.fwd, self.rc, guidelen, right, True)
        self.a, self.b = self.d.scan_bounds(cohort, len(self.fwd))
        self.scanned_bp = int((self.b.astype(np.int64) - self.a).clip(min=0).sum())
        self.ascii_dev = torch.from_numpy(self.d.ascii).to(torch.device("cuda", self.device))
        self.batch = None

    def prepare_resident(self) -> None:
        self.batch = _cabi.Batch(self.ctx, None, self.d.slot_off, self.d.lens, device_ptr=self.ascii_dev.data_ptr())
        self.batch.set_posmap(self.d.seg)
        self.batch.set_alleles(self.d.alleles)
        self.batch.set_scan(self.a, self.b, self.d.is_ref)  # bounds are batch metadata: uploaded once

    def step_resident(self, fused=None) -> _cabi.Result:
        """encode + unphased search; texts resident in HBM, table stays on the device (see
        Workload.step_resident for `fused`)."""
        if self.batch is None:
            self.prepare_resident()
        if fused is None or fused:
            self.ctx.set_fused(2 if fused is None else 1)
            return _cabi.encode_search(self.ctx, self.batch, self.ascii_dev.data_ptr(), self.params)
        self.batch.repack(self.ascii_dev.data_ptr())
        return _cabi.search(self.ctx, self.batch, self.params)

    def host_buffers(self):
        torch = self.torch
        if not hasattr(self, "_host"):
            pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
            A, S = self.d.alleles, self.d.seg
            self._host = {
                "ascii": pin(self.d.ascii), "out": None,
                "seg": marshal.SegmentTable(pin(S.seg_off), pin(S.seg_rel), pin(S.seg_gen), pin(S.seg_step)),
                "alleles": marshal.AlleleTable(pin(A.va_off), pin(A.va_idx), pin(A.va_ent_off), pin(A.va_ref)),
            }  # fmt: skip
        return self._host

    def step_host(self, n_groups: int = 0):
        """Everything starts and ends in host memory: texts, position maps and variant_alleles
        tables go up, the batch is packed and searched, the guide table comes down (the calls
        `encode_haplotypes` + `search` make through the C-ABI). Returns (table, h2d, d2h bytes)."""
        hb = self.host_buffers()
        t0 = self.ctx.traffic()
        batch = _cabi.Batch(self.ctx, hb["ascii"], self.d.slot_off, self.d.lens)
        batch.set_posmap(hb["seg"])
        batch.set_alleles(hb["alleles"])
        res = _cabi.search(self.ctx, batch, self.params, self.a, self.b, self.d.is_ref)
        n, w = res.n_guides, res.text_stride
        if hb["out"] is None or len(hb["out"]["hap"]) < n:
            hb["out"] = _cabi.alloc_table(int(n * 1.05) + 1024, w, pinned=True)
        table = res.table(hb["out"])
        res.close()
        batch.close()
        t1 = self.ctx.traffic()
        return table, t1[0] - t0[0], t1[1] - t0[1]

    step_host_twocall = step_host

    # ---- sizes for the roofline (algorithmic bytes) ----
    def pack_algorithmic_bytes(self) -> float:
        return 1.625 * self.d.total_slots

    def scan_algorithmic_bytes(self, n_hits: int) -> float:
        """Every haplotype of an unphased cohort is dense in variant bases (IUPAC codes every few
        bases), so every scanned base is read: 0.5 B/bp planes, 0.125 B/bp case plane for the
        non-REF haplotypes, 8 B per emitted hit record."""
        bp = (self.b.astype(np.int64) - self.a).clip(min=0)
        ref_bp = int(bp[self.d.is_ref.astype(bool)].sum())
        return 0.5 * self.scanned_bp + 0.125 * (self.scanned_bp - ref_bp) + 8.0 * n_hits

    table_algorithmic_bytes = Workload.table_algorithmic_bytes
    fused_auto = Workload.fused_auto
    fused_algorithmic_bytes = Workload.fused_algorithmic_bytes

    def oracle_subset(self, hap_indices):
        """The flat arrays of a subset of haplotypes (oracle/c_oracle.search's inputs)."""
        u = self.d.take(hap_indices)
        a, b = u.scan_bounds(self.cohort, len(self.fwd))
        return u, a, b
