"""TEST SUPPORT ONLY: a stand-in for the device layer (`_cabi.Batch`, `_cabi.search`,
`Result.table / annotate`) that runs the kernels' own __host__ __device__ core compiled for the
CPU (libhawkcheck.so through tests/hostcheck.py). It exists so that the drop-in's HOST logic --
install() on the genuine `crisprhawk.crisprhawk` module, the lazy GuideList, the annotation
seam, the eviction fallback -- can be driven by the reference's real driver functions in the
build container, which has the reference but no GPU. Never importable from the product."""

from __future__ import annotations

import numpy as np

from crispr_hawk_b200 import _cabi, encoder, search_guides
from tests import hostcheck


class FakeContext:
    def info(self):
        return {"sm_count": 0, "total_mem": 32 << 30, "free_mem": 32 << 30}


class FakeBatch:
    def __init__(self, texts):
        self.ctx, self.texts, self.n_hap = FakeContext(), list(texts), len(texts)
        self.has_posmap = self.has_alleles = self.has_variants = False
        self.seg = self.va = self.vt = None
        self.closed = False

    def set_posmap(self, seg):
        self.seg, self.has_posmap = seg, True

    def set_alleles(self, va):
        self.va, self.has_alleles = va, True

    def set_variants(self, vt):
        self.vt, self.has_variants = vt, True

    def export_nibbles(self, i):
        from oracle import c_oracle

        return c_oracle.encode(self.texts[i].encode("ascii"))

    def device_bytes(self):
        return sum(len(t) for t in self.texts)

    def close(self):
        self.closed = True


def fake_from_edits(ctx, ref_ascii, region_start, edit_off, edit_pos, edit_reflen, edit_altlen, edit_altoff, alt_pool):
    """hawk_batch_create_from_edits on the host: texts (ALT characters lower-case), lengths and
    run-length coordinate maps by the reference's conventions (haplotype.py:106-159)."""
    from crispr_hawk_b200 import marshal

    ref = bytes(np.asarray(ref_ascii, np.uint8)).decode("ascii")
    pool = bytes(np.asarray(alt_pool, np.uint8)).decode("ascii")
    texts, posmaps = [], []
    for h in range(len(edit_off) - 1):
        parts, pm, cur = [], [], 0
        for e in range(int(edit_off[h]), int(edit_off[h + 1])):
            p, rl, al, ao = int(edit_pos[e]), int(edit_reflen[e]), int(edit_altlen[e]), int(edit_altoff[e])
            parts.append(ref[cur:p])
            pm.extend(range(region_start + cur, region_start + p))
            parts.append(pool[ao : ao + al].lower())
            pm.extend([region_start + p] * al)  # inserted bases repeat the anchor's coordinate
            cur = p + rl
        parts.append(ref[cur:])
        pm.extend(range(region_start + cur, region_start + len(ref)))
        texts.append("".join(parts))
        posmaps.append(np.asarray(pm, np.int64))
    b = FakeBatch(texts)
    offs, rels, gens, steps = [0], [], [], []
    for pm in posmaps:
        r, g, st = marshal.posmap_segments(pm)
        rels.append(r), gens.append(g), steps.append(st)
        offs.append(offs[-1] + len(r))
    b.seg = marshal.SegmentTable(np.asarray(offs, np.int64), np.concatenate(rels).astype(np.int32),
                                 np.concatenate(gens).astype(np.int32), np.concatenate(steps).astype(np.uint8))  # fmt: skip
    b.has_posmap = True
    b.lens = np.array([len(t) for t in texts], np.int32)
    b.export_text = lambda i: texts[i]
    return b


class FakeResult:
    def __init__(self, table, params):
        self.tab, self.params, self.handle = table, params, object()
        self.n_guides = len(table["hap"])
        self.window = params.pam_len + params.guide_len + 20
        self.text_stride = (self.window + 15) // 16 * 16
        self.n_hits = (self.n_guides, 0)
        self.closed = False

    def device_bytes(self):
        return self.n_guides * (21 + self.text_stride)

    def table(self, buffers=None, want_text=True):
        out = {k: v for k, v in self.tab.items() if want_text or k != "text"}
        out["bucket"] = out["bucket"].astype(np.uint32)
        return out

    def annotate(self, batch, **kw):
        return hostcheck.annotate_flat(self.tab, self.params, batch.seg, batch.vt)

    def collapse(self, is_ref):
        """hawk_result_collapse: rows ordered by (start, stop, group), ties in emission order."""
        t, core = self.tab, slice(10, 10 + self.window - 20)
        keys = [(int(t["start"][i]), int(t["stop"][i]), int(t["strand"][i]), int(is_ref[t["hap"][i]]), t["text"][i, core].tobytes())
                for i in range(self.n_guides)]  # fmt: skip
        perm = np.array(sorted(range(self.n_guides), key=lambda i: (keys[i], i)), dtype=np.uint32)
        head = np.array([k == 0 or keys[perm[k]] != keys[perm[k - 1]] for k in range(self.n_guides)], dtype=np.uint8)
        return perm, head, False

    def cfdon(self, is_ref, mm, pam2):
        return hostcheck.cfdon_flat(self.tab, self.params, is_ref, mm, pam2)

    def featurize(self, lead=4, kmers=True, onehot=False, onehot_device_ptr=0, kmers_out=None):
        return hostcheck.featurize_flat(self.tab, self.params, lead, kmers, onehot)

    def close(self):
        self.closed, self.handle = True, None


def _search(ctx, batch, params, a, b, is_ref):
    va = batch.va
    if va is None:
        z = np.zeros(batch.n_hap + 1, np.int64)
        va = type("VA", (), dict(va_off=z, va_idx=np.zeros(1, np.int32), va_ent_off=np.zeros(1, np.int64), va_ref=np.zeros(1, np.uint8)))()
    try:
        table = hostcheck.search_flat(batch.texts, params, a, b, is_ref, batch.seg, va)
    except KeyError as e:
        raise _cabi.HawkLibraryError(str(e), _cabi.HAWK_EALLELES) from e
    return FakeResult(table, params)


def activate(monkeypatch):
    """Route the product's device calls to the CPU emulation for the duration of a test."""
    monkeypatch.setattr(encoder, "pack_texts", lambda texts, debug, ctx=None: FakeBatch(texts))
    monkeypatch.setattr(_cabi, "search", _search)
    monkeypatch.setattr(_cabi.Batch, "from_edits", staticmethod(fake_from_edits))
    monkeypatch.setattr(_cabi.Context, "default", classmethod(lambda cls, device=None: FakeContext()))
    monkeypatch.setattr(_cabi, "load_library", lambda path=None: None)
    monkeypatch.setattr(search_guides, "LIVE_TABLES", search_guides._LiveTables(cap_bytes=1 << 40))
