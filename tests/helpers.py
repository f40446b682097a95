"""Test support: golden-fixture loading and duck-typed stand-ins for the
reference's `Region` / `Haplotype` (haplotype.py:23-77, region.py:15) so the
product's `search()` can be driven where the reference is not installed."""

from __future__ import annotations

import gzip
import json
import os
from typing import List

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class FxSequence:
    def __init__(self, text: str):
        self._sequence = text
        self._sequence_raw = list(text)

    @property
    def sequence(self) -> str:
        return self._sequence

    def __len__(self):
        return len(self._sequence)

    def __getitem__(self, idx):
        return self._sequence_raw[idx]


class FxCoordinate:
    def __init__(self, contig, start, stop):
        self.contig, self.start, self.stop = contig, start, stop

    def __str__(self):
        return f"{self.contig}:{self.start}-{self.stop}"


class FxRegion:
    def __init__(self, contig: str, start: int, stop: int):
        self.contig, self.start, self.stop = contig, start, stop
        self.coordinates = FxCoordinate(contig, start, stop)


class FxHap:
    def __init__(self, d: dict):
        self.sequence = FxSequence(d["seq"])
        pm = d["posmap"]
        self.posmap = dict(enumerate(pm))
        self.posmap_rev = {g: i for i, g in enumerate(pm)}
        self.start, self.stop = d["start"], d["stop"]
        self.samples, self.variants = d["samples"], d["variants"]
        self.afs = {k: (float("nan") if v is None else v) for k, v in d["afs"].items()}
        self.variant_alleles = {
            int(k): [tuple(e) for e in v] for k, v in d["variant_alleles"].items()
        }
        self.id = d["id"]

    def __len__(self):
        return len(self.sequence)

    def __getitem__(self, idx):
        return self.sequence[idx]


def load_golden(name: str) -> List[dict]:
    with gzip.open(os.path.join(GOLDEN_DIR, name + ".json.gz"), "rb") as fh:
        return json.loads(fh.read().decode())


def all_golden_cases() -> List[dict]:
    out = []
    for name in ("kat", "random", "edge", "config1", "config4"):
        out.extend(load_golden(name))
    return out


def golden_guides(case: dict):
    """Expected guides of a fixture as comparable tuples (same shape as OracleGuide)."""
    from oracle.hawk_oracle import OracleGuide

    G, P = case["guidelen"], len(case["pam"])
    return [
        OracleGuide(g[0], g[1], g[2], g[3], g[4], g[5], g[6], g[7], tuple(g[8]), G, P)
        for g in case["guides"]
    ]


def fixture_objects(case: dict):
    region = FxRegion(case["contig"], case["region_start"], case["region_stop"])
    haps = [FxHap(h) for h in case["haps"]]
    return region, haps


def table_to_tuples(table, haps, guidelen, pamlen, right):
    """Guide table (hawk_result_fetch layout) -> OracleGuide tuples in final order.

    Same host steps as crispr_hawk_b200.search_guides.search, minus object creation."""
    import numpy as np

    from oracle.hawk_oracle import OracleGuide

    order = np.argsort(table["bucket"], kind="stable")
    out = []
    for i in order.tolist():
        h = haps[int(table["hap"][i])]
        s = int(table["strand"][i])
        rp = (not right) if s == 1 else bool(right)
        pivot = int(table["pos"][i]) - (0 if rp else guidelen)
        gpm = tuple(h.posmap[pivot + j] for j in range(guidelen + pamlen))
        out.append(
            OracleGuide(int(table["start"][i]), int(table["stop"][i]), s,
                        table["text"][i].tobytes().decode("ascii"), rp, h.samples, h.variants, h.id,
                        gpm, guidelen, pamlen)
        )  # fmt: skip
    return out


def split_hits(recs, n_hap):
    import numpy as np

    hap = (recs >> np.uint64(32)).astype(np.int64)
    pos = (recs & np.uint64(0xFFFFFFFF)).astype(np.int64)
    cuts = np.searchsorted(hap, np.arange(n_hap + 1))
    return [pos[cuts[h] : cuts[h + 1]].tolist() for h in range(n_hap)]
