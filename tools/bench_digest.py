"""Print the headline numbers of bench lines: python tools/bench_digest.py file.json [...]"""
import json
import sys

for path in sys.argv[1:]:
    try:
        lines = [ln for ln in open(path).read().splitlines() if ln.startswith("{")]
        d = json.loads(lines[-1])
    except Exception as e:  # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    e2e = d.get("e2e") or {}
    print(f"== {path}: impl={d.get('impl')} n_gpus={d.get('n_gpus')} {d['ms_per_step']:.3f} ms/step value {d['value']:.4g} {d['unit']}"
          f" e2e {e2e.get('ms_per_step') or 0:.2f} ms ({e2e.get('value') or 0:.4g}) clocks {(d.get('clocks') or {}).get('reasons')}")
    r = d.get("roofline")
    if r:
        print("   roofline", r["kernel"], f"{r['achieved']:.0f} GB/s", {k: round(v, 3) for k, v in r.items() if k.endswith("frac") and v is not None})
    k = d.get("kernels")
    if k:
        print("   kernels ms", {n: round(v["ms"], 3) for n, v in k.items()}, "launches", d.get("gpu_launches"))
    fe = e2e.get("from_edit_lists")
    if fe:
        print(f"   from edit lists {fe['ms_per_step']:.2f} ms, rows only {fe.get('rows_only', {}).get('ms_per_step', 0):.2f} ms")
    cb = d.get("cpu_baseline")
    if cb:
        print(f"   cpu {cb['value']:.4g} {cb['unit']} on {cb['cores']} threads ({cb['kind']})")
    nr = d.get("next_rows") or {}
    for key in ("N2_annotate", "N2_collapse", "N4_cfdon", "N4_features"):
        if key in nr:
            print("  ", key, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in nr[key].items() if a.endswith("ms")})
    if "api_search" in nr:
        for name, v in nr["api_search"].items():
            if isinstance(v, dict):
                print("   api_search", name, {a: round(b, 1) for a, b in v.items() if a.endswith("_ms")})
    if "from_variant_records" in nr:
        v = nr["from_variant_records"]
        print("   from_variant_records", round(v["haplotypes_built_ms"], 1), round(v["build_plus_search_ms"], 1))
    for key in ("final_merge", "c5"):
        if d.get(key):
            print("  ", key, json.dumps(d[key])[:600])
