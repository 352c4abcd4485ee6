#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import csv
import re
import sys
from collections import defaultdict


def main(path, out=None):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "nsecond": 1, "msecond": 1e6, "second": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, ns))
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    tot = sum(v[1] for v in agg.values()) or 1.0
    lines = [f"# launch list summary of {path} ({len(rows)} launches, {tot/1e6:.3f} ms total under ncu, cold-cache/serialised)",
             "", "| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{n}` | {c} | {ns/1e6:.3f} | {100*ns/tot:.1f}% | {ns/c/1e3:.1f} |")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
