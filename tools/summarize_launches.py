#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals (profiles/*.md).
usage: python tools/summarize_launches.py gpurun_out/launches.csv [skip_first_n]"""
import csv
import re
import sys
from collections import defaultdict


def main(path, skip=0):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
    rows = rows[skip:]
    tot = sum(ns for _, ns in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for name, ns in rows:
        name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
        name = re.sub(r"\(.*", "", name)
        m = re.match(r"([\w:]+)(<[^>]*>)?", name)
        name = (m.group(1) + (m.group(2) or "")) if m else name
        agg[name][0] += 1
        agg[name][1] += ns
    print(f"# {len(rows)} launches, {tot / 1e6:.2f} ms total device time (cold-cache, serialised: compare SHARES)")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| `{name[:70]}` | {n} | {ns / 1e6:.2f} | {100 * ns / tot:.1f}% | {ns / n / 1e3:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
