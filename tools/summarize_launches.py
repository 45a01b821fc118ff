#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

  python tools/summarize_launches.py gpurun_out/launches.csv [skip_first_n_launches] > profiles/<name>.md

The per-launch times of such a pass are cold-cache and serialised; what is judged is each kernel's
SHARE of the step, not the absolute (B200_PROFILING.md).
"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("vb::", "").replace("(anonymous namespace)::", "")
    return name[:110]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        rows.append((short(r["Kernel Name"]), float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
    rows = rows[skip:]
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns, _, _ in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    total = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if not k.startswith("at::") and "nccl" not in k.lower())
    print(f"source: `{path}` (launches skipped at the head: {skip})\n")
    print(f"launches: {len(rows)}, summed kernel time: {total / 1e6:.3f} ms, "
          f"hand-written kernels' share: {100 * ours / max(total, 1):.1f} %\n")
    print("| kernel | launches | total ms | share % | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {ns / 1e6:.3f} | {100 * ns / total:.2f} | {ns / c / 1e3:.1f} |")


if __name__ == "__main__":
    main()
