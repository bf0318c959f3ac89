#!/usr/bin/env python
"""ncu `--metrics gpu__time_duration.sum --csv` launch list -> per-kernel markdown summary (count, total, mean, share)."""
import collections
import csv
import re
import sys


def main(path: str, title: str = "") -> None:
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("mb200::", "").replace("unnamed>::", "").replace("<", "<")
        v = float(row["Metric Value"]) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"### {title or path}\n")
    print("| kernel | launches | total ms | mean us | share of listed time |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1] / 1e6:.2f} | {v[1] / v[0] / 1e3:.1f} | {100 * v[1] / tot:.1f} % |")
    print(f"\nlisted launches: {sum(v[0] for v in agg.values())}, listed device time: {tot / 1e6:.1f} ms "
          "(ncu serialises launches and runs them cold: compare shares, not absolutes)\n")


if __name__ == "__main__":
    main(*sys.argv[1:])
