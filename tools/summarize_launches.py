"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot = collections.defaultdict(float)
cnt = collections.Counter()
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void |lrcn::", "", name)
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(f"| kernel | launches | total µs | share |\n|---|---:|---:|---:|")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"| `{k}` | {cnt[k]} | {v:.1f} | {100 * v / T:.1f}% |")
print(f"| **sum** | {sum(cnt.values())} | {T:.1f} | 100% |")
