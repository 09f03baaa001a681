"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel (shares, not absolutes)."""
import collections
import csv
import re
import sys


def summarize(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"us": 1e3, "ms": 1e6, "ns": 1, "s": 1e9}.get(row["Metric Unit"], 1)
        k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void unnamed>::", "").replace("unnamed>::", "")
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | total ms | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {v[0]} | {v[1]/1e6:.3f} | {v[1]/v[0]/1e3:.1f} | {100*v[1]/tot:.1f}% |")
    out.append(f"\nTotal {tot/1e6:.2f} ms")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(sys.argv[1]))
