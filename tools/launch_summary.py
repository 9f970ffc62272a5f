"""Summarise an ncu launch list (gpu__time_duration.sum CSV): per-kernel share of the LAST step.

    python tools/launch_summary.py gpurun_out/launches.csv [first_id last_id]
"""
import csv
import re
import sys
from collections import defaultdict


def rows(path):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
            yield int(r["ID"]), r["Kernel Name"], us


def main():
    rs = list(rows(sys.argv[1]))
    if len(sys.argv) > 3:
        rs = [r for r in rs if int(sys.argv[2]) <= r[0] <= int(sys.argv[3])]
    agg, cnt = defaultdict(float), defaultdict(int)
    for _, k, us in rs:
        k = re.sub(r"\(.*", "", k)[:100]
        agg[k] += us
        cnt[k] += 1
    tot = sum(agg.values())
    print(f"# {len(rs)} launches, sum {tot / 1e3:.2f} ms")
    print("    sum_us  share  count  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        print(f"{v:10.1f} {100 * v / tot:5.1f}% {cnt[k]:6d}  {k}")


if __name__ == "__main__":
    main()
