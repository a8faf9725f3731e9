"""Profiling aid: reduce an `ncu --metrics gpu__time_duration.sum --csv` launch list to one row per kernel
(launches, total, average, share of the captured device time).  Usage: summarise_launches.py launches.csv "header note" > summary.csv"""
import csv
import re
import sys
from collections import OrderedDict


def main(path: str, note: str) -> None:
    rows = [l for l in open(path, newline="") if l.startswith('"')]
    per = OrderedDict()
    for r in csv.DictReader(rows):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*$", "", r["Kernel Name"])[:80]
        ns = float(r["Metric Value"].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1.0)
        n, t = per.get(name, (0, 0.0))
        per[name] = (n + 1, t + ns)
    total = sum(t for _, t in per.values())
    count = sum(n for n, _ in per.values())
    print(f"# {note}")
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES, not absolute times)")
    print(f"# total captured device time: {total / 1e6:.3f} ms over {count} launches")
    print("kernel,launches,total_us,avg_us,share")
    for name, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"{name},{n},{t / 1e3:.1f},{t / 1e3 / n:.2f},{t / total:.4f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu launch list summary")
