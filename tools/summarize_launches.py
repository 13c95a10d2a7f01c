"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=25):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void\s+", "", name)
        rows.append((name, v * scale))
    tot = sum(t for _, t in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, t in rows:
        agg[n][0] += 1
        agg[n][1] += t
    print(f"launches={len(rows)} total={tot/1e3:.3f} ms (serialised, cold cache)")
    print(f"{'kernel':70s} {'n':>5s} {'us':>10s} {'share':>7s}")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{n[:70]:70s} {c:5d} {t:10.1f} {100*t/tot:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
