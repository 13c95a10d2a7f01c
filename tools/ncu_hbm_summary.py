"""Per-kernel memory evidence from `ncu -i rep --page raw --csv` (stdin): duration, DRAM bytes, achieved DRAM GB/s
against the measured copy peak, L1/L2 throughput %, registers.  One line per launch, plus a per-kernel mean.
usage: ncu -i x.ncu-rep --page raw --csv | python tools/ncu_hbm_summary.py [peak_GBs]"""
import csv
import re
import sys
from collections import defaultdict

peak = float(sys.argv[1]) if len(sys.argv) > 1 else 6552.3
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
units = rows[1]
col = {h: i for i, h in enumerate(hdr)}


def get(r, name, default=0.0):
    i = col.get(name)
    if i is None or r[i] == "":
        return default
    try:
        v = float(r[i].replace(",", ""))
    except ValueError:
        return default
    u = units[i]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3,
             "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
    return v * scale


agg = defaultdict(list)
print(f"{'kernel':44s} {'us':>8s} {'dram rd MB':>10s} {'dram wr MB':>10s} {'GB/s':>8s} {'of peak':>7s} {'L1 %':>6s} {'L2 %':>6s} {'regs':>5s}")
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    name = re.sub(r"^void\s+", "", name).replace("md::", "").replace("<unnamed>::", "")
    us = get(r, "gpu__time_duration.sum")
    rd, wr = get(r, "dram__bytes_read.sum"), get(r, "dram__bytes_write.sum")
    l1 = get(r, "l1tex__throughput.avg.pct_of_peak_sustained_active")
    l2 = get(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed")
    regs = get(r, "launch__registers_per_thread")
    gbs = (rd + wr) / us / 1e3 if us else 0.0
    agg[name].append((us, rd, wr, gbs, l1, l2, regs))
    print(f"{name[:44]:44s} {us:8.1f} {rd / 1e6:10.2f} {wr / 1e6:10.2f} {gbs:8.0f} {gbs / peak:7.3f} {l1:6.1f} {l2:6.1f} {regs:5.0f}")
print()
print("per-kernel mean:")
for name, v in agg.items():
    n = len(v)
    m = [sum(x[i] for x in v) / n for i in range(7)]
    print(f"{name[:44]:44s} n={n:2d} {m[0]:8.1f} us  dram {(m[1] + m[2]) / 1e6:8.2f} MB  {m[3]:7.0f} GB/s = {m[3] / peak:5.3f} of {peak:.0f}  L1 {m[4]:5.1f}%  L2 {m[5]:5.1f}%")
