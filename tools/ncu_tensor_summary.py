"""Tensor-pipe evidence from `ncu -i rep --page raw --csv` (stdin): per launch duration, tensor-pipe active %, issue
active %, L2 throughput %, DRAM bytes, registers.  usage: ncu -i x.ncu-rep --page raw --csv | python tools/ncu_tensor_summary.py"""
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
WANT = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("dram__bytes_read.sum", "dram rd"),
        ("dram__bytes_write.sum", "dram wr"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    name = re.sub(r"^void\s+", "", name).replace("md::", "")
    parts = []
    for key, label in WANT:
        i = col.get(key)
        if i is None:
            continue
        parts.append(f"{label} {r[i]} {units[i]}".strip())
    print(name[:60], "|", "; ".join(parts))
