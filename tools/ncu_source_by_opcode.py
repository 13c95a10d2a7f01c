"""Aggregates `ncu -i rep --page source --csv` (stdin) warp-stall samples by SASS opcode (first token of the instruction,
predicate stripped) and prints the share per opcode with the dominant stall reasons.  usage: ... | python tools/ncu_source_by_opcode.py"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(sys.stdin))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Sampling" in c for c in r):
        hdr = r
        rows = rows[i + 1:]
        break
if hdr is None:
    sys.exit("no source table")
col = {h: i for i, h in enumerate(hdr)}
samp = col["Warp Stall Sampling (All Samples)"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = defaultdict(lambda: [0, defaultdict(int)])
tot = 0
for r in rows:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[samp] or 0)
    except ValueError:
        continue
    if not n:
        continue
    ins = re.sub(r"^@!?U?P\w+\s+", "", r[col["Source"]].strip())
    op = ins.split()[0] if ins else "?"
    op = op.split(".")[0] if not op.startswith(("MUFU", "LDTM", "STTM", "UTC", "SYNCS", "FENCE")) else op
    agg[op][0] += n
    tot += n
    for s in stalls:
        try:
            v = int(r[col[s]] or 0)
        except ValueError:
            v = 0
        if v:
            agg[op][1][s] += v
print(f"total samples {tot}")
for op, (n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    top = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print(f"{op:28s} {n:6d} {100 * n / tot:5.1f}%   {top}")
