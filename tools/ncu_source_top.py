"""Condenses `ncu -i rep --page source --csv` (stdin) to the hottest instructions: per kernel launch, the top-N SASS
lines by warp-stall samples with every non-zero numeric column.  usage: ncu -i x.ncu-rep --page source --csv | python tools/ncu_source_top.py [N]"""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 80
rows = list(csv.reader(sys.stdin))
# the source page prints one table per kernel: a header row starts with '#' or contains 'Source'
hdr = None
block = []
blocks = []
for r in rows:
    if any(c.strip() in ("Source", "# Source", "Address") for c in r[:3]) or (r and r[0].strip().startswith("#")):
        if hdr and block:
            blocks.append((hdr, block))
        hdr, block = r, []
    elif hdr:
        block.append(r)
if hdr and block:
    blocks.append((hdr, block))
for bi, (hdr, block) in enumerate(blocks):
    samp = [i for i, h in enumerate(hdr) if "Sampling" in h and "All" in h]
    if not samp:
        samp = [i for i, h in enumerate(hdr) if "Samples" in h]
    if not samp:
        print("no sampling column in", hdr[:12])
        continue
    sc = samp[0]

    def val(r, i):
        try:
            return float(r[i].replace(",", ""))
        except Exception:  # noqa: BLE001
            return 0.0
    tot = sum(val(r, sc) for r in block)
    print(f"=== table {bi}: {len(block)} lines, total samples {tot:.0f}; columns: {hdr}")
    order = sorted(range(len(block)), key=lambda k: -val(block[k], sc))[:top]
    for k in sorted(order):
        r = block[k]
        nz = [f"{hdr[i]}={r[i]}" for i in range(len(r)) if i != sc and i > 1 and val(r, i) != 0.0 and "Sampling" not in hdr[i]
              and ("stall" in hdr[i].lower() or "Stall" in hdr[i])]
        print(f"{k:6d} {val(r, sc):8.0f} {100 * val(r, sc) / max(tot, 1):5.1f}%  {r[1][:90] if len(r) > 1 else ''}  | {r[0][:60]} | " + " ".join(nz[:8]))
