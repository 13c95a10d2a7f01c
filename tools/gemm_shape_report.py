"""Joins an MD_TRACE log of conv_gemm launches with an ncu launch list: time and TFLOP/s per GEMM shape."""
import collections
import csv
import re
import sys


def main(csv_path, log_path, top=30):
    lines = [l for l in open(csv_path) if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", ""), float(r["Metric Value"].replace(",", "")) / 1e3))
    gem = [t for n, t in rows if n.startswith("conv_gemm")]
    tr = [l.strip() for l in open(log_path) if l.startswith("conv_gemm")][-len(gem):]
    agg = collections.OrderedDict()
    for l, t in zip(tr, gem):
        d = dict(kv.split("=") for kv in l.split()[1:])
        fl = 2.0 * int(d["B"]) * int(d["D"]) * int(d["H"]) * int(d["W"]) * int(d["Cin"]) * int(d["taps"]) * int(d["N"])
        key = (d["B"], d["D"], d["H"], d["W"], d["Cin"], d["taps"], d["N"], d["BN"], d["tiles"], d["act"], d.get("f32"), d.get("bf16"), d.get("res"))
        a = agg.setdefault(key, [0, 0.0, fl])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    print(f"{len(gem)} conv_gemm launches, {tot:.0f} us total")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"B={k[0]:>2} D={k[1]:>2} H={k[2]:>2} W={k[3]:>5} Cin={k[4]:>5} taps={k[5]:>2} N={k[6]:>5} BN={k[7]:>3} tiles={k[8]:>9} "
              f"act={k[9]} f32={k[10]} bf16={k[11]} res={k[12]} n={a[0]:2d} us={a[1]/a[0]:7.1f} tot={a[1]:6.0f} TF/s={a[2]/(a[1]/a[0])/1e6:6.0f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
