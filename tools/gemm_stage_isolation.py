"""Which pipeline stage bounds each conv_gemm shape of the step: re-times tools/gemm_suite.py with MD_GEMM_DBG switching
off the epilogue work (1), the MMAs (2), the TMA loads (4) and their combinations (results are garbage in those modes,
only the timings mean something).  Prints cold microseconds per shape and mode."""
import os
import re
import subprocess
import sys

top = sys.argv[1] if len(sys.argv) > 1 else "34"
modes = [("full", 0), ("no_epi", 1), ("no_mma", 2), ("no_tma", 4), ("tma_only", 3), ("mma_only", 5), ("epi_only", 6), ("empty", 7)]
cols = {}
shapes = []
for name, m in modes:
    env = dict(os.environ, MD_GEMM_DBG=str(m))
    out = subprocess.run([sys.executable, "tools/gemm_suite.py", "--top", top], env=env, capture_output=True, text=True).stdout
    rows = [l for l in out.splitlines() if l.startswith("B=")]
    if not shapes:
        shapes = [l[:75] for l in rows]
    cols[name] = [float(re.search(r"cold\s+([\d.]+)", l).group(1)) for l in rows]
    tot = re.search(r"TOTAL per step: cold ([\d.]+)", out)
    print(f"mode {name:9s}: total cold {tot.group(1) if tot else '?'} ms", flush=True)
print(f"{'shape':75s} " + " ".join(f"{n:>8s}" for n, _ in modes))
for i, s in enumerate(shapes):
    print(s + " " + " ".join(f"{cols[n][i]:8.1f}" for n, _ in modes))
