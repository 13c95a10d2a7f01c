#!/bin/bash
# round-2 visit 4: bf16 GroupNorm-only intermediates, row-prefetching GN apply, conv_in on tcgen05, no cast passes
O=gpurun_out/r02d; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 300 python bench.py --no-cpu --no-eager > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
python - <<PY
import json
d=json.loads(open("$O/bench.json").read()); print("%.2f steps/s %.3f ms e2e %.2f launches %d"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
for k in d["roofline"]["kernels"]: print(k["kernel"], k["shape"], "%.1f us frac %.3f"%(k["us"], k["frac"]))
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $O/launches.csv python tools/profile_step.py 16 > $O/launches.log 2>&1
python tools/summarize_launches.py $O/launches.csv 60 > $O/launches_summary.txt 2>&1
head -40 $O/launches_summary.txt
