#!/bin/bash
# round-2 visit 11: LN3 without write-back (v2 through the ff2 epilogue), attention variant default; full suite + bench
O=gpurun_out/r02l; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 300 python bench.py --no-cpu --no-eager --no-kernels > $O/bench.json 2> $O/bench.err
python - <<PY
import json
d=json.loads(open("$O/bench.json").read()); print("%.2f steps/s %.3f ms e2e %.2f launches %d"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
PY
timeout 200 python tools/time_step.py 2 4 8 16 > $O/time_step.log 2>&1; cat $O/time_step.log
