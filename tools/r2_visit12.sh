#!/bin/bash
# round-2 visit 12: weight tiles requested before the grid dependency resolves; tests, bench, views sweep, launch list at 2 views
O=gpurun_out/r02m; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemm or conv or split or stats" > $O/pytest_gemm.log 2>&1; tail -3 $O/pytest_gemm.log
timeout 300 python bench.py --no-cpu --no-eager --no-kernels > $O/bench.json 2> $O/bench.err
python - <<PY
import json
d=json.loads(open("$O/bench.json").read()); print("%.2f steps/s %.3f ms e2e %.2f launches %d"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
PY
timeout 200 python tools/time_step.py 2 4 8 16 > $O/time_step.log 2>&1; cat $O/time_step.log
timeout 200 python tools/gemm_suite.py > $O/suite.log 2>&1; tail -1 $O/suite.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $O/launches_v2.csv python tools/profile_step.py 2 > $O/launches_v2.log 2>&1
python tools/summarize_launches.py $O/launches_v2.csv 30 > $O/launches_v2_summary.txt 2>&1; head -34 $O/launches_v2_summary.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
