#!/bin/bash
# round-2 visit 2: full GPU test suite with the new configs + default bench
O=gpurun_out/r02b; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_configs.py > $O/pytest_old.log 2>&1; tail -3 $O/pytest_old.log
timeout 1500 python -m pytest tests/test_gpu_configs.py -m gpu -q -s > $O/pytest_new.log 2>&1; tail -25 $O/pytest_new.log
MD_ENC_SINGLE=1 timeout 300 python bench.py --no-cpu --no-eager --no-kernels > $O/bench_encsingle.json 2> $O/bench_encsingle.err
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
python - <<PY
import json
for n in ("bench_encsingle","bench"):
    try:
        d=json.loads(open("$O/%s.json"%n).read()); print(n, "%.2f steps/s %.3f ms e2e %.2f launches %d"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
        if "library_baseline" in d: print(json.dumps(d["library_baseline"])); print(json.dumps(d.get("cpu_baseline")))
    except Exception as e: print(n, "failed", e)
PY
