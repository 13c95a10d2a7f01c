#!/bin/bash
# round-2 visit 8: split items first + batched read-back (tests, GEMM suite hybrid on/off, bench on/off)
O=gpurun_out/r02i; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemm or conv or split or stats" > $O/pytest_gemm.log 2>&1; tail -3 $O/pytest_gemm.log
MD_HYBRID=0 timeout 200 python tools/gemm_suite.py > $O/suite_off.log 2>&1; tail -1 $O/suite_off.log
timeout 200 python tools/gemm_suite.py > $O/suite_on.log 2>&1; tail -1 $O/suite_on.log
MD_HYBRID=0 timeout 300 python bench.py --no-cpu --no-eager --no-kernels > $O/bench_off.json 2> $O/bench_off.err
timeout 300 python bench.py --no-cpu --no-eager --no-kernels > $O/bench_on.json 2> $O/bench_on.err
python - <<PY
import json
for n in ("off","on"):
    try:
        d=json.loads(open("$O/bench_%s.json"%n).read()); print(n, "%.2f steps/s %.3f ms e2e %.2f launches %d"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
    except Exception as e: print(n, "failed", e)
PY
