#!/bin/bash
# round-2 visit 1: warp-uniform GEMM issue loops vs the round-1 library (A/B in one box)
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "gemm or conv or split or stats or step or unet" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
MD_BUILD_TAG=old timeout 200 python tools/gemm_suite.py > $O/suite_old.log 2>&1; tail -1 $O/suite_old.log
timeout 200 python tools/gemm_suite.py > $O/suite_new.log 2>&1; tail -1 $O/suite_new.log
MD_CG2=2 timeout 200 python tools/gemm_suite.py > $O/suite_new_cg2.log 2>&1; tail -1 $O/suite_new_cg2.log
MD_BUILD_TAG=kprof timeout 200 python tools/kprof_batch.py > $O/kprof.log 2>&1
MD_BUILD_TAG=old timeout 200 python bench.py --no-cpu > $O/bench_old.json 2> $O/bench_old.err
timeout 200 python bench.py --no-cpu > $O/bench_new.json 2> $O/bench_new.err
MD_CG2=1 timeout 200 python bench.py --no-cpu > $O/bench_new_cg2.json 2> $O/bench_new_cg2.err
python - <<PY
import json
for n in ("old","new","new_cg2"):
    try:
        d=json.loads(open("$O/bench_%s.json"%n).read()); print(n, "%.2f steps/s %.3f ms"%(d["value"], d["ms_per_step"]))
    except Exception as e: print(n, "failed", e)
PY
