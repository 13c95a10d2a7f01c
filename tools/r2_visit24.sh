#!/bin/bash
# round-2 visit 24: split-row attention (MD_ATT_VARIANT=2): parity of the attention tests, timing vs variants 0 / 1, bench
O=gpurun_out/r02y; mkdir -p $O
MD_ATT_VARIANT=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "attention" > $O/pytest_att.log 2>&1; tail -4 $O/pytest_att.log
for v in 1 2; do MD_ATT_VARIANT=$v timeout 100 python tools/time_attention.py > $O/att$v.log 2>&1; echo "variant $v"; cat $O/att$v.log; done
MD_ATT_VARIANT=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "unet_forward or denoise_step" > $O/pytest_step.log 2>&1; tail -3 $O/pytest_step.log
for v in 1 2; do MD_ATT_VARIANT=$v timeout 300 python bench.py --no-cpu --no-eager --no-kernels --no-vae > $O/bench_v$v.json 2> $O/bench_v$v.err; done
python - <<PY
import json
for n in ("v1","v2"):
    try:
        d=json.loads(open("$O/bench_%s.json"%n).read()); print(n, "%.2f steps/s %.3f ms e2e %.2f launches %d"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
    except Exception as e: print(n, "failed", e, open("$O/bench_%s.err"%n).read()[-400:])
PY
