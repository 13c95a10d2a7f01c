#!/bin/bash
# round-2 visit 27: attention with separate K / V stage barriers: parity, timing (variants 0 and 1), bench
O=gpurun_out/r02ac; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "attention" > $O/pytest_att.log 2>&1; tail -4 $O/pytest_att.log
for v in 1 0; do MD_ATT_VARIANT=$v timeout 100 python tools/time_attention.py > $O/att$v.log 2>&1; echo "variant $v"; cat $O/att$v.log; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "unet_forward or denoise_step" > $O/pytest_step.log 2>&1; tail -3 $O/pytest_step.log
timeout 300 python bench.py --no-cpu --no-eager --no-kernels --no-vae > $O/bench.json 2> $O/bench.err
python - <<PY
import json
d=json.loads(open("$O/bench.json").read()); print("%.2f steps/s %.3f ms e2e %.2f launches %d"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
PY
