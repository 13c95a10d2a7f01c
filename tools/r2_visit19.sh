#!/bin/bash
# round-2 visit 19: GroupNorm tail inside the producing GEMM (A/B against MD_GN_TAIL=0): op test, parity, bench, views sweep
O=gpurun_out/r02t; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "group_norm_tail" > $O/pytest_tail.log 2>&1; tail -5 $O/pytest_tail.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_vae.py -m gpu -x -q -k "unet_forward or denoise_step or trajectory or graph or vae_decode or vae_encode or chunk" > $O/pytest_parity.log 2>&1; tail -4 $O/pytest_parity.log
MD_GN_TAIL=0 timeout 300 python bench.py --no-cpu --no-eager --no-kernels --no-vae > $O/bench_off.json 2> $O/bench_off.err
timeout 300 python bench.py --no-cpu --no-eager --no-kernels --no-vae > $O/bench_on.json 2> $O/bench_on.err
python - <<PY
import json
for n in ("off","on"):
    try:
        d=json.loads(open("$O/bench_%s.json"%n).read()); print(n, "%.2f steps/s %.3f ms e2e %.2f launches %d"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
    except Exception as e: print(n, "failed", e, open("$O/bench_%s.err"%n).read()[-400:])
PY
timeout 200 python tools/time_step.py 2 4 8 16 > $O/time_step.log 2>&1; cat $O/time_step.log
