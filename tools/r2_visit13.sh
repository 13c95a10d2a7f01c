#!/bin/bash
# round-2 visit 13 (2 GPUs): sharded step vs the reference golden, bench at N=2
O=gpurun_out/r02n; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "multi_gpu" > $O/pytest_mgpu.log 2>&1; tail -3 $O/pytest_mgpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 50 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 1500 $O/bench_n2.json | head -c 600; echo
timeout 600 python bench.py --no-cpu --no-eager --no-kernels > $O/bench_n1.json 2> $O/bench_n1.err
python - <<PY
import json
for n in ("n1","n2"):
    try:
        d=json.loads([l for l in open("$O/bench_%s.json"%n) if l.startswith("{")][-1]); print(n, "%.2f steps/s %.3f ms e2e %.2f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]))
    except Exception as e: print(n, "failed", e)
PY
