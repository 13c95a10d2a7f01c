#!/bin/bash
# round-2 visit 20 (4 GPUs): strong scaling lines at N = 1, 2, 4 with the round-2 library; sharded parity test
O=gpurun_out/r02u; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "multi_gpu" > $O/pytest_mgpu.log 2>&1; tail -2 $O/pytest_mgpu.log
for N in 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 50 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
done
timeout 600 python bench.py --no-cpu --no-eager --no-kernels --no-vae > $O/bench_n1.json 2> $O/bench_n1.err
python - <<PY
import json
for n in ("n1","n2","n4"):
    try:
        d=json.loads([l for l in open("$O/bench_%s.json"%n) if l.startswith("{")][-1]); print(n, "%.2f steps/s %.3f ms e2e %.2f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]))
    except Exception as e: print(n, "failed", e, open("$O/bench_%s.err"%n).read()[-500:])
PY
