#!/bin/bash
# round-2 visit 43 (8 GPUs): BASELINE config 5 ends (8 and 64 views) and config 4 (SMPL-X body) with the final code
O=gpurun_out/r02av; mkdir -p $O
F="--no-cpu --no-eager --no-kernels --no-vae"
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 8 --steps 20 --warmup 3 $F $2 > $O/bench_$1.json 2> $O/bench_$1.err; }
run n8_v64 "--views 64" 29742
run n8_smplx "--config smplx" 29745
run n8_v8 "--views 8" 29744
python - <<PY
import json
for n in ("n8_v64","n8_smplx","n8_v8"):
    try:
        d=json.loads([l for l in open("$O/bench_%s.json"%n) if l.startswith("{")][-1]); print(n, "%.2f steps/s %.3f ms e2e %.2f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), d["config"]["workload"][:40], d["config"].get("exchange","")[:12])
    except Exception as e: print(n, "failed", e, open("$O/bench_%s.err"%n).read()[-600:])
PY
