#!/bin/bash
# round-2 visit 23 (8 GPUs): strong scaling at N = 8, view sweep (BASELINE config 5) and the SMPL-X config at N = 8
O=gpurun_out/r02x; mkdir -p $O
run() { # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 8 --steps 30 --warmup 3 $2 > $O/bench_$1.json 2> $O/bench_$1.err
}
run n8_v16 "" 29541
run n8_v64 "--views 64" 29542
run n8_v32 "--views 32" 29543
run n8_v8 "--views 8" 29544
run n8_smplx "--config smplx" 29545
python - <<PY
import json
for n in ("n8_v16","n8_v64","n8_v32","n8_v8","n8_smplx"):
    try:
        d=json.loads([l for l in open("$O/bench_%s.json"%n) if l.startswith("{")][-1]); print(n, d["metric"], "%.2f steps/s %.3f ms e2e %.2f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), d["config"]["workload"][:40])
    except Exception as e: print(n, "failed", e, open("$O/bench_%s.err"%n).read()[-600:])
PY
