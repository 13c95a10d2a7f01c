#!/bin/bash
# round-2 visit 36 (2 GPUs): NVLink peer exchange of the vertex-feature sums vs the NCCL all-reduce: parity against the
# reference golden (plain, capture, replay), then bench lines for both
O=gpurun_out/r02ao; mkdir -p $O
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
MD_PEER=1 run 29711 tools/mgpu_check.py $O/mgpu_peer.json > $O/mgpu_peer.log 2>&1; grep MGPU $O/mgpu_peer.log || tail -20 $O/mgpu_peer.log
MD_PEER=0 run 29712 tools/mgpu_check.py $O/mgpu_nccl.json > $O/mgpu_nccl.log 2>&1; grep MGPU $O/mgpu_nccl.log || tail -5 $O/mgpu_nccl.log
for m in 1 0; do
  MD_PEER=$m run $((29720+m)) bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu --no-eager --no-kernels --no-vae > $O/bench_peer$m.json 2> $O/bench_peer$m.err
  python - <<PY
import json
l=[x for x in open("$O/bench_peer$m.json") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("peer=$m", d["value"], "steps/s", d["ms_per_step"], "ms e2e", d["e2e"]["value"], d["config"].get("exchange"), "launches", d["gpu_launches"])
else:
    print("peer=$m no line"); print(open("$O/bench_peer$m.err").read()[-1500:])
PY
done
