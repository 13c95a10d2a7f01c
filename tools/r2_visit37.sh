#!/bin/bash
# round-2 visit 37 (8 GPUs): 16 views over 8 ranks with the NVLink peer exchange and with the NCCL all-reduce, then 4 ranks
O=gpurun_out/r02ap; mkdir -p $O
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
show() { python - <<PY
import json
l=[x for x in open("$1") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("$1", d["n_gpus"], "gpus", round(d["value"],2), "steps/s", round(d["ms_per_step"],3), "ms e2e", round(d["e2e"]["value"],2), d["config"].get("exchange","")[:24], d["clocks"]["reasons"])
else:
    print("$1 no line"); print(open("$1".replace(".json",".err")).read()[-1500:])
PY
}
F="--no-cpu --no-eager --no-kernels --no-vae"
MD_PEER=1 run 8 29731 bench.py --gpus 8 --steps 50 --warmup 5 $F > $O/bench_n8_peer.json 2> $O/bench_n8_peer.err; show $O/bench_n8_peer.json
MD_PEER=0 run 8 29732 bench.py --gpus 8 --steps 30 --warmup 5 $F > $O/bench_n8_nccl.json 2> $O/bench_n8_nccl.err; show $O/bench_n8_nccl.json
MD_PEER=1 run 4 29733 bench.py --gpus 4 --steps 30 --warmup 5 $F > $O/bench_n4_peer.json 2> $O/bench_n4_peer.err; show $O/bench_n4_peer.json
