#!/bin/bash
# round-2 visit 44 (2 GPUs): sharded sampling through the drop-in API against single-GPU sampling
O=gpurun_out/r02aw; mkdir -p $O
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29761 tools/mgpu_sample_check.py $O/mgpu_sample.json > $O/mgpu_sample.log 2>&1
grep MGPU_SAMPLE $O/mgpu_sample.log || tail -30 $O/mgpu_sample.log
