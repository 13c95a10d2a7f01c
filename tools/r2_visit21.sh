#!/bin/bash
# round-2 visit 21: split-K only where the K loop shortens by more than the exchange latency (A/B via MD_SPLIT_MIN)
O=gpurun_out/r02v; mkdir -p $O
for T in 0 16 28 48; do
MD_SPLIT_MIN=$T timeout 200 python tools/time_step.py 2 4 8 16 > $O/time_step_$T.log 2>&1; echo "split_min=$T"; cat $O/time_step_$T.log
done
