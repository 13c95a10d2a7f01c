#!/bin/bash
# round-2 visit 39: exhaustive GEMM configuration sweep at 2, 4 and 16 views
O=gpurun_out/r02ar; mkdir -p $O
for n in 2 4 16; do timeout 500 python tools/gemm_autotune.py $n > $O/autotune_n$n.log 2>&1; grep -A12 "^TOTAL" $O/autotune_n$n.log | head -14; done
