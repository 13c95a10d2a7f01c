#!/bin/bash
# round-2 visit 40: minimum K blocks per split (MD_SPLIT_DIV 4 vs 8 vs 12) on the step at 2 / 4 / 16 views
O=gpurun_out/r02as; mkdir -p $O
for d in 4 8 12; do echo "split_div=$d"; MD_SPLIT_DIV=$d timeout 300 python tools/time_step.py 2 4 16 --steps=40 2>&1 | tee $O/time_step_div$d.log; done
