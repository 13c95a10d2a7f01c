#!/bin/bash
# round-2 visit 28: pipelined softmax (variant 3) on top of the K / V barrier split
O=gpurun_out/r02ad; mkdir -p $O
MD_ATT_VARIANT=3 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "attention" > $O/pytest_att.log 2>&1; tail -3 $O/pytest_att.log
for v in 1 3; do MD_ATT_VARIANT=$v timeout 100 python tools/time_attention.py > $O/att$v.log 2>&1; echo "variant $v"; cat $O/att$v.log; done
