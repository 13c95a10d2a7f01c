#!/bin/bash
# round-2 visit 3: chunk-invariance diagnosis at N=64, the re-bind test, launch list of the current step
O=gpurun_out/r02c; mkdir -p $O
timeout 600 python tools/diag_chunks.py 64 16 32 > $O/diag64.log 2>&1; cat $O/diag64.log | tail -5
MD_CG2=0 timeout 600 python tools/diag_chunks.py 32 16 32 > $O/diag32_nocg2.log 2>&1; cat $O/diag32_nocg2.log | tail -5
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "other_step_counts" > $O/pytest_rebind.log 2>&1; tail -5 $O/pytest_rebind.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $O/launches.csv python tools/profile_step.py 16 > $O/launches.log 2>&1
python tools/summarize_launches.py $O/launches.csv 60 > $O/launches_summary.txt 2>&1
head -64 $O/launches_summary.txt
