#!/bin/bash
# round-2 visit 35: launch lists of one step at 16 views (refresh after the sparse-conv work) and at 2 views (what a rank of the 8-GPU job runs)
O=gpurun_out/r02an; mkdir -p $O
for n in 16 2; do
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_n$n.csv python tools/profile_step.py $n > $O/profile_n$n.log 2>&1
  python tools/summarize_launches.py $O/launches_n$n.csv 70 > $O/summary_n$n.txt; head -3 $O/summary_n$n.txt
done
