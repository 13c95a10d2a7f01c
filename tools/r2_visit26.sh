#!/bin/bash
# round-2 visit 26: source-level stall profile of the level-0 attention launch
O=gpurun_out/r02af; mkdir -p $O
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -o $O/att -f python tools/attn_one.py > $O/att.log 2>&1
ncu -i $O/att.ncu-rep --page source --csv 2>> $O/att.log | python tools/ncu_source_top.py 60 > $O/att_source_top.txt 2>&1
ncu -i $O/att.ncu-rep --page raw --csv 2>> $O/att.log > $O/att_raw.csv
python - <<PY
import csv
rows=list(csv.reader(open("$O/att_raw.csv")))
h=rows[0]; r=rows[2]
for i,k in enumerate(h):
    if any(t in k for t in ("warp_issue_stalled","pipe_xu","pipe_tensor","issue_active","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","smsp__cycles_active.avg","sm__cycles_elapsed.max","tmem","pipe_alu","pipe_fma","inst_executed.sum")) and r[i] not in ("","0"):
        print(k, r[i], rows[1][i])
PY
head -80 $O/att_source_top.txt
rm -f $O/att.ncu-rep
