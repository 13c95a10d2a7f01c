#!/bin/bash
O=gpurun_out/r02ag; mkdir -p $O
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -o $O/att -f python tools/attn_one.py > $O/att.log 2>&1
ncu -i $O/att.ncu-rep --page source --csv 2>> $O/att.log | python tools/ncu_source_by_opcode.py > $O/att_by_opcode.txt 2>&1
cat $O/att_by_opcode.txt
ncu -i $O/att.ncu-rep --page raw --csv 2>> $O/att.log | python tools/ncu_tensor_summary.py
rm -f $O/att.ncu-rep
