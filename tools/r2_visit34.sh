#!/bin/bash
# round-2 visit 34: sparse tile kernel, one vs two tap groups per CTA: parity, stage timing, launch list, step timing
O=gpurun_out/r02am; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spatial_volume or denoise_step" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for v in 0 1 2; do MD_SPARSE_TILE=$v timeout 200 python tools/time_volume.py 16; done
timeout 200 python tools/time_volume.py 16 body
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sparse|smpl|resample" -c 12 --csv --log-file $O/launches.csv python tools/time_volume.py 16 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02am/launches.csv")) if len(r) > 5 and r[0].isdigit()]
for r in rows[:12]: print(r[4][:66], r[-1])
PY
for v in 0 2; do MD_SPARSE_TILE=$v timeout 200 python tools/time_step.py 2 4 16 > $O/time_step_$v.log 2>&1; echo "tile=$v"; cat $O/time_step_$v.log; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sparse_conv_tile" -c 7 -f -o $O/sparse_tile python tools/time_volume.py 16 > $O/ncu_full.log 2>&1; tail -1 $O/ncu_full.log
