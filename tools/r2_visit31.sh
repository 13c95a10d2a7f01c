#!/bin/bash
# round-2 visit 31: register-tiled sparse convolution (A/B via MD_SPARSE_TILE): parity, stage timing, launch list, step timing
O=gpurun_out/r02aj; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spatial_volume or denoise_step" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for v in 0 1; do MD_SPARSE_TILE=$v timeout 200 python tools/time_volume.py 16; MD_SPARSE_TILE=$v timeout 200 python tools/time_volume.py 16 body; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sparse|smpl|resample" --csv --log-file $O/launches.csv python tools/time_volume.py 16 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02aj/launches.csv")) if len(r) > 5 and r[0].isdigit()]
for r in rows[:12]: print(r[4][:60], r[-1])
PY
for v in 0 1; do MD_SPARSE_TILE=$v timeout 200 python tools/time_step.py 2 16 > $O/time_step_$v.log 2>&1; echo "tile=$v"; cat $O/time_step_$v.log; done
