#!/bin/bash
# round-2 visit 32: sparse tile kernel with cp.async weight staging: parity, stage timing, launch list, full capture
O=gpurun_out/r02ak; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spatial_volume" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for v in 0 1; do MD_SPARSE_TILE=$v timeout 200 python tools/time_volume.py 16; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sparse|smpl|resample" -c 12 --csv --log-file $O/launches.csv python tools/time_volume.py 16 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02ak/launches.csv")) if len(r) > 5 and r[0].isdigit()]
for r in rows[:12]: print(r[4][:60], r[-1])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sparse_conv_tile" -c 6 -f -o $O/sparse_tile python tools/time_volume.py 16 > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log
