#!/bin/bash
# round-2 visit 10: batch-construction tests, VAE shell test, frustum gather rewrite, ncu evidence for the HBM-bound kernels
O=gpurun_out/r02k; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_vae.py -m gpu -x -q > $O/pytest_new.log 2>&1; tail -5 $O/pytest_new.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "frustum or spatial or reference_shaped or step" > $O/pytest_geo.log 2>&1; tail -3 $O/pytest_geo.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"frustum_gather|vertex_features|volume_resample|depth_attention|gn_apply_fused|layer_norm|sparse_conv|target_encoder|smpl_scatter" -c 60 \
  -o $O/hbm -f python tools/profile_step.py 16 > $O/hbm.log 2>&1
ncu -i $O/hbm.ncu-rep --page raw --csv > $O/hbm_raw.csv 2>> $O/hbm.log
python tools/ncu_hbm_summary.py < $O/hbm_raw.csv > $O/hbm_summary.txt 2>&1; tail -22 $O/hbm_summary.txt
sz=$(stat -c %s $O/hbm.ncu-rep 2>/dev/null || echo 0); if [ "$sz" -gt 30000000 ]; then rm -f $O/hbm.ncu-rep; fi
timeout 300 python bench.py --no-cpu --no-eager --no-kernels > $O/bench.json 2> $O/bench.err
python - <<PY
import json
d=json.loads(open("$O/bench.json").read()); print("%.2f steps/s %.3f ms e2e %.2f launches %d"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
PY
