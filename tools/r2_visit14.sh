#!/bin/bash
# round-2 visit 14: evidence run — launch list of the current step, ncu --set full of the three roofline convolution
# shapes and the level-0 attention, full default bench (with the torch-eager, CPU and VAE blocks), reference arm
O=gpurun_out/r02o; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $O/launches.csv python tools/profile_step.py 16 > $O/launches.log 2>&1
python tools/summarize_launches.py $O/launches.csv 70 > $O/launches_summary.txt 2>&1; head -12 $O/launches_summary.txt
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -o $O/roof -f \
  python tools/gemm_cases.py roof_conv640_l0 roof_conv320_l0 roof_conv1280_l1 geglu_l0 > $O/roof.log 2>&1
ncu -i $O/roof.ncu-rep --page raw --csv > $O/roof_raw.csv 2>> $O/roof.log
python tools/ncu_tensor_summary.py < $O/roof_raw.csv > $O/roof_summary.txt 2>&1; cat $O/roof_summary.txt; grep -h "warm" $O/roof.log
timeout 300 ncu --profile-from-start off --set full --clock-control none -o $O/att -f python tools/attn_one.py > $O/att.log 2>&1
ncu -i $O/att.ncu-rep --page raw --csv > $O/att_raw.csv 2>> $O/att.log
python tools/ncu_tensor_summary.py < $O/att_raw.csv > $O/att_summary.txt 2>&1; cat $O/att_summary.txt
rm -f $O/roof.ncu-rep $O/att.ncu-rep
( time timeout 900 python bench.py > $O/bench.json 2> $O/bench.err ) 2> $O/bench_time.txt; tail -3 $O/bench_time.txt
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err ) 2> $O/bench_ref_time.txt; tail -3 $O/bench_ref_time.txt
python - <<PY
import json
d=json.loads(open("$O/bench.json").read())
print("%.2f steps/s %.3f ms e2e %.2f launches %d"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
print(json.dumps(d["roofline"]["frac_of_burst"]), json.dumps(d.get("vae_decode")), json.dumps(d.get("library_baseline"))[:300], json.dumps(d.get("cpu_baseline")))
for k in d["roofline"].get("kernels", []): print(k["kernel"], k["shape"], "%.1f us frac %.3f"%(k["us"], k["frac"]))
print(open("$O/bench_ref.json").read()[:400])
PY
