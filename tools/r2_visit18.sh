#!/bin/bash
# round-2 visit 18: bf16 internal stream of the spatial transformer (A/B against MD_ST_FP32=1): parity + bench
O=gpurun_out/r02s; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -s -k "unet_forward or denoise_step or trajectory or layer_norm" > $O/pytest_parity.log 2>&1; tail -6 $O/pytest_parity.log; grep "drift" $O/pytest_parity.log
MD_ST_FP32=1 timeout 300 python bench.py --no-cpu --no-eager --no-kernels --no-vae > $O/bench_fp32.json 2> $O/bench_fp32.err
timeout 300 python bench.py --no-cpu --no-eager --no-kernels --no-vae > $O/bench_bf16.json 2> $O/bench_bf16.err
python - <<PY
import json
for n in ("fp32","bf16"):
    try:
        d=json.loads(open("$O/bench_%s.json"%n).read()); print(n, "%.2f steps/s %.3f ms e2e %.2f launches %d"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
    except Exception as e: print(n, "failed", e)
PY
timeout 300 python - > $O/eps_err.log 2>&1 <<PY
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
from morphablediffusion_b200 import synth
from morphablediffusion_b200.engine import Engine
gold = np.load("tests/golden/step_n16_persp.npz")
n = 16
sd = synth.make_state_dict()
batch = synth.make_batch(n, "perspective", "flame", int(gold["seed"]))
x_t, x_input, clip = synth.make_inputs(n, 32, int(gold["seed"]))
noise = torch.randn(x_t.shape, generator=torch.Generator().manual_seed(int(gold["noise_seed"])))
eng = Engine(max_views_per_call=16); eng.load_state_dict(sd); eng.bind(batch, "perspective")
x = x_t[0].cuda().contiguous()
eps = eng.denoise_step(x, x_input[0].cuda().contiguous(), clip[0,0].cuda().contiguous(), int(gold["index"]), float(gold["cfg_scale"]), noise=noise[0].cuda().contiguous(), want_eps=True)
torch.cuda.synchronize()
ref = torch.from_numpy(gold["eps"])[0]
print("eps rel-L2 vs reference golden (N=16):", float((eps.cpu()-ref).norm()/ref.norm()), "mode", "fp32" if os.environ.get("MD_ST_FP32") else "bf16")
PY
cat $O/eps_err.log
MD_ST_FP32=1 timeout 300 python - >> $O/eps_err.log 2>&1 <<PY
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
from morphablediffusion_b200 import synth
from morphablediffusion_b200.engine import Engine
gold = np.load("tests/golden/step_n16_persp.npz")
n = 16
sd = synth.make_state_dict()
batch = synth.make_batch(n, "perspective", "flame", int(gold["seed"]))
x_t, x_input, clip = synth.make_inputs(n, 32, int(gold["seed"]))
noise = torch.randn(x_t.shape, generator=torch.Generator().manual_seed(int(gold["noise_seed"])))
eng = Engine(max_views_per_call=16); eng.load_state_dict(sd); eng.bind(batch, "perspective")
x = x_t[0].cuda().contiguous()
eps = eng.denoise_step(x, x_input[0].cuda().contiguous(), clip[0,0].cuda().contiguous(), int(gold["index"]), float(gold["cfg_scale"]), noise=noise[0].cuda().contiguous(), want_eps=True)
torch.cuda.synchronize()
ref = torch.from_numpy(gold["eps"])[0]
print("eps rel-L2 vs reference golden (N=16):", float((eps.cpu()-ref).norm()/ref.norm()), "mode", "fp32" if os.environ.get("MD_ST_FP32") else "bf16")
PY
tail -1 $O/eps_err.log
