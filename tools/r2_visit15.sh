#!/bin/bash
# round-2 visit 15: first-stage encoder on the GPU, shell encode/decode, full suite
O=gpurun_out/r02p; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_vae.py -m gpu -x -q > $O/pytest_vae.log 2>&1; tail -15 $O/pytest_vae.log
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_vae.py > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python - > $O/vae_time.log 2>&1 <<PY
import sys, torch
sys.path.insert(0, ".")
from morphablediffusion_b200 import synth
from morphablediffusion_b200.engine import Engine
sd = dict(synth.make_state_dict()); sd.update(synth.make_vae_encoder_state_dict())
eng = Engine(max_views_per_call=16); eng.load_state_dict(sd)
for n in (1, 16):
    x = torch.rand(n, 3, 256, 256, device="cuda") * 2 - 1
    for _ in range(2): m = eng.vae_encode_moments(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): m = eng.vae_encode_moments(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"vae_encode {n} image(s) @256x256: {ms:.2f} ms  ({n * 0.273 / ms:.3f} PFLOP/s of 273 GFLOP/image)")
PY
cat $O/vae_time.log
