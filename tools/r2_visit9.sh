#!/bin/bash
# round-2 visit 9: first-stage decoder on the GPU (parity vs the reference Decoder goldens), full suite, decode timing
O=gpurun_out/r02j; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_vae.py -m gpu -x -q > $O/pytest_vae.log 2>&1; tail -15 $O/pytest_vae.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_vae.py > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python - > $O/vae_time.log 2>&1 <<PY
import sys, torch
sys.path.insert(0, ".")
from morphablediffusion_b200 import synth
from morphablediffusion_b200.engine import Engine
sd = dict(synth.make_state_dict()); sd.update(synth.make_vae_state_dict())
eng = Engine(max_views_per_call=16); eng.load_state_dict(sd)
x = torch.randn(16, 4, 32, 32, device="cuda")
for _ in range(2): img = eng.vae_decode(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): img = eng.vae_decode(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"vae_decode 16 views @256x256: {ms:.2f} ms  ({16 * 0.622 / ms:.1f} TFLOP/s of 622 GFLOP/view)  workspace peak {eng.workspace_peak() / 2**30:.2f} GB")
PY
cat $O/vae_time.log
