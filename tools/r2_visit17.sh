#!/bin/bash
# round-2 visit 17: CLIP image tower on the GPU, prepare() fully on the library, full suite
O=gpurun_out/r02r; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_vae.py -m gpu -x -q > $O/pytest_vae.log 2>&1; tail -15 $O/pytest_vae.log
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_vae.py > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python - > $O/clip_time.log 2>&1 <<PY
import sys, torch
sys.path.insert(0, ".")
from morphablediffusion_b200 import synth
from morphablediffusion_b200.engine import Engine
sd = dict(synth.make_state_dict()); sd.update(synth.make_clip_state_dict())
eng = Engine(max_views_per_call=16); eng.load_state_dict(sd)
for n in (1, 16):
    x = torch.rand(n, 3, 256, 256, device="cuda") * 2 - 1
    for _ in range(2): m = eng.clip_embed(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): m = eng.clip_embed(x)
    e1.record(); torch.cuda.synchronize()
    print(f"clip_embed {n} image(s): {e0.elapsed_time(e1) / 5:.2f} ms")
PY
cat $O/clip_time.log
