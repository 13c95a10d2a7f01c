"""Device-timed denoise steps for a list of view counts on one GPU: `python tools/time_step.py 2 4 8 16 [--steps K]`."""
import sys

import torch

sys.path.insert(0, ".")
from morphablediffusion_b200 import synth  # noqa: E402
from morphablediffusion_b200.engine import Engine  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
steps = 20
for a in sys.argv[1:]:
    if a.startswith("--steps="):
        steps = int(a.split("=")[1])
views = [int(a) for a in args] or [16]
sd = synth.make_state_dict()
for n in views:
    batch = synth.make_batch(n)
    x_t, x_input, clip = synth.make_inputs(n)
    eng = Engine(max_views_per_call=n)
    eng.load_state_dict(sd)
    eng.bind(batch, "perspective")
    x = x_t[0].cuda().contiguous()
    xi = x_input[0].cuda().contiguous()
    cl = clip[0, 0].cuda().contiguous()
    for i in range(4):
        eng.denoise_step(x, xi, cl, 49 - i, 2.0, seed=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        eng.denoise_step(x, xi, cl, 45 - (i % 40), 2.0, seed=1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"views={n} ms_per_step={ms:.3f} steps_per_s={1e3 / ms:.2f} finite={bool(torch.isfinite(x).all())}", flush=True)
    del eng
    torch.cuda.empty_cache()
