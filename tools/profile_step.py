"""One denoise step (16 views) bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import sys

import torch

sys.path.insert(0, ".")
from morphablediffusion_b200 import synth  # noqa: E402
from morphablediffusion_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sd = synth.make_state_dict()
batch = synth.make_batch(n)
x_t, x_input, clip = synth.make_inputs(n)
eng = Engine(max_views_per_call=min(n, 16))
eng.load_state_dict(sd)
eng.bind(batch, "perspective")
x = x_t[0].cuda().contiguous()
xi = x_input[0].cuda().contiguous()
cl = clip[0, 0].cuda().contiguous()
for i in range(2):
    eng.denoise_step(x, xi, cl, 49 - i, 2.0, seed=1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.denoise_step(x, xi, cl, 40, 2.0, seed=1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("workspace peak GB", eng.workspace_peak() / 2**30)
