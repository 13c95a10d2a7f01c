"""Chunk invariance of one denoise step: the same N-view step in chunks of different sizes, twice each, with the
per-view rel-L2 spread.  `python tools/diag_chunks.py [N] [chunk ...]` (development)."""
import sys

import torch

sys.path.insert(0, ".")
from morphablediffusion_b200 import synth  # noqa: E402
from morphablediffusion_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
chunks = [int(a) for a in sys.argv[2:]] or [16, 32]
sd = synth.make_state_dict()
batch = synth.make_batch(n)
x_t, x_input, clip = synth.make_inputs(n)


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


outs = {}
for chunk in chunks:
    for rep in range(2):
        eng = Engine(max_views_per_call=chunk)
        eng.load_state_dict(sd)
        eng.bind(batch, "perspective")
        x = x_t[0].cuda().contiguous()
        eps = eng.denoise_step(x, x_input[0].cuda().contiguous(), clip[0, 0].cuda().contiguous(), 33, 2.0, seed=5,
                               want_eps=True)
        torch.cuda.synchronize()
        outs[(chunk, rep)] = eps.float().cpu()
        eng.close()
keys = list(outs)
base = outs[keys[0]]
for k in keys[1:]:
    e = outs[k]
    per_view = [rel(e[v], base[v]) for v in range(n)]
    print(f"chunk={k[0]} rep={k[1]} vs chunk={keys[0][0]} rep=0: rel={rel(e, base):.3e} "
          f"per-view min={min(per_view):.2e} max={max(per_view):.2e} argmax={per_view.index(max(per_view))}", flush=True)
