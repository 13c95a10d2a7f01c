"""Device-timed md_spatial_volume (target-view encoder -> vertex features -> sparse conv net -> resample) on the bench mesh:
`python tools/time_volume.py [views] [body]`; MD_SPARSE_TILE=0 switches the wide sparse layers back to the quad kernel."""
import sys

import torch

sys.path.insert(0, ".")
from morphablediffusion_b200 import synth  # noqa: E402
from morphablediffusion_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
mesh = "body" if "body" in sys.argv[2:] else "flame"
proj = "orthographic" if mesh == "body" else "perspective"
sd = synth.make_state_dict()
batch = synth.make_batch(n, proj, mesh)
x_t, _, _ = synth.make_inputs(n)
eng = Engine(max_views_per_call=n)
eng.load_state_dict(sd)
eng.bind(batch, proj)
x = x_t[0].cuda().contiguous()
te = eng.embed_time(500)
for _ in range(5):
    vol = eng.spatial_volume(x, te)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    vol = eng.spatial_volume(x, te)
e1.record()
torch.cuda.synchronize()
print(f"mesh={mesh} views={n} spatial_volume {e0.elapsed_time(e1) / 50 * 1e3:.1f} us  checksum {float(vol.double().sum()):.6f}")
