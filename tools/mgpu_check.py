"""Multi-GPU parity against the REFERENCE golden: the 16 views of tests/golden/step_n16_persp.npz sharded over the
ranks (one exchange of the vertex-feature sums per step: NVLink peer push, or the NCCL all-reduce with MD_PEER=0), epsilon and x_{t-1} gathered and compared with what
the real reference modules produced for the unsharded step.
    torchrun --nproc-per-node G --master-addr 127.0.0.1 tools/mgpu_check.py [out.json]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from morphablediffusion_b200 import synth  # noqa: E402
from morphablediffusion_b200.engine import Engine, comm_unique_id  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
gold = np.load(os.path.join(ROOT, "tests", "golden", "step_n16_persp.npz"))
N, index, scale, seed = int(gold["n_views"]), int(gold["index"]), float(gold["cfg_scale"]), int(gold["seed"])
sd = synth.make_state_dict(seed=seed)
batch = synth.make_batch(N, str(gold["projection"]), str(gold["mesh"]), seed)
x_t, x_input, clip = synth.make_inputs(N, 32, seed)
noise = torch.randn(x_t.shape, generator=torch.Generator().manual_seed(int(gold["noise_seed"])))
n_local = N // world
view0 = rank * n_local
eng = Engine(max_views_per_call=n_local)
eng.load_state_dict(sd)
uid = [comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
eng.init_comm(rank, world, uid[0])
if os.environ.get("MD_PEER", "1") != "0":
    eng.init_peer_exchange(dist)
eng.bind(batch, str(gold["projection"]), view0=view0, n_local=n_local)
xin, cl = x_input[0].cuda().contiguous(), clip[0, 0].cuda().contiguous()
nz = noise[0, view0:view0 + n_local].cuda().contiguous()
res = {}
rel = lambda a, b: float((a.cpu() - b).norm() / b.norm())
x_start = x_t[0, view0:view0 + n_local].cuda().contiguous()
x = x_start.clone()


def gathered(t):
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return torch.cat(parts, 0)


# three calls with the same buffers: plain launches, graph capture, graph replay — each must match the reference
for call in range(3):
    x.copy_(x_start)
    eng.denoise_step(x, xin, cl, index, scale, noise=nz)
    torch.cuda.synchronize()
    xp = gathered(x)
    if rank == 0:
        res[f"x_prev_rel_l2_call{call}"] = rel(xp, torch.from_numpy(gold["x_prev"])[0])
x.copy_(x_start)
eps = eng.denoise_step(x, xin, cl, index, scale, noise=nz, want_eps=True)
torch.cuda.synchronize()
ge = gathered(eps)
if rank == 0:
    res["eps_rel_l2"] = rel(ge, torch.from_numpy(gold["eps"])[0])
if rank == 0:
    res.update(world=world, views=N, views_per_gpu=n_local, tolerance=3e-2,
               exchange="peer" if eng.peer_exchange_attached() else "nccl",
               ok=bool(max(v for k, v in res.items() if "rel_l2" in k) < 3e-2))
    line = json.dumps(res)
    print("MGPU " + line, flush=True)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write(line + "\n")
dist.barrier()
dist.destroy_process_group()
