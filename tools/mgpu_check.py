"""Multi-GPU parity: views sharded over ranks (one all-reduce per step) == unsharded single-GPU step.
torchrun --nproc-per-node 2 tools/mgpu_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from morphablediffusion_b200 import synth  # noqa: E402
from morphablediffusion_b200.engine import Engine, comm_unique_id  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
N = 8
sd = synth.make_state_dict()
batch = synth.make_batch(N)
x_t, x_input, clip = synth.make_inputs(N)
n_local = N // world
view0 = rank * n_local
eng = Engine(max_views_per_call=n_local)
eng.load_state_dict(sd)
uid = [comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
eng.init_comm(rank, world, uid[0])
eng.bind(batch, "perspective", view0=view0, n_local=n_local)
xin, cl = x_input[0].cuda().contiguous(), clip[0, 0].cuda().contiguous()
x = x_t[0, view0:view0 + n_local].cuda().contiguous()
for index in (49, 48):
    eng.denoise_step(x, xin, cl, index, 2.0, seed=6033)
torch.cuda.synchronize()
gathered = [torch.empty_like(x) for _ in range(world)]
dist.all_gather(gathered, x)
if rank == 0:
    full = Engine(max_views_per_call=N)
    full.load_state_dict(sd)
    full.bind(batch, "perspective")
    xf = x_t[0].cuda().contiguous()
    for index in (49, 48):
        full.denoise_step(xf, xin, cl, index, 2.0, seed=6033)
    torch.cuda.synchronize()
    got = torch.cat(gathered, 0)
    rel = float((got - xf).norm() / xf.norm())
    print(f"sharded x{world} vs unsharded after 2 steps: rel_l2 = {rel:.3e}  ({'OK' if rel < 2e-3 else 'FAIL'})", flush=True)
dist.barrier()
dist.destroy_process_group()
