"""Sharded sampling through the drop-in API on G GPUs: SyncMultiviewDiffusion.enable_view_sharding() + sample(), against the
same model sampling all views on rank 0 alone (same x_T, seed and conditioning; 5 DDIM steps).
    torchrun --nproc-per-node G --master-addr 127.0.0.1 tools/mgpu_sample_check.py [out.json]"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from morphablediffusion_b200 import batch as B, synth  # noqa: E402
from morphablediffusion_b200.ldm_api import SyncDDIMSampler, SyncMultiviewDiffusion  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
N, STEPS = 16, 5
sd = dict(synth.make_state_dict())
sd.update(synth.make_vae_state_dict())
sd.update(synth.make_vae_encoder_state_dict())
sd.update(synth.make_clip_state_dict())
unet_config = {"target": "ldm.models.diffusion.attention.DepthWiseAttention",
               "params": dict(volume_dims=[64, 128, 256, 512], image_size=32, in_channels=8, out_channels=4,
                              model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
                              channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                              transformer_depth=1, context_dim=768, use_checkpoint=True, legacy=False)}


def build():
    m = SyncMultiviewDiffusion(unet_config, None, projection="perspective", view_num=N, cfg_scale=2.0, sample_steps=STEPS,
                               batch_view_num=N, output_num=1)
    m.load_state_dict(sd, strict=False)
    return m.cuda().eval()


img = torch.rand(256, 256, 3, generator=torch.Generator().manual_seed(8)) * 2 - 1
data = B.build_batch(img, synth.head_mesh() * 0.37, n_views=N)
rel = lambda a, b: float((a - b).norm() / b.norm())
res = {"world": world, "views": N, "steps": STEPS}

model = build()
assert model.enable_view_sharding(dist)
res["exchange"] = "peer" if model._get_engine().peer_exchange_attached() else "nccl"
sampler = SyncDDIMSampler(model, STEPS, latent_size=32)
torch.manual_seed(5 + 17 * rank)                # the ranks' generators disagree on purpose: rank 0 rules
_, clip, info = model.prepare(data)
x_sh, inter = sampler.sample(info, clip, 2.0, log_every_t=2, batch=data)
torch.manual_seed(5 + 17 * rank)
imgs = model.sample(sampler, data, 2.0, N)
torch.cuda.synchronize()
ok_shape = tuple(imgs.shape) == (1, N, 3, 256, 256) and bool(torch.isfinite(imgs).all())
parts = [torch.empty_like(x_sh) for _ in range(world)]
dist.all_gather(parts, x_sh)
res["ranks_agree"] = all(torch.equal(parts[0], p) for p in parts)
try:
    model._bound_engine(data)       # what every stage-level method (all views bound) goes through
    res["stage_methods_refused"] = False
except RuntimeError:
    res["stage_methods_refused"] = True
if rank == 0:
    plain = build()
    s2 = SyncDDIMSampler(plain, STEPS, latent_size=32)
    torch.manual_seed(5)
    _, clip2, info2 = plain.prepare(data)
    x_pl, inter2 = s2.sample(info2, clip2, 2.0, log_every_t=2, batch=data)
    torch.cuda.synchronize()
    res["latent_rel_l2_vs_single_gpu"] = rel(x_sh, x_pl)
    res["inter_rel_l2"] = max(rel(a, b) for a, b in zip(inter["x_inter"], inter2["x_inter"]))
    res["images_ok"] = ok_shape
    res["tolerance"] = 5e-2
    res["ok"] = bool(res["ranks_agree"] and ok_shape and res["stage_methods_refused"]
                     and res["latent_rel_l2_vs_single_gpu"] < 5e-2)
    line = json.dumps(res)
    print("MGPU_SAMPLE " + line, flush=True)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write(line + "\n")
dist.barrier()
dist.destroy_process_group()
