"""Developer check: every stage of the CUDA path against the oracle / golden fixtures (run on a B200)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from morphablediffusion_b200 import synth  # noqa: E402
from morphablediffusion_b200.engine import Engine, voxelize  # noqa: E402
from oracle import ldm_oracle as O  # noqa: E402

dev = "cuda"


def rel(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs().max().item()
    return err, ref.abs().max().item(), ((got - ref).norm() / ref.norm().clamp_min(1e-12)).item()


def report(name, got, ref, tol):
    e, m, r = rel(got, ref)
    ok = r <= tol
    print(f"[{'OK' if ok else 'FAIL'}] {name}: max_abs={e:.3e} ref_max={m:.3e} rel_l2={r:.3e} (tol {tol})", flush=True)
    return ok


def main():
    which = sys.argv[1:] or ["vox", "vol", "fr", "unet", "step"]
    torch.manual_seed(0)
    t0 = time.time()
    sd = synth.make_state_dict()
    print(f"state dict built in {time.time()-t0:.1f}s", flush=True)
    N = 4
    batch = synth.make_batch(N, "perspective", "flame")
    x_t, x_input, clip = synth.make_inputs(N)
    ok = True

    if "vox" in which:
        v = batch["vertices"][0].to(dev)
        coord, out_sh, bounds = voxelize(v)
        oc, osh, ob = O.voxelize(batch["vertices"][0])
        same = torch.equal(coord.cpu(), oc) and torch.equal(out_sh.cpu(), osh) and torch.equal(bounds.cpu(), ob)
        print(f"[{'OK' if same else 'FAIL'}] voxelize bit-exact: out_sh={out_sh.tolist()}", flush=True)
        ok &= same

    eng = Engine(smpl_num_views=N)
    t0 = time.time()
    eng.load_state_dict(sd)
    print(f"weights loaded in {time.time()-t0:.1f}s", flush=True)
    eng.bind(batch, "perspective")
    cfg = O.VolumeCfg("perspective", num_views=N)
    tstep = 981
    t = torch.full((1,), tstep, dtype=torch.long)
    with torch.no_grad():
        t_embed = O.embed_time(sd, t)
        v_embed = O.get_viewpoint_embedding(batch)

    vol_ref = None
    if "vol" in which or "fr" in which:
        with torch.no_grad():
            vol_ref = O.construct_spatial_volume(sd, cfg, x_t, t_embed, v_embed, batch)
    if "vol" in which:
        vol = eng.spatial_volume(x_t[0].to(dev), tstep)
        torch.cuda.synchronize()
        ok &= report("spatial_volume", vol, vol_ref, 2e-3)

    if "fr" in which:
        with torch.no_grad():
            fr_ref, _ = O.construct_view_frustum_volume(sd, cfg, vol_ref, t_embed, v_embed, torch.tensor([[1, 2]]), batch)
        fr = eng.frustum_feats(vol_ref.to(dev), 1, 2, tstep)
        torch.cuda.synchronize()
        for k in sorted(fr_ref, reverse=True):
            ok &= report(f"frustum level {k}", fr[k], fr_ref[k], 3e-2)

    if "unet" in which:
        gold = np.load("tests/golden/unet_b2.npz")
        g = torch.Generator().manual_seed(int(gold["input_seed"]))
        x = torch.randn(2, 8, 32, 32, generator=g)
        tt = torch.tensor([981, 401])
        ctx = torch.randn(2, 1, 768, generator=g)
        src = {32: torch.randn(2, 64, 48, 32, 32, generator=g), 16: torch.randn(2, 128, 24, 16, 16, generator=g),
               8: torch.randn(2, 256, 12, 8, 8, generator=g), 4: torch.randn(2, 512, 6, 4, 4, generator=g)}
        out = eng.unet_forward(x.to(dev), tt, ctx.to(dev), {k: v.to(dev) for k, v in src.items()})
        torch.cuda.synchronize()
        ok &= report("unet_forward vs reference golden", out, torch.from_numpy(gold["out"]), 3e-2)

    if "step" in which:
        gold = np.load("tests/golden/step_n4_persp.npz")
        index = int(gold["index"])
        g = torch.Generator().manual_seed(int(gold["noise_seed"]))
        noise = torch.randn(x_t.shape, generator=g)
        xl = x_t[0].to(dev).contiguous()
        eps = eng.denoise_step(xl, x_input[0].to(dev).contiguous(), clip[0, 0].to(dev).contiguous(), index,
                               float(gold["cfg_scale"]), noise=noise[0].to(dev).contiguous(), want_eps=True)
        torch.cuda.synchronize()
        ok &= report("denoise_step eps vs reference golden", eps, torch.from_numpy(gold["eps"])[0], 3e-2)
        ok &= report("denoise_step x_prev vs reference golden", xl, torch.from_numpy(gold["x_prev"])[0], 3e-2)
        print("workspace peak GB:", eng.workspace_peak() / 2**30)
    print("ALL OK" if ok else "SOME FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
