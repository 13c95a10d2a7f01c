"""ORACLE tooling: generate tests/golden/* by running the REAL reference modules (build container only).

  python -m oracle.make_golden [--skip-n16]

For every configuration the reference's own `SyncDDIMSampler.denoise_apply` code path
(/root/reference/ldm/models/diffusion/morphable_diffusion.py:701-739) is executed on CPU/fp32 with the seeded
synthetic state dict and inputs of morphablediffusion_b200/synth.py, and its outputs are stored (sub-sampled
where large).  The same run also evaluates oracle/ldm_oracle.py and prints the deviation — this is the step that
pins the restatement to the reference.
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from morphablediffusion_b200 import synth  # noqa: E402
from oracle import ldm_oracle as O  # noqa: E402
from oracle import ref_import  # noqa: E402

GOLD = ROOT / "tests" / "golden"


def maxerr(a, b):
    return float((a - b).abs().max()), float(b.abs().max())


def load_synth(model, seed=6033):
    sd = synth.make_state_dict(seed=seed)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [k for k in missing if not k.startswith(("betas", "alphas", "sqrt_", "posterior"))]
    assert not missing and not unexpected, (missing[:5], unexpected[:5])
    return sd


def ref_eps(model, x_t, x_input, clip, t, cfg_scale, batch, bvn):
    """The ε half of the reference's denoise_apply (lines 713-737), calling the reference's own methods."""
    B, N, C, H, W = x_t.shape
    v_embed = model.get_viewpoint_embedding(batch)
    t_embed = model.embed_time(t)
    vol = model.spatial_volume.construct_spatial_volume(x_t, t_embed, v_embed, batch)
    e_t = []
    frustum0 = None
    for ni in range(0, N, bvn):
        xs = x_t[:, ni:ni + bvn]
        VN = xs.shape[1]
        xs = xs.reshape(B * VN, C, H, W)
        ts = t.view(B, 1).repeat(1, VN).view(B * VN)
        idx = torch.arange(N)[ni:ni + bvn].unsqueeze(0).repeat(B, 1)
        clip_, feats, xc = model.get_target_view_feats(x_input, vol, clip, t_embed, v_embed, idx, batch)
        if frustum0 is None:
            frustum0 = {k: v[:1].clone() for k, v in feats.items()}
        e = model.model.predict_with_unconditional_scale(xs, ts, clip_, feats, xc, cfg_scale)
        e_t.append(e.view(B, VN, 4, H, W))
    return torch.cat(e_t, 1), vol, frustum0


def run_config(name, n_views, projection, mesh, index=49, cfg_scale=2.0, seed=6033, bvn=4, latent=32):
    torch.manual_seed(0)
    t0 = time.time()
    model, ns = ref_import.build_reference_model(projection=projection, view_num=n_views, cfg_scale=cfg_scale,
                                                 latent=latent)
    sd = load_synth(model, seed)
    batch = synth.make_batch(n_views, projection, mesh, seed, image_size=latent * 8)
    x_t, x_input, clip = synth.make_inputs(n_views, latent, seed)
    step = int(model.sampler.ddim_timesteps[index])
    t = torch.full((1,), step, dtype=torch.long)
    g = torch.Generator().manual_seed(seed + 7)
    noise = torch.randn(x_t.shape, generator=g)
    with torch.no_grad():
        eps, vol, fr0 = ref_eps(model, x_t, x_input, clip, t, cfg_scale, batch, bvn)
        # reference DDIM update with our noise: replicate denoise_apply_impl's arithmetic via its own tensors
        s = model.sampler
        a_t, a_prev = s.ddim_alphas[index], s.ddim_alphas_prev[index]
        sig, s1m = s.ddim_sigmas[index], s.ddim_sqrt_one_minus_alphas[index]
        x0 = (x_t - s1m * eps) / a_t.sqrt()
        x_prev = a_prev.sqrt() * x0 + torch.clamp(1. - a_prev - sig ** 2, min=1e-7).sqrt() * eps + sig * noise
        # is_step0 path of the reference's own denoise_apply_impl (no RNG) as a cross-check of the update
        x_prev0 = s.denoise_apply_impl(x_t, index, eps, is_step0=True)
        assert torch.allclose(x_prev0, x_prev - sig * noise, atol=1e-6)
    t_ref = time.time() - t0
    # ---- oracle on the same inputs
    t0 = time.time()
    cfg = O.VolumeCfg(projection=projection, num_views=n_views, input_image_size=latent * 8)
    sched = O.make_schedule()
    with torch.no_grad():
        o_eps, parts = O.denoise_eps(sd, cfg, x_t, x_input, clip, t, cfg_scale, batch, bvn, return_parts=True)
        o_prev = O.ddim_update(sched, x_t, index, o_eps, noise)
    t_or = time.time() - t0
    print(f"[{name}] ref {t_ref:.1f}s oracle {t_or:.1f}s  eps err/max {maxerr(o_eps, eps)}  "
          f"vol {maxerr(parts['spatial_volume'], vol)}  x_prev {maxerr(o_prev, x_prev)}", flush=True)
    np.savez_compressed(
        GOLD / f"{name}.npz",
        n_views=n_views, projection=projection, mesh=mesh, index=index, cfg_scale=cfg_scale, seed=seed, latent=latent,
        eps=eps.numpy(), x_prev=x_prev.numpy(), noise_seed=seed + 7,
        vol_sub=vol[:, :, ::4, ::4, ::4].numpy(), vol_absmean=float(vol.abs().mean()),
        **{f"frustum0_{k}": v[:, ::8, ::2, ::2, ::2].numpy() for k, v in fr0.items()},
        timestep=step)
    return model, sd


step_noise = synth.step_noise


def run_trajectory(name, n_views=2, steps=50, cfg_scale=2.0, seed=6033, keep=(40, 30, 20, 10, 0)):
    """The reference's SyncDDIMSampler.sample loop (morphable_diffusion.py:742-776) for all `steps` DDIM steps with
    the reference's own per-step methods; the only substitution is that the sigma_t * randn of denoise_apply_impl
    (:696) uses `step_noise` so the trajectory can be reproduced elsewhere."""
    torch.manual_seed(0)
    model, ns = ref_import.build_reference_model(view_num=n_views, cfg_scale=cfg_scale, sample_steps=steps)
    load_synth(model, seed)
    batch = synth.make_batch(n_views, "perspective", "flame", seed)
    x, x_input, clip = synth.make_inputs(n_views, 32, seed)
    s = model.sampler
    total = s.ddim_timesteps.shape[0]
    snaps = {}
    eps_first = None
    t0 = time.time()
    with torch.no_grad():
        for i, step in enumerate(np.flip(s.ddim_timesteps)):
            index = total - i - 1
            t = torch.full((1,), int(step), dtype=torch.long)
            eps, _, _ = ref_eps(model, x, x_input, clip, t, cfg_scale, batch, n_views)
            if eps_first is None:
                eps_first = eps.clone()
            x = s.denoise_apply_impl(x, index, eps, is_step0=True)      # the reference's update without its RNG draw
            if index != 0:
                x = x + s.ddim_sigmas[index] * step_noise(seed, index, x.shape)
            if index in keep:
                snaps[index] = x.clone()
            if i % 10 == 0:
                print(f"[{name}] step {i}/{total} |x| {float(x.abs().mean()):.4f} ({time.time() - t0:.0f}s)", flush=True)
    np.savez_compressed(GOLD / f"{name}.npz", n_views=n_views, steps=total, cfg_scale=cfg_scale, seed=seed,
                        x0=x.numpy(), eps_first=eps_first.numpy(),
                        **{f"x_at_{k}": v.numpy() for k, v in snaps.items()})
    print(f"[{name}] done in {time.time() - t0:.0f}s  |x0| mean {float(x.abs().mean()):.4f} max {float(x.abs().max()):.3f}")


def unet_only(seed=6033):
    """DepthWiseAttention.forward alone (batch 2) with a random source_dict."""
    model, ns = ref_import.build_reference_model()
    sd = load_synth(model, seed)
    g = torch.Generator().manual_seed(seed + 11)
    x = torch.randn(2, 8, 32, 32, generator=g)
    t = torch.tensor([981, 401])
    ctx = torch.randn(2, 1, 768, generator=g)
    src = {32: torch.randn(2, 64, 48, 32, 32, generator=g), 16: torch.randn(2, 128, 24, 16, 16, generator=g),
           8: torch.randn(2, 256, 12, 8, 8, generator=g), 4: torch.randn(2, 512, 6, 4, 4, generator=g)}
    with torch.no_grad():
        ref = model.model.diffusion_model(x, t, ctx, source_dict=src)
        ours = O.unet_forward(sd, x, t, ctx, src, prefix="model.diffusion_model.")
    print("[unet_b2] err/max", maxerr(ours, ref), flush=True)
    np.savez_compressed(GOLD / "unet_b2.npz", out=ref.numpy(), seed=seed, input_seed=seed + 11)


def vae_decode_golden(name, n_views, latent, seed=6033):
    """decode_first_stage (morphable_diffusion.py:468-471) through the reference's own Decoder class
    (ldm/modules/diffusionmodules/model.py:462-569) and a Conv2d post_quant_conv (autoencoder.py:303,330-333; the
    AutoencoderKL class itself needs `taming`, which is not installed, so its two-line decode() is spelled out)."""
    import importlib
    ref_import.reference()
    m = importlib.import_module("ldm.modules.diffusionmodules.model")
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
              num_res_blocks=2, attn_resolutions=[], dropout=0.0)   # morphable_diffusion.py:399-414
    dec = m.Decoder(**dd).eval()
    sd = synth.make_vae_state_dict(seed)
    dec.load_state_dict({k[len("first_stage_model.decoder."):]: v for k, v in sd.items() if ".decoder." in k})
    pq = torch.nn.Conv2d(4, 4, 1)
    pq.load_state_dict({"weight": sd["first_stage_model.post_quant_conv.weight"],
                        "bias": sd["first_stage_model.post_quant_conv.bias"]})
    x = torch.randn(n_views, 4, latent, latent, generator=torch.Generator().manual_seed(seed + 77))
    t0 = time.time()
    with torch.no_grad():
        ref = dec(pq(x / 0.18215))
        ours = O.vae_decode(sd, x / 0.18215)
    print(f"[{name}] oracle err/max", maxerr(ours, ref), f"{time.time() - t0:.0f}s", flush=True)
    np.savez_compressed(GOLD / f"{name}.npz", image=ref.numpy().astype(np.float16) if latent >= 32 else ref.numpy(),
                        n_views=n_views, latent=latent, seed=seed, input_seed=seed + 77)


def vae_encode_golden(name, n, size, seed=6033):
    """AutoencoderKL.encode up to the moments (autoencoder.py:324-328) through the reference's own Encoder class
    (model.py:368-459) and a Conv2d quant_conv (autoencoder.py:302), on seeded weights and a seeded image in [-1, 1]."""
    import importlib
    ref_import.reference()
    m = importlib.import_module("ldm.modules.diffusionmodules.model")
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
              num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    enc = m.Encoder(**dd).eval()
    sd = synth.make_vae_encoder_state_dict(seed)
    enc.load_state_dict({k[len("first_stage_model.encoder."):]: v for k, v in sd.items() if ".encoder." in k})
    qc = torch.nn.Conv2d(8, 8, 1)
    qc.load_state_dict({"weight": sd["first_stage_model.quant_conv.weight"], "bias": sd["first_stage_model.quant_conv.bias"]})
    x = torch.rand(n, 3, size, size, generator=torch.Generator().manual_seed(seed + 78)) * 2 - 1
    t0 = time.time()
    with torch.no_grad():
        ref = qc(enc(x))
        ours = O.vae_encode_moments(sd, x)
    print(f"[{name}] oracle err/max", maxerr(ours, ref), f"{time.time() - t0:.0f}s", flush=True)
    np.savez_compressed(GOLD / f"{name}.npz", moments=ref.numpy(), n=n, size=size, seed=seed, input_seed=seed + 78)


def clip_golden(name, n, size, seed=6033):
    """CLIP image embedding of a seeded image on seeded ViT-L/14 weights.  The `clip` package is not installed, so the
    independent implementation used to pin the restatement is HuggingFace transformers' CLIPVisionModelWithProjection
    (same architecture, its own code): the seeded weights are renamed into its state dict and both must agree."""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    sd = synth.make_clip_state_dict(seed)
    p = "clip_image_encoder.model.visual."
    cfg = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                           image_size=224, patch_size=14, projection_dim=768, hidden_act="quick_gelu", layer_norm_eps=1e-5)
    hf = CLIPVisionModelWithProjection(cfg).eval()
    m = {"vision_model.embeddings.class_embedding": sd[p + "class_embedding"],
         "vision_model.embeddings.patch_embedding.weight": sd[p + "conv1.weight"],
         "vision_model.embeddings.position_embedding.weight": sd[p + "positional_embedding"],
         "vision_model.pre_layrnorm.weight": sd[p + "ln_pre.weight"], "vision_model.pre_layrnorm.bias": sd[p + "ln_pre.bias"],
         "vision_model.post_layernorm.weight": sd[p + "ln_post.weight"], "vision_model.post_layernorm.bias": sd[p + "ln_post.bias"],
         "visual_projection.weight": sd[p + "proj"].t().contiguous()}
    for i in range(24):
        b, h = p + f"transformer.resblocks.{i}.", f"vision_model.encoder.layers.{i}."
        wq, wk, wv = sd[b + "attn.in_proj_weight"].chunk(3, 0)
        bq, bk, bv = sd[b + "attn.in_proj_bias"].chunk(3, 0)
        m.update({h + "self_attn.q_proj.weight": wq, h + "self_attn.k_proj.weight": wk, h + "self_attn.v_proj.weight": wv,
                  h + "self_attn.q_proj.bias": bq, h + "self_attn.k_proj.bias": bk, h + "self_attn.v_proj.bias": bv,
                  h + "self_attn.out_proj.weight": sd[b + "attn.out_proj.weight"], h + "self_attn.out_proj.bias": sd[b + "attn.out_proj.bias"],
                  h + "layer_norm1.weight": sd[b + "ln_1.weight"], h + "layer_norm1.bias": sd[b + "ln_1.bias"],
                  h + "layer_norm2.weight": sd[b + "ln_2.weight"], h + "layer_norm2.bias": sd[b + "ln_2.bias"],
                  h + "mlp.fc1.weight": sd[b + "mlp.c_fc.weight"], h + "mlp.fc1.bias": sd[b + "mlp.c_fc.bias"],
                  h + "mlp.fc2.weight": sd[b + "mlp.c_proj.weight"], h + "mlp.fc2.bias": sd[b + "mlp.c_proj.bias"]})
    missing, unexpected = hf.load_state_dict(m, strict=False)
    missing = [k for k in missing if "position_ids" not in k]
    assert not missing and not unexpected, (missing[:5], unexpected[:5])
    x = torch.rand(n, 3, size, size, generator=torch.Generator().manual_seed(seed + 79)) * 2 - 1
    t0 = time.time()
    with torch.no_grad():
        pre = O.clip_preprocess(x)
        ref = hf(pixel_values=pre).image_embeds.unsqueeze(1)
        ours = O.clip_image_embed(sd, x)
    print(f"[{name}] oracle vs transformers err/max", maxerr(ours, ref), f"{time.time() - t0:.0f}s", flush=True)
    np.savez_compressed(GOLD / f"{name}.npz", embed=ref.numpy(), pre_sub=pre[:, :, ::16, ::16].numpy(), n=n, size=size,
                        seed=seed, input_seed=seed + 79)


def train_forward_golden(name, n_views=4, seed=6033):
    """The forward half of the reference's training_step (morphable_diffusion.py:520-541), line by line through the
    reference's own methods (add_noise with a seeded torch generator state, embed_time, construct_spatial_volume,
    get_target_view_feats, UNetWrapper.forward(is_train=True)); the Lightning logging calls (:543-548) are left out."""
    model, ns = ref_import.build_reference_model(view_num=n_views)
    sd = load_synth(model, seed)
    batch = synth.make_batch(n_views, "perspective", "flame", seed)
    x_t, x_input, clip = synth.make_inputs(n_views, 32, seed)          # x_t doubles as the clean target latents
    B = 1
    time_steps = torch.tensor([437])
    target_index = torch.tensor([[2]])
    torch.manual_seed(seed + 80)
    with torch.no_grad():
        x_noisy, noise = model.add_noise(x_t, time_steps)
        v_embed = model.get_viewpoint_embedding(batch)
        t_embed = model.embed_time(time_steps)
        vol = model.spatial_volume.construct_spatial_volume(x_noisy, t_embed, v_embed, batch)
        clip_, feats, xc = model.get_target_view_feats(x_input, vol, clip, t_embed, v_embed, target_index, batch)
        x_noisy_ = x_noisy[torch.arange(B)[:, None], target_index][:, 0]
        pred = model.model(x_noisy_, time_steps, clip_, feats, xc, is_train=True)
        target = noise[torch.arange(B)[:, None], target_index][:, 0]
        loss = torch.nn.functional.mse_loss(target, pred, reduction="none").mean()
        o_loss, o_pred = O.training_forward(sd, O.VolumeCfg("perspective", num_views=n_views), batch, x_t, x_input, clip,
                                            time_steps, noise, target_index)
    print(f"[{name}] loss ref {float(loss):.6f} oracle {float(o_loss):.6f}  pred err/max {maxerr(o_pred, pred)}", flush=True)
    np.savez_compressed(GOLD / f"{name}.npz", n_views=n_views, seed=seed, time_step=437, target_index=2,
                        noise_seed=seed + 80, noise=noise.numpy(), pred=pred.numpy(), loss=float(loss))


def spec_dump():
    model, ns = ref_import.build_reference_model()
    skip = ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
            "posterior_variance", "posterior_log_variance_clipped")
    d = {k: list(v.shape) for k, v in model.state_dict().items() if k not in skip}
    (GOLD / "ref_state_dict_spec.json").write_text(json.dumps(d, indent=0))
    print("[spec] keys", len(d))
    # first-stage decoder (reference Decoder class; post_quant_conv is autoencoder.py:303 Conv2d(embed_dim, z_channels, 1))
    import importlib
    m = importlib.import_module("ldm.modules.diffusionmodules.model")
    dec = m.Decoder(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
                    num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    dv = {"first_stage_model.decoder." + k: list(v.shape) for k, v in dec.state_dict().items()}
    dv["first_stage_model.post_quant_conv.weight"] = [4, 4, 1, 1]
    dv["first_stage_model.post_quant_conv.bias"] = [4]
    (GOLD / "ref_vae_decoder_spec.json").write_text(json.dumps(dv, indent=0))
    print("[spec] first-stage decoder keys", len(dv))
    enc = m.Encoder(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
                    num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    ev = {"first_stage_model.encoder." + k: list(v.shape) for k, v in enc.state_dict().items()}
    ev["first_stage_model.quant_conv.weight"] = [8, 8, 1, 1]
    ev["first_stage_model.quant_conv.bias"] = [8]
    (GOLD / "ref_vae_encoder_spec.json").write_text(json.dumps(ev, indent=0))
    print("[spec] first-stage encoder keys", len(ev))


if __name__ == "__main__":
    GOLD.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    want = lambda n: not only or n in only
    if want("spec"):
        spec_dump()
    if want("unet_b2"):
        unet_only()
    if want("step_n4_persp"):
        run_config("step_n4_persp", 4, "perspective", "flame", index=49)
    if want("step_n4_ortho"):
        run_config("step_n4_ortho", 4, "orthographic", "body", index=20)
    if want("step_n16_persp") and "--skip-n16" not in sys.argv:
        run_config("step_n16_persp", 16, "perspective", "flame", index=49)
    # round 2: the remaining BASELINE.json configurations
    if want("step_n16_ortho_body"):   # config 4: SMPL-X-sized body (10 475 points), 16 orthographic views
        run_config("step_n16_ortho_body", 16, "orthographic", "body", index=30, bvn=8)
    if want("step_n8_persp"):         # config 5: view-count sweep (smpl_feature_extractor.num_views = N)
        run_config("step_n8_persp", 8, "perspective", "flame", index=10, bvn=8)
    if want("step_n32_persp"):
        run_config("step_n32_persp", 32, "perspective", "flame", index=40, bvn=8)
    if want("step_n4_lat64"):         # config 1: 4 views @ 64x64 latent (input_image_size = 512)
        run_config("step_n4_lat64", 4, "perspective", "flame", index=49, latent=64)
    if want("vae_n2_lat8"):           # §8f rank 1: VAE decode, small latent (CPU oracle test) and the real size (GPU test)
        vae_decode_golden("vae_n2_lat8", 2, 8)
    if want("vae_n2_lat32"):
        vae_decode_golden("vae_n2_lat32", 2, 32)
    if want("vae_n1_lat64"):          # BASELINE config 1's latent size (512x512 image)
        vae_decode_golden("vae_n1_lat64", 1, 64)
    if want("vae_enc_n2_64"):         # §8f rank 2 (VAE half): encoder moments, small image (CPU test) and 256x256 (GPU test)
        vae_encode_golden("vae_enc_n2_64", 2, 64)
    if want("vae_enc_n2_256"):
        vae_encode_golden("vae_enc_n2_256", 2, 256)
    if want("clip_n2_256"):           # §8f rank 2 (CLIP half): image embedding of two 256x256 images
        clip_golden("clip_n2_256", 2, 256)
    if want("train_n4"):              # §8f rank 3, forward half: the training loss of one step
        train_forward_golden("train_n4")
    if want("traj_n2_50"):            # a2: the 50-step sampler loop
        run_trajectory("traj_n2_50", 2, 50)
