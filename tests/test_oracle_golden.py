"""The oracle (CPU restatement) against golden vectors produced by the REAL reference modules
(oracle/make_golden.py, run in the build container where /root/reference is mounted)."""
import os

import numpy as np
import pytest
import torch

from morphablediffusion_b200 import synth
from oracle import ldm_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-4  # fp32 CPU restatement vs fp32 CPU reference: reassociation only


def rel(a, b):
    return float((a - b).norm() / b.norm())


def test_unet_forward_matches_reference(state_dict):
    gold = np.load(os.path.join(GOLD, "unet_b2.npz"))
    g = torch.Generator().manual_seed(int(gold["input_seed"]))
    x = torch.randn(2, 8, 32, 32, generator=g)
    t = torch.tensor([981, 401])
    ctx = torch.randn(2, 1, 768, generator=g)
    src = {32: torch.randn(2, 64, 48, 32, 32, generator=g), 16: torch.randn(2, 128, 24, 16, 16, generator=g),
           8: torch.randn(2, 256, 12, 8, 8, generator=g), 4: torch.randn(2, 512, 6, 4, 4, generator=g)}
    with torch.no_grad():
        out = O.unet_forward(state_dict, x, t, ctx, src, prefix="model.diffusion_model.")
    assert rel(out, torch.from_numpy(gold["out"])) < TOL


@pytest.mark.parametrize("name", ["step_n4_persp", "step_n4_ortho"])
def test_denoise_step_matches_reference(state_dict, name):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    n, proj, mesh = int(gold["n_views"]), str(gold["projection"]), str(gold["mesh"])
    index, scale, seed = int(gold["index"]), float(gold["cfg_scale"]), int(gold["seed"])
    batch = synth.make_batch(n, proj, mesh, seed)
    x_t, x_input, clip = synth.make_inputs(n, 32, seed)
    cfg = O.VolumeCfg(projection=proj, num_views=n)
    sched = O.make_schedule()
    assert int(sched["timesteps"][index]) == int(gold["timestep"])
    t = torch.full((1,), int(gold["timestep"]), dtype=torch.long)
    noise = torch.randn(x_t.shape, generator=torch.Generator().manual_seed(int(gold["noise_seed"])))
    with torch.no_grad():
        eps, parts = O.denoise_eps(state_dict, cfg, x_t, x_input, clip, t, scale, batch, 4, return_parts=True)
        x_prev = O.ddim_update(sched, x_t, index, eps, noise)
    assert rel(eps, torch.from_numpy(gold["eps"])) < TOL
    assert rel(x_prev, torch.from_numpy(gold["x_prev"])) < TOL
    assert rel(parts["spatial_volume"][:, :, ::4, ::4, ::4], torch.from_numpy(gold["vol_sub"])) < TOL


def test_vae_decode_matches_reference_decoder():
    """decode_first_stage restatement vs the reference's own Decoder class (tests/golden/vae_n2_lat8.npz, 8x8 latents so
    the CPU run takes a second); the same function at 32x32 latents is the checker of the GPU test."""
    gold = np.load(os.path.join(GOLD, "vae_n2_lat8.npz"))
    n, latent, seed = int(gold["n_views"]), int(gold["latent"]), int(gold["seed"])
    sd = synth.make_vae_state_dict(seed)
    x = torch.randn(n, 4, latent, latent, generator=torch.Generator().manual_seed(int(gold["input_seed"])))
    with torch.no_grad():
        img = O.vae_decode(sd, x / 0.18215)
    assert img.shape == (n, 3, 8 * latent, 8 * latent)
    assert rel(img, torch.from_numpy(gold["image"])) < TOL


def test_vae_encode_matches_reference_encoder():
    """AutoencoderKL.encode restatement (moments) vs the reference's own Encoder class on a 64x64 image."""
    gold = np.load(os.path.join(GOLD, "vae_enc_n2_64.npz"))
    n, size, seed = int(gold["n"]), int(gold["size"]), int(gold["seed"])
    sd = synth.make_vae_encoder_state_dict(seed)
    x = torch.rand(n, 3, size, size, generator=torch.Generator().manual_seed(int(gold["input_seed"]))) * 2 - 1
    with torch.no_grad():
        mom = O.vae_encode_moments(sd, x)
    assert mom.shape == (n, 8, size // 8, size // 8)
    assert rel(mom, torch.from_numpy(gold["moments"])) < TOL
    z = O.vae_posterior_sample(mom)                      # mode() * 0.18215
    assert torch.allclose(z, mom[:, :4] * 0.18215)


def test_clip_image_embedder_matches_independent_implementation(clip_state_dict):
    """FrozenCLIPImageEmbedder.encode restatement vs tests/golden/clip_n2_256.npz, which oracle/make_golden.py produced with
    HuggingFace transformers' CLIPVisionModelWithProjection on the same seeded weights (the `clip` package itself is
    absent); also the preprocess (bicubic align_corners resize + CLIP normalisation) on a sub-sampled grid."""
    gold = np.load(os.path.join(GOLD, "clip_n2_256.npz"))
    n, size = int(gold["n"]), int(gold["size"])
    x = torch.rand(n, 3, size, size, generator=torch.Generator().manual_seed(int(gold["input_seed"]))) * 2 - 1
    with torch.no_grad():
        pre = O.clip_preprocess(x)
        emb = O.clip_image_embed(clip_state_dict, x)
    assert emb.shape == (n, 1, 768)
    assert rel(pre[:, :, ::16, ::16], torch.from_numpy(gold["pre_sub"])) < 1e-6
    assert rel(emb, torch.from_numpy(gold["embed"])) < TOL


def test_training_forward_matches_reference(state_dict):
    """Forward half of training_step (loss of one step) vs the reference's own methods (tests/golden/train_n4.npz)."""
    gold = np.load(os.path.join(GOLD, "train_n4.npz"))
    n, seed = int(gold["n_views"]), int(gold["seed"])
    batch = synth.make_batch(n, "perspective", "flame", seed)
    x, x_input, clip = synth.make_inputs(n, 32, seed)
    with torch.no_grad():
        loss, pred = O.training_forward(state_dict, O.VolumeCfg("perspective", num_views=n), batch, x, x_input, clip,
                                        torch.tensor([int(gold["time_step"])]), torch.from_numpy(gold["noise"]),
                                        torch.tensor([[int(gold["target_index"])]]))
    assert rel(pred, torch.from_numpy(gold["pred"])) < TOL
    assert abs(float(loss) - float(gold["loss"])) < 1e-4 * float(gold["loss"])


def test_schedule_constants():
    s = O.make_schedule()
    assert s["timesteps"][0] == 1 and s["timesteps"][-1] == 981 and len(s["timesteps"]) == 50
    assert float(s["alphas_prev"][0]) > float(s["alphas"][0])  # a_prev[0] = acp[0], a[0] = acp[1]
    assert torch.all(s["alphas"][1:] < s["alphas"][:-1])


def test_voxelize_rule_edge_cases():
    # half-way cases round to even (torch.round), out_sh is forced to a multiple of 4
    v = torch.tensor([[0.0, 0.0, 0.0], [0.0025, 0.0075, 0.0125], [0.1, 0.2, 0.3]])
    coord, out_sh, bounds = O.voxelize(v)
    assert coord.dtype == torch.int32 and out_sh.dtype == torch.int32
    assert torch.all(out_sh % 4 == 0)
    assert coord[0].tolist() == [0, 0, 0]
    assert coord[2].tolist() == [60, 40, 20]
    assert torch.equal(bounds, torch.stack([v.min(0).values, v.max(0).values]))
