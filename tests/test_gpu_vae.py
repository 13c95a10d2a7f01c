"""GPU parity of the first-stage decode (`pytest -m gpu`): md_vae_decode against golden images
produced by the reference's own Decoder class on the seeded weights (oracle/make_golden.py vae_n2_lat8 / vae_n2_lat32).
Tolerance: bf16 tensor-core operands with fp32 accumulation against the fp32 reference, rel-L2 <= 3e-2 and max-abs
<= 4 % of the image range (the bar of the denoise step)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
BF16_REL = 3e-2
BF16_MAX = 4e-2


def rel(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-20))


def maxrel(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-20))


@pytest.fixture(scope="module")
def vae_engine(vae_state_dict):
    from morphablediffusion_b200 import synth
    from morphablediffusion_b200.engine import Engine
    eng = Engine(max_views_per_call=16)
    sd = dict(synth.make_state_dict())
    sd.update(vae_state_dict)
    eng.load_state_dict(sd)
    assert eng.has_vae()
    yield eng
    eng.close()


@pytest.mark.parametrize("name", ["vae_n2_lat8", "vae_n2_lat32", "vae_n1_lat64"])
def test_vae_decode_vs_reference_golden(vae_engine, name):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    n, latent = int(gold["n_views"]), int(gold["latent"])
    x = torch.randn(n, 4, latent, latent, generator=torch.Generator().manual_seed(int(gold["input_seed"])))
    img = vae_engine.vae_decode(x.cuda())
    torch.cuda.synchronize()
    ref = torch.from_numpy(gold["image"].astype(np.float32))
    assert img.shape == ref.shape
    assert rel(img, ref) < BF16_REL and maxrel(img, ref) < BF16_MAX, (rel(img, ref), maxrel(img, ref))


def test_vae_decode_full_size_properties(vae_engine):
    """16 views at 32x32 latents (BASELINE config 2's decode): finite, image-scaled, and every view equals the same
    view decoded alone (views are independent: no cross-sample leakage through the batched GEMMs / statistics)."""
    x = torch.randn(16, 4, 32, 32, generator=torch.Generator().manual_seed(3)).cuda()
    img = vae_engine.vae_decode(x)
    one = vae_engine.vae_decode(x[5:6])
    torch.cuda.synchronize()
    assert img.shape == (16, 3, 256, 256) and torch.isfinite(img).all()
    assert rel(img[5:6], one) < 1e-2


@pytest.mark.parametrize("name", ["vae_enc_n2_64", "vae_enc_n2_256"])
def test_vae_encode_vs_reference_golden(state_dict, vae_encoder_state_dict, name):
    """md_vae_encode (moments = quant_conv(Encoder(image)), SURVEY §8f rank 2, VAE half) vs the reference Encoder class."""
    from morphablediffusion_b200.engine import Engine
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    n, size = int(gold["n"]), int(gold["size"])
    x = torch.rand(n, 3, size, size, generator=torch.Generator().manual_seed(int(gold["input_seed"]))) * 2 - 1
    sd = dict(state_dict)
    sd.update(vae_encoder_state_dict)
    eng = Engine(max_views_per_call=16)
    try:
        eng.load_state_dict(sd)
        assert eng.has_vae_encoder() and not eng.has_vae()
        mom = eng.vae_encode_moments(x.cuda())
        torch.cuda.synchronize()
    finally:
        eng.close()
    ref = torch.from_numpy(gold["moments"])
    assert mom.shape == ref.shape
    assert rel(mom, ref) < BF16_REL and maxrel(mom, ref) < BF16_MAX, (rel(mom, ref), maxrel(mom, ref))


def test_shell_encode_first_stage(state_dict, vae_encoder_state_dict):
    """encode_first_stage through the drop-in class: mode() equals the library's mean * 0.18215 and the golden; sample()
    follows torch's generator like DiagonalGaussianDistribution.sample."""
    from morphablediffusion_b200.ldm_api import SyncMultiviewDiffusion
    from oracle import ldm_oracle as O
    gold = np.load(os.path.join(GOLD, "vae_enc_n2_256.npz"))
    sd = dict(state_dict)
    sd.update(vae_encoder_state_dict)
    unet_config = {"target": "ldm.models.diffusion.attention.DepthWiseAttention",
                   "params": dict(volume_dims=[64, 128, 256, 512], image_size=32, in_channels=8, out_channels=4,
                                  model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
                                  channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                                  transformer_depth=1, context_dim=768, use_checkpoint=True, legacy=False)}
    model = SyncMultiviewDiffusion(unet_config, None, projection="perspective", view_num=2, cfg_scale=2.0)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().eval()
    x = torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(int(gold["input_seed"]))) * 2 - 1
    ref_mom = torch.from_numpy(gold["moments"])
    z_mode = model.encode_first_stage(x.cuda(), sample=False)
    assert rel(z_mode, O.vae_posterior_sample(ref_mom)) < BF16_REL
    torch.manual_seed(11)
    z = model.encode_first_stage(x.cuda(), sample=True)
    torch.manual_seed(11)
    noise = torch.randn(2, 4, 32, 32)
    assert rel(z, O.vae_posterior_sample(ref_mom, noise)) < BF16_REL


def test_clip_embed_vs_golden(state_dict, clip_state_dict):
    """md_clip_embed (CLIP half) vs the golden embedding (HuggingFace CLIPVisionModelWithProjection on the
    seeded ViT-L/14 weights, fp32): bicubic resize + normalisation + 24-layer image tower + projection."""
    from morphablediffusion_b200.engine import Engine
    gold = np.load(os.path.join(GOLD, "clip_n2_256.npz"))
    n, size = int(gold["n"]), int(gold["size"])
    x = torch.rand(n, 3, size, size, generator=torch.Generator().manual_seed(int(gold["input_seed"]))) * 2 - 1
    sd = dict(state_dict)
    sd.update(clip_state_dict)
    eng = Engine(max_views_per_call=16)
    try:
        eng.load_state_dict(sd)
        assert eng.has_clip()
        emb = eng.clip_embed(x.cuda())
        torch.cuda.synchronize()
    finally:
        eng.close()
    ref = torch.from_numpy(gold["embed"])
    assert emb.shape == ref.shape
    assert rel(emb, ref) < BF16_REL, rel(emb, ref)


def test_quickgelu_epilogue():
    from morphablediffusion_b200 import _native as nat
    torch.manual_seed(2)
    A = torch.randn(300, 256, device="cuda").to(torch.bfloat16)
    Wt = (torch.randn(512, 256, device="cuda") / 16).to(torch.bfloat16)
    bias = torch.randn(512, device="cuda")
    o = torch.zeros(300, 512, device="cuda")
    ob = torch.zeros(300, 512, device="cuda", dtype=torch.bfloat16)
    nat.conv_gemm(A, Wt, B=1, D=1, H=1, W=300, Cin=256, N=512, taps=[(0, 0, 0)], bias=bias, out_f32=o, act="quickgelu")
    nat.conv_gemm(A, Wt, B=1, D=1, H=1, W=300, Cin=256, N=512, taps=[(0, 0, 0)], bias=bias, out_bf16=ob, act="quickgelu")
    y = A.float() @ Wt.float().t() + bias
    ref = y * torch.sigmoid(1.702 * y)
    assert rel(o, ref) < 1e-5 and rel(ob, ref) < 5e-3


def test_model_sample_end_to_end(state_dict, vae_state_dict, vae_encoder_state_dict, clip_state_dict, tmp_path):
    """SyncMultiviewDiffusion.sample (morphable_diffusion.py:567-587) through the drop-in classes, every stage on the
    library: prepare (VAE encode + CLIP embed), the DDIM loop, decode of all views; return_inter_results and the
    validation / test hooks (:600-624), which only sample and write an image strip."""
    from morphablediffusion_b200 import batch as B, synth
    from morphablediffusion_b200.ldm_api import SyncDDIMSampler, SyncMultiviewDiffusion
    sd = dict(state_dict)
    sd.update(vae_state_dict)
    sd.update(vae_encoder_state_dict)
    sd.update(clip_state_dict)
    unet_config = {"target": "ldm.models.diffusion.attention.DepthWiseAttention",
                   "params": dict(volume_dims=[64, 128, 256, 512], image_size=32, in_channels=8, out_channels=4,
                                  model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
                                  channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                                  transformer_depth=1, context_dim=768, use_checkpoint=True, legacy=False)}
    n = 4
    model = SyncMultiviewDiffusion(unet_config, None, projection="perspective", view_num=n, cfg_scale=2.0, sample_steps=4,
                                   batch_view_num=4, output_num=1)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.split(".")[0] in ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod",
                                                      "sqrt_one_minus_alphas_cumprod", "posterior_variance",
                                                      "posterior_log_variance_clipped") for k in missing)
    model = model.cuda().eval()
    assert model._get_engine().has_clip() and model._get_engine().has_vae() and model._get_engine().has_vae_encoder()
    img = torch.rand(256, 256, 3, generator=torch.Generator().manual_seed(8)) * 2 - 1
    data = B.build_batch(img, synth.head_mesh() * 0.37, n_views=n)
    sampler = SyncDDIMSampler(model, 4, latent_size=32)
    torch.manual_seed(5)
    x = model.sample(sampler, data, 2.0, 4)
    assert x.shape == (1, n, 3, 256, 256) and torch.isfinite(x).all()
    torch.manual_seed(5)
    x2, inter = model.sample(sampler, data, 2.0, 4, return_inter_results=True, inter_interval=2, inter_view_interval=2)
    assert rel(x2, x) < 2e-2                                   # same torch seed -> same posterior sample, x_T, step seeds
    assert inter.shape[:2] == (1, 2) and inter.shape[3:] == (3, 256, 256) and torch.isfinite(inter).all()
    strip = B.image_strip(x)
    assert strip.shape == (256, n * 256, 3)
    model.image_dir, model.outdir = str(tmp_path / "log"), str(tmp_path / "test")
    model.validation_step(data, 0)
    model.test_step(data, 3)
    assert any((tmp_path / "log" / "images" / "val").iterdir()) and any((tmp_path / "test").iterdir())
    import pytest as _pt
    with _pt.raises(NotImplementedError):
        model.training_step(data)


def test_vae_missing_weights_fail_loudly(state_dict):
    from morphablediffusion_b200 import _native as nat
    from morphablediffusion_b200.engine import Engine
    eng = Engine()
    eng.load_state_dict(state_dict)
    assert not eng.has_vae()
    with pytest.raises(nat.MdiffError):
        eng.vae_decode(torch.zeros(1, 4, 32, 32, device="cuda"))
    eng.close()


def test_shell_decode_first_stage(state_dict, vae_state_dict):
    """SyncMultiviewDiffusion.decode_first_stage (morphable_diffusion.py:468-471) through the drop-in class, with the
    first-stage tensors loaded from a reference-keyed state dict."""
    from morphablediffusion_b200.ldm_api import SyncMultiviewDiffusion
    gold = np.load(os.path.join(GOLD, "vae_n2_lat32.npz"))
    sd = dict(state_dict)
    sd.update(vae_state_dict)
    unet_config = {"target": "ldm.models.diffusion.attention.DepthWiseAttention",
                   "params": dict(volume_dims=[64, 128, 256, 512], image_size=32, in_channels=8, out_channels=4,
                                  model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
                                  channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                                  transformer_depth=1, context_dim=768, use_checkpoint=True, legacy=False)}
    model = SyncMultiviewDiffusion(unet_config, None, projection="perspective", view_num=2, cfg_scale=2.0)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not [k for k in missing if k.startswith("first_stage_model.decoder.")]
    model = model.cuda().eval()
    x = torch.randn(2, 4, 32, 32, generator=torch.Generator().manual_seed(int(gold["input_seed"])))
    img = model.decode_first_stage(x.cuda())
    assert rel(img, torch.from_numpy(gold["image"].astype(np.float32))) < BF16_REL


def test_gemm_weight_pitch_and_row_softmax():
    """The two additions under the AttnBlock: a GEMM whose weight operand is a column slice of a wider matrix (keys
    inside the fused q|k activation, md_conv_gemm_args.Wpitch) and the row softmax; together: softmax(q k^T / sqrt(c))."""
    from morphablediffusion_b200 import _native as nat
    torch.manual_seed(9)
    S, C = 256, 512
    qk = torch.randn(S, 2 * C, device="cuda").to(torch.bfloat16)
    sc = torch.zeros(S, S, device="cuda")
    nat.conv_gemm(qk, qk[:, C:], B=1, D=1, H=1, W=S, Cin=C, Cpitch=2 * C, N=S, taps=[(0, 0, 0)], out_f32=sc,
                  out_scale=C ** -0.5, Wpitch=2 * C)
    ref = (qk[:, :C].float() @ qk[:, C:].float().t()) * C ** -0.5
    assert rel(sc, ref) < 1e-5
    pr = torch.zeros(S, S, device="cuda", dtype=torch.bfloat16)
    nat.check(nat.lib.md_op_softmax_rows(sc.data_ptr(), pr.data_ptr(), S, S, nat.cur_stream()), "softmax_rows")
    assert rel(pr, F.softmax(ref, dim=1)) < 5e-3
