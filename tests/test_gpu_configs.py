"""GPU parity for the remaining BASELINE.json configurations and the 50-step sampler loop (`pytest -m gpu`).

Golden vectors come from the REAL reference modules (oracle/make_golden.py, run in the build container):
  * config 1: 4 views @ 64x64 latent (SpatialVolumeNet(input_image_size=512), 512-px intrinsics);
  * config 4: SMPL-X-sized body (10 475 points), 16 orthographic views;
  * config 5: view-count sweep, N = 8 and N = 32 (smpl_feature_extractor.num_views = N), N = 64 by properties;
  * a2: SyncDDIMSampler.sample for all 50 DDIM steps with shared step noise.
Tolerances: the bf16 tensor-core path is held to rel-L2 <= 3e-2 / max-abs <= 4 % of range per step (same bar as
tests/test_gpu_parity.py); the 50-step trajectory bound is stated in its test.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
BF16_REL = 3e-2
BF16_MAX = 4e-2

UNET_PARAMS = dict(volume_dims=[64, 128, 256, 512], image_size=32, in_channels=8, out_channels=4,
                   model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2, channel_mult=[1, 2, 4, 4],
                   num_heads=8, use_spatial_transformer=True, transformer_depth=1, context_dim=768,
                   use_checkpoint=True, legacy=False)


def rel(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-20))


def maxrel(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-20))


@pytest.mark.parametrize("name,chunk", [("step_n16_ortho_body", 16), ("step_n8_persp", 8), ("step_n32_persp", 16),
                                        ("step_n4_lat64", 4)])
def test_denoise_step_baseline_configs(state_dict, name, chunk):
    """md_denoise_step vs the reference's own denoise_apply outputs at the sizes BASELINE.json names."""
    from morphablediffusion_b200 import synth
    from morphablediffusion_b200.engine import Engine
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    n, proj, mesh = int(gold["n_views"]), str(gold["projection"]), str(gold["mesh"])
    index, scale, seed, latent = int(gold["index"]), float(gold["cfg_scale"]), int(gold["seed"]), int(gold["latent"])
    batch = synth.make_batch(n, proj, mesh, seed, image_size=latent * 8)
    x_t, x_input, clip = synth.make_inputs(n, latent, seed)
    noise = torch.randn(x_t.shape, generator=torch.Generator().manual_seed(int(gold["noise_seed"])))
    eng = Engine(latent_size=latent, image_size=latent * 8, max_views_per_call=chunk)
    try:
        eng.load_state_dict(state_dict)
        eng.bind(batch, proj)
        assert eng.ddim_timestep(index) == int(gold["timestep"])
        x = x_t[0].cuda().contiguous()
        eps = eng.denoise_step(x, x_input[0].cuda().contiguous(), clip[0, 0].cuda().contiguous(), index, scale,
                               noise=noise[0].cuda().contiguous(), want_eps=True)
        torch.cuda.synchronize()
    finally:
        eng.close()
    eps_ref, xp_ref = torch.from_numpy(gold["eps"])[0], torch.from_numpy(gold["x_prev"])[0]
    assert rel(eps, eps_ref) < BF16_REL and maxrel(eps, eps_ref) < BF16_MAX, (rel(eps, eps_ref), maxrel(eps, eps_ref))
    assert rel(x, xp_ref) < BF16_REL


def test_view_sweep_n64_properties(state_dict):
    """N = 64 (top of BASELINE config 5) has no CPU golden (a reference step takes minutes); properties instead:
    (1) every view's epsilon is finite and of unit scale; (2) views are processed independently given the shared
    spatial volume, so the 64-view step in chunks of 16 equals the same step in chunks of 32 to the run-to-run level of two
    identical chunk-16 runs;
    (3) DDIM noise is keyed by the global view index, so the chunking does not change x_prev either."""
    from morphablediffusion_b200 import synth
    from morphablediffusion_b200.engine import Engine
    n = 64
    batch = synth.make_batch(n)
    x_t, x_input, clip = synth.make_inputs(n)
    outs = []
    for chunk in (16, 16, 32):
        eng = Engine(max_views_per_call=chunk)
        eng.load_state_dict(state_dict)
        eng.bind(batch, "perspective")
        x = x_t[0].cuda().contiguous()
        eps = eng.denoise_step(x, x_input[0].cuda().contiguous(), clip[0, 0].cuda().contiguous(), 33, 2.0, seed=5,
                               want_eps=True)
        torch.cuda.synchronize()
        outs.append((eps.cpu(), x.cpu()))
        eng.close()
    (e16, x16), (e16b, x16b), (e32, x32) = outs
    assert torch.isfinite(e16).all() and 0.3 < float(e16.std()) < 3.0
    # run-to-run level of the SAME configuration (fp32 atomics in the GroupNorm statistics and split-K arrival order
    # reorder sums; bf16 roundings downstream then flip): measured 7e-3 at N = 64, the same size as the error against
    # the fp32 reference.  A different chunking must stay at that level, and within the single-step parity bound.
    noise = rel(e16b, e16)
    assert noise < BF16_REL
    assert rel(e32, e16) < max(2.0 * noise, 2e-3) and rel(x32, x16) < max(2.0 * rel(x16b, x16), 2e-3), \
        (rel(e32, e16), noise)
    per_view = [rel(e32[v], e16[v]) for v in range(n)]
    assert max(per_view) < 3.0 * max(noise, 1e-3), max(per_view)   # no single view stands out (no chunk-edge bug)


def _shell(n_views, state_dict, sample_steps=50):
    from morphablediffusion_b200.ldm_api import SyncMultiviewDiffusion
    unet_config = {"target": "ldm.models.diffusion.attention.DepthWiseAttention", "params": dict(UNET_PARAMS)}
    model = SyncMultiviewDiffusion(unet_config, None, projection="perspective", view_num=n_views, cfg_scale=2.0,
                                   sample_steps=sample_steps)
    model.load_state_dict(state_dict, strict=False)
    return model.cuda().eval()


def test_sampler_50_step_trajectory_vs_reference(state_dict):
    """SyncDDIMSampler.sample (reference morphable_diffusion.py:742-776) for all 50 steps through the drop-in classes,
    against the same loop run with the reference's own modules in fp32 (tests/golden/traj_n2_50.npz) with shared x_T
    and shared per-step noise.  Stated drift bound: rel-L2(x_t) <= 6e-2 at every recorded index including x_0, and
    the first step's epsilon within the single-step tolerance (bf16 operands against fp32, 50 compounding steps)."""
    from morphablediffusion_b200 import synth
    from morphablediffusion_b200.synth import step_noise
    gold = np.load(os.path.join(GOLD, "traj_n2_50.npz"))
    n, steps, seed, scale = int(gold["n_views"]), int(gold["steps"]), int(gold["seed"]), float(gold["cfg_scale"])
    model = _shell(n, state_dict, steps)
    batch = {k: v.cuda() for k, v in synth.make_batch(n, "perspective", "flame", seed).items()}
    x_T, x_input, clip = synth.make_inputs(n, 32, seed)
    noise = torch.stack([step_noise(seed, i, x_T.shape) for i in range(steps)], 0)   # [steps, B, N, 4, h, w]
    x0, inter = model.sampler.sample({"x": x_input.cuda(), "elevation": None}, clip.cuda(), unconditional_scale=scale,
                                     log_every_t=10, batch_view_num=n, batch=batch, x_T=x_T, step_noise=noise.cuda())
    torch.cuda.synchronize()
    assert torch.isfinite(x0).all()
    # x_inter is recorded at index 49 (first step) and at every index divisible by 10
    logged = [steps - 1] + [i for i in range(steps - 2, -1, -1) if i % 10 == 0]
    assert len(inter["x_inter"]) == len(logged)
    drift = {}
    for idx, xi in zip(logged, inter["x_inter"]):
        key = f"x_at_{idx}"
        if key in gold.files:
            drift[idx] = rel(xi, torch.from_numpy(gold[key]))
    drift["x0"] = rel(x0, torch.from_numpy(gold["x0"]))
    print("trajectory drift (rel-L2 by DDIM index):", {k: round(v, 5) for k, v in drift.items()})
    assert set(drift) >= {40, 30, 20, 10, 0, "x0"}
    assert max(drift.values()) < 6e-2, drift


def test_sampler_with_other_step_counts(state_dict):
    """generate_face.py --sample_steps 20: SyncDDIMSampler(model, 20) must denoise index 19 at timestep 951 with the
    20-step alphas / sigmas (the library used to apply its own 50-step table)."""
    from morphablediffusion_b200 import synth
    from morphablediffusion_b200.ldm_api import SyncDDIMSampler
    from oracle import ldm_oracle as O
    n = 2
    model = _shell(n, state_dict)
    batch_cpu = synth.make_batch(n)
    batch = {k: v.cuda() for k, v in batch_cpu.items()}
    x_t, x_input, clip = synth.make_inputs(n)
    for steps, eta, index in ((20, 1.0, 19), (25, 0.0, 7)):
        sampler = SyncDDIMSampler(model, steps, "uniform", eta, latent_size=32)
        ts_val = int(sampler.ddim_timesteps[index])
        assert ts_val == index * (1000 // steps) + 1
        ts = torch.full((1,), ts_val, device="cuda", dtype=torch.long)
        out = sampler.denoise_apply(x_t.cuda(), {"x": x_input.cuda(), "elevation": None}, clip.cuda(), ts, index, 2.0,
                                    batch_view_num=n, is_step0=True, batch=batch)
        sched = O.make_schedule(steps, eta)
        with torch.no_grad():
            ref = O.denoise_apply(state_dict, O.VolumeCfg("perspective", num_views=n), sched, x_t, x_input, clip,
                                  ts.cpu(), index, 2.0, batch_cpu, noise=torch.zeros_like(x_t), batch_view_num=n)
        assert rel(out, ref) < BF16_REL, (steps, rel(out, ref))
    with pytest.raises(IndexError):
        sampler.denoise_apply(x_t.cuda(), {"x": x_input.cuda()}, clip.cuda(), None, 25, 2.0, batch=batch)
    # the binding key is content + identity, never addresses: an unchanged batch must not re-bind, an in-place camera
    # edit must (with random weights the epsilon itself barely depends on one camera, so the re-bind is observed
    # directly and through the frustum features of the edited view)
    eng = model._engine
    sampler.denoise_apply(x_t.cuda(), {"x": x_input.cuda()}, clip.cuda(), None, 7, 2.0, is_step0=True, batch=batch)
    binds = eng.bind_count
    sampler.denoise_apply(x_t.cuda(), {"x": x_input.cuda()}, clip.cuda(), None, 7, 2.0, is_step0=True, batch=batch)
    assert eng.bind_count == binds
    vol = eng.spatial_volume(x_t[0].cuda(), 500.0)
    f_a = eng.frustum_feats(vol, 0, n, 500.0)[32]
    batch["target_RT"][0, 1, :, 3] += torch.tensor([0.4, -0.3, 0.5], device="cuda")
    sampler.denoise_apply(x_t.cuda(), {"x": x_input.cuda()}, clip.cuda(), None, 7, 2.0, is_step0=True, batch=batch)
    assert eng.bind_count == binds + 1
    f_b = eng.frustum_feats(vol, 0, n, 500.0)[32]
    same, moved = rel(f_b[0], f_a[0]), rel(f_b[1], f_a[1])   # untouched view: bf16 run-to-run level (measured 4e-3)
    assert same < 1e-2 and moved > 3 * max(same, 3e-3), (same, moved)


def test_training_forward_loss_vs_reference_golden(state_dict):
    """SURVEY §8f rank 3, forward half: SyncMultiviewDiffusion.training_loss (the forward of training_step, morphable_
    diffusion.py:520-541, composed of the library's stage calls) against the reference's own methods with the same
    time step, noise and target view (tests/golden/train_n4.npz).  training_step itself refuses to run with gradients."""
    from morphablediffusion_b200 import synth
    gold = np.load(os.path.join(GOLD, "train_n4.npz"))
    n, seed = int(gold["n_views"]), int(gold["seed"])
    model = _shell(n, state_dict)
    batch = {k: v.cuda() for k, v in synth.make_batch(n, "perspective", "flame", seed).items()}
    x, x_input, clip = synth.make_inputs(n, 32, seed)
    loss, pred = model.training_loss(batch, x=x.cuda(), time_steps=torch.tensor([int(gold["time_step"])], device="cuda"),
                                     noise=torch.from_numpy(gold["noise"]).cuda(),
                                     target_index=torch.tensor([[int(gold["target_index"])]], device="cuda"),
                                     prepared=(clip.cuda(), {"x": x_input.cuda()}))
    assert rel(pred, torch.from_numpy(gold["pred"])) < BF16_REL, rel(pred, torch.from_numpy(gold["pred"]))
    assert abs(float(loss) - float(gold["loss"])) < 2e-2 * float(gold["loss"])
    with pytest.raises(NotImplementedError):
        model.training_step(batch)


def test_stage_methods_compose_like_the_reference(state_dict):
    """The body of the reference's denoise_apply (morphable_diffusion.py:713-737) written against the drop-in classes' stage
    methods — embed_time, construct_spatial_volume, get_target_view_feats, UNetWrapper.predict_with_unconditional_scale —
    must give the reference's epsilon (tests/golden/step_n4_persp.npz)."""
    from morphablediffusion_b200 import synth
    gold = np.load(os.path.join(GOLD, "step_n4_persp.npz"))
    n, seed, scale = int(gold["n_views"]), int(gold["seed"]), float(gold["cfg_scale"])
    model = _shell(n, state_dict)
    batch = {k: v.cuda() for k, v in synth.make_batch(n, "perspective", "flame", seed).items()}
    x_t, x_input, clip = (a.cuda() for a in synth.make_inputs(n, 32, seed))
    B, N, C, H, W = x_t.shape
    t = torch.full((B,), int(gold["timestep"]), device="cuda", dtype=torch.long)
    v_embed = model.get_viewpoint_embedding(batch)
    t_embed = model.embed_time(t)
    vol = model.spatial_volume.construct_spatial_volume(x_t, t_embed, v_embed, batch)
    e_t = []
    for ni in range(0, N, 2):                                             # batch_view_num = 2
        xs = x_t[:, ni:ni + 2].reshape(B * 2, C, H, W)
        ts = t.view(B, 1).repeat(1, 2).view(B * 2)
        idx = torch.arange(N, device="cuda")[ni:ni + 2].unsqueeze(0).repeat(B, 1)
        clip_, feats, xc = model.get_target_view_feats(x_input, vol, clip, t_embed, v_embed, idx, batch)
        e = model.model.predict_with_unconditional_scale(xs, ts, clip_, feats, xc, scale)
        e_t.append(e.view(B, 2, 4, H, W))
    eps = torch.cat(e_t, 1)
    assert rel(eps, torch.from_numpy(gold["eps"])) < BF16_REL, rel(eps, torch.from_numpy(gold["eps"]))


def test_sample_seeds_differ_between_calls_and_items(state_dict):
    """step noise is no longer one fixed Philox stream: it follows torch's generator per sample() call
    and differs between batch items; the same torch seed reproduces the same sample."""
    from morphablediffusion_b200 import synth
    n = 2
    model = _shell(n, state_dict, sample_steps=4)
    b1 = synth.make_batch(n)
    batch = {k: torch.cat([v, v], 0).cuda() for k, v in b1.items()}
    _, x_input, clip = synth.make_inputs(n)
    info = {"x": torch.cat([x_input, x_input], 0).cuda(), "elevation": None}
    clip2 = torch.cat([clip, clip], 0).cuda()
    x_T = torch.randn(1, n, 4, 32, 32).repeat(2, 1, 1, 1, 1)
    torch.manual_seed(1)
    a, _ = model.sampler.sample(info, clip2, 2.0, batch_view_num=n, batch=batch, x_T=x_T)
    torch.manual_seed(1)
    b, _ = model.sampler.sample(info, clip2, 2.0, batch_view_num=n, batch=batch, x_T=x_T)
    c, _ = model.sampler.sample(info, clip2, 2.0, batch_view_num=n, batch=batch, x_T=x_T)
    assert rel(a, b) < 5e-3                 # same torch seed -> same Philox seeds (bf16 run-to-run level)
    assert rel(a[0], a[1]) > 1e-2           # identical items, different step noise
    assert rel(a, c) > 1e-2                 # next call draws new seeds


@pytest.mark.parametrize("exchange", ["peer", "nccl", "peer_fallback"])
def test_multi_gpu_sharded_step_vs_reference_golden(tmp_path, exchange):
    """BASELINE config 3: the 16 views sharded over 2 ranks reproduce the REFERENCE golden of the unsharded step —
    epsilon and x_{t-1}, for plain launches, graph capture and graph replay — with the vertex-feature sums exchanged
    over NVLink peer memory inside the step's kernels (default) and through the NCCL all-reduce (MD_PEER=0);
    "peer_fallback": rank 1 pretends it cannot map its peers, and every rank must end up on the NCCL path.
    Needs two GPUs (gpurun --gpus 2); tools/mgpu_check.py is the worker."""
    import json
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "mgpu.json"
    port = 29600 + os.getpid() % 300 + {"peer": 0, "nccl": 301, "peer_fallback": 602}[exchange]
    env = dict(os.environ, MD_PEER="0" if exchange == "nccl" else "1")
    if exchange == "peer_fallback":
        env["MD_PEER_TEST_FAIL"] = "1"
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(root, "tools", "mgpu_check.py"), str(out)],
                       capture_output=True, text=True, timeout=900, cwd=root, env=env)
    assert p.returncode == 0, p.stderr[-3000:]
    res = json.loads(out.read_text())
    assert res["exchange"] == ("peer" if exchange == "peer" else "nccl"), res
    assert res["ok"] and res["eps_rel_l2"] < BF16_REL, res


def test_multi_gpu_sharded_sampling_through_the_drop_in_api(tmp_path):
    """SyncMultiviewDiffusion.enable_view_sharding() + sample() on 2 ranks (16 views, 5 DDIM steps) against the same model
    sampling every view on one GPU: identical latents on both ranks, within the bf16 tolerance of the single-GPU run,
    stage-level methods refused while sharded.  Needs two GPUs; tools/mgpu_sample_check.py is the worker."""
    import json
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "mgpu_sample.json"
    port = 30900 + os.getpid() % 300
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(root, "tools", "mgpu_sample_check.py"), str(out)],
                       capture_output=True, text=True, timeout=900, cwd=root)
    assert p.returncode == 0, p.stderr[-3000:]
    res = json.loads(out.read_text())
    assert res["ok"] and res["ranks_agree"] and res["latent_rel_l2_vs_single_gpu"] < 5e-2, res
