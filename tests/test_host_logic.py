"""Host-side logic that needs no GPU: synthetic generators, camera maths, config plumbing, and the N>1 view
sharding exchange (world_size 2 over gloo)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from morphablediffusion_b200 import synth
from morphablediffusion_b200.engine import viewpoint_embedding
from oracle import ldm_oracle as O


def test_synth_is_deterministic():
    a = synth.init_tensor("model.diffusion_model.out.2.weight", (4, 320, 3, 3))
    b = synth.init_tensor("model.diffusion_model.out.2.weight", (4, 320, 3, 3))
    assert torch.equal(a, b) and float(a.abs().sum()) > 0
    b1, b2 = synth.make_batch(4), synth.make_batch(4)
    assert all(torch.equal(b1[k], b2[k]) for k in b1)
    assert b1["coord"].dtype == torch.int32 and b1["vertices"].shape == (1, 5023, 3)
    assert torch.all(b1["out_sh"] % 4 == 0)


def test_virtual_cameras_match_scipy_euler():
    from scipy.spatial.transform import Rotation as Rot
    K, RT = synth.virtual_cameras(16)
    for i, ang in enumerate(np.linspace(-90, 90, 16)):
        R = Rot.from_euler("xyz", (-180, ang, 0), True).as_matrix()
        pos = np.array([4.5 * np.sin(np.radians(ang)), 0, 4.5 * np.cos(np.radians(ang))])
        assert np.allclose(RT[i, :, :3].numpy(), R, atol=1e-6)
        assert np.allclose(RT[i, :, 3].numpy(), -R @ pos, atol=1e-5)
    assert K.shape == (16, 4, 4) and float(K[0, 0, 0]) == pytest.approx(1545.2376, rel=1e-6)


def test_viewpoint_embedding_matches_oracle():
    b = synth.make_batch(8)
    b["target_azimuth"] = torch.linspace(0, 315, 8).unsqueeze(0)
    b["target_elevation"] = torch.full((1, 8), 30.0)
    assert torch.allclose(viewpoint_embedding(b), O.get_viewpoint_embedding(b))


def test_unique_voxel_batch_has_no_collisions():
    b = synth.make_batch(4, unique_voxels=True)
    c = b["coord"][0].long()
    lin = (c[:, 0] * 100000 + c[:, 1]) * 100000 + c[:, 2]
    assert lin.unique().numel() == lin.numel()
    b2 = synth.make_batch(4)
    c2 = b2["coord"][0].long()
    lin2 = (c2[:, 0] * 100000 + c2[:, 1]) * 100000 + c2[:, 2]
    assert lin2.unique().numel() < lin2.numel()  # the realistic mesh does collide (SURVEY: ~21 % for FLAME)


def test_instantiate_from_config_resolves_reference_targets():
    from morphablediffusion_b200 import ldm_api
    assert ldm_api.get_obj_from_str("ldm.models.diffusion.attention.DepthWiseAttention") is ldm_api.DepthWiseAttention
    with pytest.raises(KeyError):
        ldm_api.instantiate_from_config({})
    with pytest.raises(NotImplementedError):
        ldm_api.SpatialVolumeNet(256, 4, 16, projection="fisheye")


# ----------------------------------------------------------------------------- N > 1: view sharding over gloo
def _vertex_feature_sum(sd, cfg, x_views, t_embed, v_embed, batch, views):
    """Sum over `views` of the per-view vertex features (the quantity ranks all-reduce each step)."""
    import torch.nn.functional as F
    V = cfg.V
    verts = O.spatial_volume_verts(V, cfg.length, 1)
    Nv = batch["vertices"].shape[1]
    grid = (batch["vertices"] / cfg.length)[:, :, None, None, :]
    acc = torch.zeros(16, Nv)
    for ni in views:
        f = O.noisy_target_view_encoder(sd, x_views[:, ni], t_embed, v_embed[:, ni])
        coords = O.get_warp_coordinates(verts, 32, 256, batch["target_K"][:, ni], batch["target_RT"][:, ni],
                                        cfg.projection).view(1, V, V * V, 2)
        vol = F.grid_sample(f, coords, mode="bilinear", padding_mode="zeros", align_corners=True).view(1, 16, V, V, V)
        acc += F.grid_sample(vol, grid, mode="bilinear", padding_mode="zeros", align_corners=True)[0, :, :, 0, 0]
    return acc


def _shard_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    N = 4
    keys = lambda k: k.startswith(("time_embed.", "spatial_volume.target_encoder", "spatial_volume.smpl"))
    sd = synth.make_state_dict(keys=keys)
    batch = synth.make_batch(N)
    x_t, _, _ = synth.make_inputs(N)
    cfg = O.VolumeCfg(num_views=N)
    with torch.no_grad():
        t_embed = O.embed_time(sd, torch.tensor([981]))
        v_embed = O.get_viewpoint_embedding(batch)
        n_local = N // world
        mine = range(rank * n_local, (rank + 1) * n_local)
        part = _vertex_feature_sum(sd, cfg, x_t, t_embed, v_embed, batch, mine)
        dist.all_reduce(part)  # the one per-step exchange of the sharded path
        if rank == 0:
            full = _vertex_feature_sum(sd, cfg, x_t, t_embed, v_embed, batch, range(N))
            out.put(float((part - full).abs().max()))
    dist.barrier()
    dist.destroy_process_group()


def test_view_sharding_allreduce_equals_unsharded_sum():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-5


# ----------------------------------------------------------------------------- N > 1: sampler loop with sharded views
class _FakeEngine:
    """Stands in for the CUDA engine in the sampler-loop test: one 'step' adds a value that depends on the GLOBAL view
    index, the DDIM index and the seed, so a sharded run matches the unsharded one only if the view ranges, the seed
    broadcast and the gather are right."""

    def __init__(self):
        self.view0, self.n_local = 0, None

    def set_ddim(self, steps, eta):
        pass

    def denoise_step(self, x, xin, clip, index, scale, noise=None, seed=0):
        n = x.shape[0]
        assert self.n_local in (None, n)
        g = torch.arange(self.view0, self.view0 + n, dtype=torch.float32).view(n, 1, 1, 1)
        x.mul_(0.9).add_(0.01 * (g + 1.0) * (index + 1) + float(seed % 1000) * 1e-4)
        if noise is not None:
            x.add_(noise)


class _FakeModel:
    num_timesteps = 1000
    view_num = 4
    _device = torch.device("cpu")

    def __init__(self, shard):
        self.alphas_cumprod = torch.linspace(0.999, 0.01, 1000)
        self._shard = shard
        self.eng = _FakeEngine()

    def _bound_engine(self, batch_item, shard=False):
        assert shard == (self._shard is not None)
        n = batch_item["target_K"].shape[1]
        if shard:
            rank, world, _ = self._shard
            self.eng.view0, self.eng.n_local = rank * (n // world), n // world
        else:
            self.eng.view0, self.eng.n_local = 0, n
        return self.eng


def _sampler_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from morphablediffusion_b200.ldm_api import SyncDDIMSampler
    B, N = 2, 4
    batch = {"target_K": torch.zeros(B, N, 3, 3)}
    info = {"x": torch.zeros(B, 4, 8, 8)}
    clip = torch.zeros(B, 1, 768)
    g = torch.Generator().manual_seed(3)
    x_T = torch.randn(B, N, 4, 8, 8, generator=g)
    noise = 0.1 * torch.randn(5, B, N, 4, 8, 8, generator=g)
    sharded = SyncDDIMSampler(_FakeModel((rank, world, dist)), 5, latent_size=8)
    torch.manual_seed(100 + rank)                       # ranks disagree on the seed (and would on x_T): rank 0 rules
    xs, inter_s = sharded.sample(info, clip, 2.0, log_every_t=2, batch=batch, x_T=x_T, step_noise=noise)
    seed_used = sharded.seed
    torch.manual_seed(200 + rank)
    xr, _ = sharded.sample(info, clip, 2.0, batch=batch)  # x_T drawn per rank, then broadcast
    # unsharded loop with the seed rank 0 drew
    plain = SyncDDIMSampler(_FakeModel(None), 5, latent_size=8)
    torch.manual_seed(100)
    xp, inter_p = plain.sample(info, clip, 2.0, log_every_t=2, batch=batch, x_T=x_T, step_noise=noise)
    gathered = [None] * world
    dist.all_gather_object(gathered, (xs, xr, seed_used))
    if rank == 0:
        same_on_ranks = all(torch.equal(gathered[0][0], o[0]) and torch.equal(gathered[0][1], o[1]) and gathered[0][2] == o[2]
                            for o in gathered)
        inter_ok = len(inter_s["x_inter"]) == len(inter_p["x_inter"]) and all(
            torch.allclose(a, b) for a, b in zip(inter_s["x_inter"], inter_p["x_inter"]))
        out.put((same_on_ranks, float((xs - xp).abs().max()), inter_ok, plain.seed == seed_used))
    dist.barrier()
    dist.destroy_process_group()


def test_sampler_loop_with_sharded_views_matches_the_unsharded_loop():
    """SyncDDIMSampler.sample under enable_view_sharding (host logic only, the engine is a stand-in): every rank steps
    its own view range, x_T and the step seed come from rank 0, latents and logged intermediates are gathered on every rank."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sampler_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same_on_ranks, err, inter_ok, seed_ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same_on_ranks and seed_ok and inter_ok and err < 1e-6


def test_alignment_map_composes_the_reference_operations():
    """generate_face.py:203-213 as one affine map (batch.alignment_map) against the operation-by-operation oracle."""
    import numpy as np
    import torch
    from morphablediffusion_b200 import batch, synth
    from oracle import ldm_oracle as O
    R = batch.so3_exponential_map(batch.MICA_POSE[:3])
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(R) - 1.0) < 1e-12
    A, b = batch.alignment_map()
    v = synth.head_mesh()[:500] * 0.4
    ref = O.align_mica_vertices(v)
    got = v.double() @ torch.from_numpy(A).double().T + torch.from_numpy(b).double()
    assert float((got - ref.double()).abs().max()) < 2e-6

