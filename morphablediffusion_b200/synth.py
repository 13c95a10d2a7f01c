"""Deterministic synthetic weights and inputs for the denoise hot path.

No dataset, checkpoint or reference asset is needed: the mesh is a procedural FLAME-sized head (5 023 vertices) or
an SMPL-X-sized body (10 475), cameras follow the reference's virtual trajectory (generate_face.py:25-45,161-173)
or an orthographic ring like assets/thuman_meta.pkl, and every parameter — including the layers the reference
zero-initialises (openaimodel.py:230-232,720; ldm/modules/attention.py:319-323; attention.py:71) — is drawn from a
per-tensor seeded normal so that parity tests are not vacuous.
"""
import math
import zlib

import numpy as np
import torch

from . import spec as _spec


def _seed_for(key, seed):
    return (zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF


def init_tensor(key, shape, seed=6033):
    """fan-in scaled normal weights, N(0,0.05) biases, norm scales ~ 1, BN running stats as in §8d."""
    g = torch.Generator().manual_seed(_seed_for(key, seed))
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.tensor(100, dtype=torch.long)
    if leaf == "running_mean":
        return torch.randn(shape, generator=g) * 0.1
    if leaf == "running_var":
        return torch.rand(shape, generator=g) + 0.5
    if len(shape) == 1:
        if leaf == "weight":  # norm scale
            return 1.0 + 0.1 * torch.randn(shape, generator=g)
        return 0.05 * torch.randn(shape, generator=g)
    if "xyzc_net" in key:      # spconv layout [O,k,k,k,I]
        fan_in = int(np.prod(shape[1:]))
    elif ".up" in key and key.endswith("conv.weight") and len(shape) == 5 and "frustum" in key:
        fan_in = shape[0] * 27 // 8  # ConvTranspose3d [I,O,k,k,k]; ~27/8 taps hit each output
    else:
        fan_in = int(np.prod(shape[1:]))
    std = 1.0 / math.sqrt(max(fan_in, 1))
    return torch.randn(shape, generator=g) * std


def make_vae_state_dict(seed=6033, prefix="first_stage_model."):
    """Seeded post_quant_conv + Decoder weights under the reference's state-dict keys."""
    return {k: init_tensor(k, shp, seed) for k, shp in _spec.vae_decoder_spec(prefix).items()}


def make_vae_encoder_state_dict(seed=6033, prefix="first_stage_model."):
    """Seeded quant_conv + Encoder weights under the reference's state-dict keys."""
    return {k: init_tensor(k, shp, seed) for k, shp in _spec.vae_encoder_spec(prefix).items()}


def make_clip_state_dict(seed=6033, prefix="clip_image_encoder.model.visual."):
    """Seeded CLIP ViT-L/14 image-tower weights under the `clip` package's key names."""
    sd = {}
    for k, shp in _spec.clip_visual_spec(prefix).items():
        g = torch.Generator().manual_seed(_seed_for(k, seed))
        leaf = k.rsplit(".", 1)[-1]
        if leaf in ("class_embedding", "positional_embedding"):
            sd[k] = torch.randn(shp, generator=g) * 0.3
        elif leaf == "proj":
            sd[k] = torch.randn(shp, generator=g) * shp[0] ** -0.5
        elif leaf == "in_proj_bias":
            sd[k] = 0.05 * torch.randn(shp, generator=g)
        elif leaf == "in_proj_weight":
            sd[k] = torch.randn(shp, generator=g) * shp[1] ** -0.5
        else:
            sd[k] = init_tensor(k, shp, seed)
    return sd


def make_state_dict(cfg=None, seed=6033, keys=None):
    sd = {}
    for k, shp in _spec.model_spec(cfg).items():
        if keys is not None and not keys(k):
            continue
        sd[k] = init_tensor(k, shp, seed)
    return sd


# ----------------------------------------------------------------------------- geometry


def virtual_cameras(n=16, radius=4.5, focal=1545.23757707405, c=128.0):
    """generate_face.py:25-45,161-173: cameras on a half circle, R = Rx(-180°)·Ry(angle) ('xyz' extrinsic euler)."""
    Ks, RTs = [], []
    for ang in np.linspace(-90, 90, n):
        a = np.radians(ang)
        pos = np.array([radius * np.sin(a), 0.0, radius * np.cos(a)])
        rx, ry = np.radians(-180.0), a
        Rx = np.array([[1, 0, 0], [0, np.cos(rx), -np.sin(rx)], [0, np.sin(rx), np.cos(rx)]])
        Ry = np.array([[np.cos(ry), 0, np.sin(ry)], [0, 1, 0], [-np.sin(ry), 0, np.cos(ry)]])
        R = Ry @ Rx  # scipy Rotation.from_euler('xyz', ...) applies x first (extrinsic)
        t = -R @ pos.reshape(3, 1)
        K = np.eye(4)
        K[:3, :3] = np.array([[focal, 0, c], [0, focal, c], [0, 0, 1]])
        RT = np.zeros((3, 4))
        RT[:, :3] = R
        RT[:, 3] = t[:, 0]
        Ks.append(K)
        RTs.append(RT)
    return torch.tensor(np.array(Ks)).float(), torch.tensor(np.array(RTs)).float()


def ortho_cameras(n=16, radius=1.5, scale=1.6667):
    """Orthographic ring like assets/thuman_meta.pkl (K 4x4 with an isotropic scale, azimuths every 360/n deg)."""
    Ks, RTs = [], []
    for i in range(n):
        a = 2 * np.pi * i / n
        pos = np.array([radius * np.sin(a), 0.0, radius * np.cos(a)])
        fwd = -pos / np.linalg.norm(pos)
        up = np.array([0.0, 1.0, 0.0])
        right = np.cross(up, fwd)
        right /= np.linalg.norm(right)
        R = np.stack([right, -up, fwd], 0)  # world -> cam, y down
        t = -R @ pos
        K = np.diag([scale, scale, scale, 1.0])
        RT = np.zeros((3, 4))
        RT[:, :3] = R
        RT[:, 3] = t
        Ks.append(K)
        RTs.append(RT)
    return torch.tensor(np.array(Ks)).float(), torch.tensor(np.array(RTs)).float()


def head_mesh(nv=5023, seed=6033):
    """Procedural FLAME-sized head: points on a bumpy ellipsoid with extents close to the FLAME template x2.5
    (0.54 x 0.56 x 0.80 in x,y,z after the reference's axis swap) plus N(0, 0.002) jitter."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(nv, generator=g) * 2 - 1
    phi = torch.rand(nv, generator=g) * 2 * math.pi
    s = torch.sqrt(1 - u * u)
    d = torch.stack([s * torch.cos(phi), s * torch.sin(phi), u], -1)
    bump = 1.0 + 0.08 * torch.sin(5 * phi) * s + 0.05 * torch.cos(7 * u)
    r = torch.tensor([0.27, 0.28, 0.40])
    v = d * r * bump[:, None] + 0.002 * torch.randn(nv, 3, generator=g)
    return v.float()


def body_points(nv=10475, seed=6033):
    """SMPL-X-sized synthetic body: union of capsules inside [-0.45,0.45]^3."""
    g = torch.Generator().manual_seed(seed + 1)
    caps = [((0, -0.05, 0), (0, 0.25, 0), 0.11), ((0, 0.30, 0), (0, 0.40, 0), 0.07),
            ((-0.08, -0.05, 0), (-0.10, -0.43, 0), 0.05), ((0.08, -0.05, 0), (0.10, -0.43, 0), 0.05),
            ((-0.13, 0.22, 0), (-0.40, 0.20, 0), 0.035), ((0.13, 0.22, 0), (0.40, 0.20, 0), 0.035)]
    w = torch.tensor([4.0, 1.0, 2.0, 2.0, 1.5, 1.5])
    which = torch.multinomial(w / w.sum(), nv, replacement=True, generator=g)
    pts = torch.zeros(nv, 3)
    for i, (a, b, rad) in enumerate(caps):
        m = which == i
        k = int(m.sum())
        a_, b_ = torch.tensor(a, dtype=torch.float32), torch.tensor(b, dtype=torch.float32)
        tt = torch.rand(k, 1, generator=g)
        dirs = torch.randn(k, 3, generator=g)
        dirs = dirs / dirs.norm(dim=1, keepdim=True)
        pts[m] = a_ + tt * (b_ - a_) + rad * dirs
    return pts.clamp(-0.45, 0.45).float()


def voxelize_cpu(vertices):
    """Host-side voxelisation rule used to BUILD synthetic batches (generate_face.py:214-225); the product's GPU
    kernel (md_voxelize) is tested against the oracle's copy of this rule, not against this helper."""
    min_xyz = vertices.min(0).values
    max_xyz = vertices.max(0).values
    voxel = torch.tensor([0.005, 0.005, 0.005])
    coord = torch.round((vertices[:, [2, 1, 0]] - min_xyz[[2, 1, 0]]) / voxel).int()
    out_sh = torch.ceil((max_xyz[[2, 1, 0]] - min_xyz[[2, 1, 0]]) / voxel).int()
    out_sh = (out_sh | 3) + 1
    return coord, out_sh, torch.stack([min_xyz, max_xyz], 0)


def make_batch(n_views=16, projection="perspective", mesh="flame", seed=6033, unique_voxels=False, image_size=256):
    """The `batch` dict wire format of generate_face.py:227-241 (B = 1), CPU tensors.  image_size scales the
    perspective intrinsics (512 px: f = 3090.475, c = 256 — BASELINE config 1)."""
    if projection == "perspective":
        r = image_size / 256.0
        K, RT = virtual_cameras(n_views, focal=1545.23757707405 * r, c=128.0 * r)
    else:
        K, RT = ortho_cameras(n_views)
    v = head_mesh(seed=seed) if mesh == "flame" else body_points(seed=seed)
    if mesh == "flame" and projection == "orthographic":
        v = v  # same mesh is fine for orthographic plumbing tests
    coord, out_sh, bounds = voxelize_cpu(v)
    if unique_voxels:  # duplicate-free variant for strict-tolerance sparse-conv tests
        lin = (coord[:, 0].long() * 100000 + coord[:, 1].long()) * 100000 + coord[:, 2].long()
        _, first = np.unique(lin.numpy(), return_index=True)
        keep = torch.from_numpy(np.sort(first))
        v = v[keep]
        coord, out_sh, bounds = voxelize_cpu(v)
    z = torch.zeros(1, n_views)
    return {
        "input_elevation": torch.zeros(1, 1), "input_azimuth": torch.zeros(1, 1),
        "target_elevation": z.clone(), "target_azimuth": z.clone(),
        "target_K": K.unsqueeze(0), "target_RT": RT.unsqueeze(0),
        "vertices": v.unsqueeze(0), "coord": coord.unsqueeze(0), "out_sh": out_sh.unsqueeze(0),
        "bounds": bounds.unsqueeze(0),
    }


def make_inputs(n_views=16, latent=32, seed=6033):
    g = torch.Generator().manual_seed(seed)
    x_t = torch.randn(1, n_views, 4, latent, latent, generator=g)
    x_input = torch.randn(1, 4, latent, latent, generator=g)
    clip = torch.randn(1, 1, 768, generator=g)
    return x_t, x_input, clip


def step_noise(seed, index, shape):
    """Shared DDIM step noise of the trajectory parity test: the draw of DDIM index `index` (the reference run of
    oracle/make_golden.py and the GPU test both regenerate it from the seed)."""
    g = torch.Generator().manual_seed(seed * 1000 + index)
    return torch.randn(shape, generator=g)
