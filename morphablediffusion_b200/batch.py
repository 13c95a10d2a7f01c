"""Batch construction on the device: the wire format either side of the denoise loop.

Mirrors generate_face.py:25-45,156-241 (virtual camera trajectory, rigid alignment of the MICA-fitted FLAME mesh,
voxelisation rule, batch dict schema) and :243-249 (decoded images -> 8-bit strip), with the per-vertex / per-pixel
work in libmdiff (md_affine_points, md_voxelize, md_images_to_u8).  Host code composes only the 3x4 alignment map.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as nat
from .engine import voxelize
from .synth import virtual_cameras

# generate_face.py:205-213: hard-coded scale and pose aligning MICA-optimised FLAME meshes with the FaceScape fits
MICA_SCALE = 1.087
MICA_POSE = (1.6811e+00, -2.6845e-02, -2.8883e-02, 8.5418e-04, -3.4041e-03, 1.0564e-02)
WORLD_SCALE = 2.5
AXIS_SWAP = ((1.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, -1.0, 0.0))


def so3_exponential_map(log_rot):
    """Rodrigues' formula (pytorch3d.transforms.so3_exponential_map, used at generate_face.py:208), float64."""
    w = np.asarray(log_rot, dtype=np.float64).reshape(3)
    theta = float(np.sqrt(max(float(w @ w), 1e-8)))
    K = np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])
    return np.eye(3) + (np.sin(theta) / theta) * K + ((1.0 - np.cos(theta)) / theta ** 2) * (K @ K)


def alignment_map(scale=MICA_SCALE, pose=MICA_POSE, world_scale=WORLD_SCALE, swap=AXIS_SWAP):
    """v' = swap @ (world_scale * (R @ (scale * v) + T))  ==  A @ v + b   (generate_face.py:206-213)."""
    R = so3_exponential_map(pose[:3])
    T = np.asarray(pose[3:], dtype=np.float64)
    M = np.asarray(swap, dtype=np.float64)
    A = world_scale * scale * (M @ R)
    b = world_scale * (M @ T)
    return A.astype(np.float32), b.astype(np.float32)


def align_vertices(vertices, A=None, b=None):
    """vertices [Nv,3] CUDA fp32 -> aligned [Nv,3] (md_affine_points)."""
    if A is None:
        A, b = alignment_map()
    v = vertices.to(torch.float32).contiguous()
    if not v.is_cuda:
        raise nat.MdiffError("batch construction runs on a CUDA device (no CPU fallback)")
    out = torch.empty_like(v)
    a9 = (C.c_float * 9)(*[float(x) for x in np.asarray(A).reshape(-1)])
    b3 = (C.c_float * 3)(*[float(x) for x in np.asarray(b).reshape(-1)])
    nat.check(nat.lib.md_affine_points(v.data_ptr(), v.shape[0], a9, b3, out.data_ptr(), nat.cur_stream()),
              "md_affine_points")
    return out


def build_batch(input_image, mesh_vertices, cameras=None, n_views=16, align=True, device="cuda"):
    """The batch dict of generate_face.py:227-241 (B = 1), every tensor on `device`.

    input_image: [256,256,3] in [-1,1]; mesh_vertices: [Nv,3] raw fitted mesh (align=True applies the MICA alignment)
    or already-aligned world coordinates; cameras: (K [N,4,4], RT [N,3,4]) or None for the virtual half circle."""
    dev = torch.device(device)
    v = torch.as_tensor(mesh_vertices, dtype=torch.float32).to(dev)
    if align:
        v = align_vertices(v)
    coord, out_sh, bounds = voxelize(v)
    K, RT = cameras if cameras is not None else virtual_cameras(n_views)
    n = K.shape[0]
    img = torch.as_tensor(input_image, dtype=torch.float32).to(dev)
    zeros = torch.zeros(1, n, device=dev)
    return {"target_image": img.unsqueeze(0).unsqueeze(0).repeat(1, n, 1, 1, 1), "input_image": img.unsqueeze(0),
            "input_elevation": torch.zeros(1, 1, device=dev), "input_azimuth": torch.zeros(1, 1, device=dev),
            "target_elevation": zeros, "target_azimuth": zeros.clone(),
            "target_K": K.to(dev).float().unsqueeze(0), "target_RT": RT.to(dev).float().unsqueeze(0),
            "vertices": v.unsqueeze(0), "out_sh": out_sh.unsqueeze(0), "coord": coord.unsqueeze(0),
            "bounds": bounds.unsqueeze(0)}


def images_to_uint8(x_sample):
    """x_sample [B,N,3,H,W] fp32 -> uint8 [B,N,H,W,3] (generate_face.py:246-249)."""
    x = x_sample.to(torch.float32).contiguous()
    B, N, _, H, W = x.shape
    out = torch.empty(B, N, H, W, 3, dtype=torch.uint8, device=x.device)
    nat.check(nat.lib.md_images_to_u8(x.data_ptr(), out.data_ptr(), B * N, H, W, nat.cur_stream()), "md_images_to_u8")
    return out


def image_strip(x_sample):
    """The saved strip of generate_face.py:250-252: views side by side, samples stacked vertically (uint8 numpy)."""
    u8 = images_to_uint8(x_sample).cpu().numpy()
    return np.concatenate([np.concatenate(list(s), 1) for s in u8], 0)
