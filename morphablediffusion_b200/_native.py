"""ctypes binding of libmdiff.so (the C ABI declared in include/mdiff.h).

The library is the product: there is no Python/torch fallback.  Importing this module without the built
shared object raises, and every wrapper raises ``MdiffError`` with the library's message on a non-zero return.
"""
import ctypes as C
import os
from pathlib import Path

_ROOT = Path(__file__).resolve().parent
LIB_PATH = _ROOT / "_lib" / "libmdiff.so"
if os.environ.get("MD_BUILD_TAG"):  # development variant built by build.py with the same tag (e.g. phase-stamp build)
    LIB_PATH = _ROOT / "_lib" / os.environ["MD_BUILD_TAG"] / "libmdiff.so"


class MdiffError(RuntimeError):
    pass


def _load():
    if not LIB_PATH.exists():
        raise MdiffError(
            f"{LIB_PATH} is missing: build it with `python -m morphablediffusion_b200.build` "
            "(or __graft_entry__.build()). There is no CPU / PyTorch fallback for the hot path.")
    return C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)


lib = _load()


class ConvGemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p),
        ("B", C.c_int), ("D", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("Cin", C.c_int), ("Cpitch", C.c_int),
        ("Wt", C.c_void_p),
        ("N", C.c_int),
        ("ntaps", C.c_int),
        ("tap", (C.c_int * 3) * 27),
        ("OD", C.c_int), ("OH", C.c_int), ("OW", C.c_int),
        ("os", C.c_int * 3), ("op", C.c_int * 3),
        ("bias", C.c_void_p),
        ("rowvec", C.c_void_p),
        ("rowvec_ld", C.c_int),
        ("res_f32", C.c_void_p),
        ("res_bf16", C.c_void_p),
        ("out_f32", C.c_void_p),
        ("out_bf16", C.c_void_p),
        ("ldo", C.c_int),
        ("act", C.c_int),
        ("out_scale", C.c_float),
        ("BN", C.c_int),
        ("col_stats", C.c_void_p),
        ("stats_ld", C.c_int),
        ("ksplit", C.c_int),
        ("in_stride", C.c_int * 3),
        ("cta_pair", C.c_int),
        ("Wpitch", C.c_int),
        ("gn_out", C.c_void_p), ("gn_gamma", C.c_void_p), ("gn_beta", C.c_void_p),
        ("gn_groups", C.c_int), ("gn_act", C.c_int), ("gn_eps", C.c_float), ("gn_barrier", C.c_void_p),
        ("tail_split", C.c_int),
    ]


lib.md_version.restype = C.c_int
lib.md_last_error.restype = C.c_char_p
lib.md_launch_count.restype = C.c_longlong
lib.md_reset_launch_count.restype = None
lib.md_op_conv_gemm.argtypes = [C.POINTER(ConvGemmArgs), C.c_void_p]
lib.md_op_conv_gemm.restype = C.c_int


def check(rc, what=""):
    if rc != 0:
        raise MdiffError(f"{what}: {lib.md_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "libmdiff takes contiguous CUDA tensors"
    return t.data_ptr()


def cur_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


ACT = {"none": 0, "silu": 1, "relu": 2, "geglu": 3, "gelu": 4, "quickgelu": 5}


def conv_gemm(A, Wt, *, B, D, H, W, Cin, N, taps, bias=None, rowvec=None, res_f32=None, res_bf16=None,
              out_f32=None, out_bf16=None, act="none", Cpitch=0, out_dims=None, os_=None, op=None, ldo=0,
              out_scale=1.0, BN=0, col_stats=None, in_stride=None, ksplit=0, cta_pair=0, tail_split=0, Wpitch=0, gn=None):
    a = ConvGemmArgs()
    a.A = (A.data_ptr() if Cpitch else ptr(A)); a.B, a.D, a.H, a.W = B, D, H, W
    a.Cin, a.Cpitch = Cin, Cpitch
    # Wpitch: Wt is a column slice of a wider row-major matrix (rows Wpitch elements apart); only its base address is used
    a.Wt = (Wt.data_ptr() if Wpitch else ptr(Wt)); a.N = N
    a.ntaps = len(taps)
    for i, t in enumerate(taps):
        for j in range(3):
            a.tap[i][j] = t[j]
    if out_dims:
        a.OD, a.OH, a.OW = out_dims
    if os_:
        for j in range(3):
            a.os[j] = os_[j]
    if op:
        for j in range(3):
            a.op[j] = op[j]
    a.bias = ptr(bias); a.rowvec = ptr(rowvec)
    a.rowvec_ld = rowvec.shape[-1] if rowvec is not None else 0
    a.res_f32 = ptr(res_f32); a.res_bf16 = ptr(res_bf16)
    a.out_f32 = ptr(out_f32); a.out_bf16 = ptr(out_bf16)
    a.ldo = ldo; a.act = ACT[act]; a.out_scale = out_scale; a.BN = BN
    a.col_stats = ptr(col_stats)
    a.ksplit = ksplit
    a.cta_pair = cta_pair
    a.tail_split = tail_split
    a.Wpitch = Wpitch
    if gn is not None:   # dict(out=bf16 tensor, gamma, beta, groups, eps, act, barrier=zeroed int32 tensor)
        a.gn_out = ptr(gn["out"]); a.gn_gamma = ptr(gn["gamma"]); a.gn_beta = ptr(gn["beta"])
        a.gn_groups = gn["groups"]; a.gn_eps = gn["eps"]; a.gn_act = ACT[gn.get("act", "none")]
        a.gn_barrier = ptr(gn["barrier"])
    if in_stride:
        for j in range(3):
            a.in_stride[j] = in_stride[j]
    check(lib.md_op_conv_gemm(C.byref(a), cur_stream()), "md_op_conv_gemm")


# ----------------------------------------------------------------------------- context-level API
class MdConfig(C.Structure):
    _fields_ = [
        ("model_channels", C.c_int), ("in_channels", C.c_int), ("out_channels", C.c_int),
        ("num_res_blocks", C.c_int), ("num_heads", C.c_int), ("context_dim", C.c_int),
        ("channel_mult", C.c_int * 4), ("attn_ds", C.c_int * 4), ("volume_dims", C.c_int * 4),
        ("latent_size", C.c_int), ("image_size", C.c_int), ("spatial_volume_size", C.c_int),
        ("frustum_depth", C.c_int), ("time_embed_dim", C.c_int), ("view_dim", C.c_int),
        ("spatial_volume_length", C.c_float), ("frustum_volume_length", C.c_float),
        ("smpl_num_views", C.c_int),
        ("ddim_steps", C.c_int), ("ddim_eta", C.c_float),
        ("max_views_per_call", C.c_int), ("workspace_bytes", C.c_ulonglong),
    ]


_vp = C.c_void_p
lib.md_default_config.argtypes = [C.POINTER(MdConfig)]
lib.md_default_config.restype = None
lib.md_create.argtypes = [C.POINTER(_vp), C.POINTER(MdConfig)]
lib.md_destroy.argtypes = [_vp]
lib.md_destroy.restype = None
lib.md_workspace_peak.argtypes = [_vp]
lib.md_workspace_peak.restype = C.c_ulonglong
lib.md_load_weights.argtypes = [_vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(_vp), C.POINTER(C.c_longlong), _vp]
lib.md_bind_sample.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]
lib.md_voxelize.argtypes = [_vp, C.c_int, _vp, _vp, _vp, _vp]
lib.md_affine_points.argtypes = [_vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), _vp, _vp]
lib.md_affine_points.restype = C.c_int
lib.md_images_to_u8.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp]
lib.md_images_to_u8.restype = C.c_int
lib.md_spatial_volume.argtypes = [_vp, _vp, _vp, _vp, _vp]
lib.md_embed_time.argtypes = [_vp, C.c_float, _vp, _vp]
lib.md_frustum_feats.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, C.POINTER(_vp), _vp]
lib.md_unet_forward.argtypes = [_vp, _vp, C.POINTER(C.c_float), _vp, C.POINTER(_vp), C.c_int, _vp, _vp]
lib.md_denoise_step.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.c_float, _vp, C.c_ulonglong, _vp, _vp]
lib.md_ddim_timestep.argtypes = [_vp, C.c_int]
lib.md_set_ddim.argtypes = [_vp, C.c_int, C.c_float]
lib.md_ddim_steps.argtypes = [_vp]
lib.md_vae_decode.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, _vp]
lib.md_has_vae.argtypes = [_vp]
lib.md_vae_encode.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, _vp]
lib.md_has_vae_encoder.argtypes = [_vp]
lib.md_clip_embed.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]
lib.md_has_clip.argtypes = [_vp]
lib.md_comm_unique_id.argtypes = [_vp]
lib.md_comm_init.argtypes = [_vp, C.c_int, C.c_int, _vp]
lib.md_peer_buffer.argtypes = [_vp, C.c_int, _vp]
lib.md_peer_attach.argtypes = [_vp, C.c_int, C.c_int, _vp]
lib.md_peer_attached.argtypes = [_vp]
lib.md_peer_detach.argtypes = [_vp]
for _f in ("md_peer_buffer", "md_peer_attach", "md_peer_attached", "md_peer_detach"):
    getattr(lib, _f).restype = C.c_int
for _f in ("md_embed_time", "md_create", "md_load_weights", "md_bind_sample", "md_voxelize", "md_spatial_volume", "md_frustum_feats",
           "md_unet_forward", "md_denoise_step", "md_ddim_timestep", "md_set_ddim", "md_ddim_steps", "md_comm_unique_id", "md_comm_init",
           "md_vae_decode", "md_has_vae", "md_vae_encode", "md_has_vae_encoder", "md_clip_embed", "md_has_clip"):
    getattr(lib, _f).restype = C.c_int

lib.md_op_group_norm.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _vp, _vp, _vp, C.c_int, _vp, _vp]
lib.md_op_group_norm_stats.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp]
lib.md_op_group_norm_stats.restype = C.c_int
lib.md_op_softmax_rows.argtypes = [_vp, _vp, C.c_longlong, C.c_int, _vp]
lib.md_op_softmax_rows.restype = C.c_int
lib.md_op_layer_norm.argtypes = [_vp, _vp, _vp, _vp, C.c_longlong, C.c_int, C.c_float, _vp]
lib.md_op_self_attention.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]
lib.md_op_self_attention_impl.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]
lib.md_op_depth_attention.argtypes = [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]
lib.md_op_cfg_ddim.argtypes = [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_ulonglong, C.c_int, _vp]
for _f in ("md_op_group_norm", "md_op_layer_norm", "md_op_self_attention", "md_op_self_attention_impl", "md_op_depth_attention",
           "md_op_cfg_ddim"):
    getattr(lib, _f).restype = C.c_int
