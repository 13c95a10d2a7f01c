"""Thin Python handle on an md_ctx (include/mdiff.h): one per GPU / rank.

Host-side glue only — tensors are torch CUDA tensors used as device memory; every computation happens inside
libmdiff.so.  The reference-shaped classes in morphablediffusion_b200/ldm_api.py are built on top of this handle.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _native as nat
from .spec import UNetConfig

PROJECTIONS = {"perspective": 0, "orthographic": 1}


def _np32(t):
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32))


class Engine:
    def __init__(self, unet_config=None, latent_size=32, image_size=256, ddim_steps=50, ddim_eta=1.0,
                 smpl_num_views=0, max_views_per_call=16, workspace_bytes=0, device=None):
        if not torch.cuda.is_available():
            raise nat.MdiffError("the Morphable Diffusion hot path needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._rank, self._world = 0, 1   # set by init_comm
        ucfg = unet_config or UNetConfig()
        cfg = nat.MdConfig()
        nat.lib.md_default_config(C.byref(cfg))
        cfg.model_channels, cfg.in_channels, cfg.out_channels = ucfg.model_channels, ucfg.in_channels, ucfg.out_channels
        cfg.num_res_blocks, cfg.num_heads, cfg.context_dim = ucfg.num_res_blocks, ucfg.num_heads, ucfg.context_dim
        for i in range(4):
            cfg.channel_mult[i] = ucfg.channel_mult[i]
            cfg.attn_ds[i] = 1 if (1 << i) in ucfg.attention_resolutions else 0
            cfg.volume_dims[i] = ucfg.volume_dims[i]
        cfg.latent_size, cfg.image_size = latent_size, image_size
        cfg.ddim_steps, cfg.ddim_eta = ddim_steps, ddim_eta
        cfg.smpl_num_views = smpl_num_views
        cfg.max_views_per_call = max_views_per_call
        cfg.workspace_bytes = workspace_bytes
        self.cfg = cfg
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            nat.check(nat.lib.md_create(C.byref(self._h), C.byref(cfg)), "md_create")
        self.S, self.V, self.D = latent_size, cfg.spatial_volume_size, cfg.frustum_depth
        self.n_views = self.view0 = self.n_local = 0
        self.ddim = (int(ddim_steps), float(ddim_eta))
        self.bind_count = 0   # md_bind_sample calls so far (tests check that edited batches re-bind)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            nat.lib.md_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd):
        """sd: {reference key: tensor}.  Tensors are moved to the GPU as fp32 for the duration of the call."""
        names, keep = [], []
        for k, v in sd.items():
            if not torch.is_tensor(v) or not v.dtype.is_floating_point:
                continue
            t = v.detach().to(self.device, torch.float32).contiguous()
            names.append(k.encode())
            keep.append(t)
        n = len(names)
        arr_n = (C.c_char_p * n)(*names)
        arr_p = (C.c_void_p * n)(*[t.data_ptr() for t in keep])
        arr_e = (C.c_longlong * n)(*[t.numel() for t in keep])
        nat.check(nat.lib.md_load_weights(self._h, n, arr_n, arr_p, arr_e, nat.cur_stream()), "md_load_weights")
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ per-sample binding
    def bind(self, batch, projection="perspective", view0=0, n_local=None, v_embed=None):
        """batch: the reference's batch dict (generate_face.py:227-241) with B = 1."""
        if projection not in PROJECTIONS:
            raise NotImplementedError(projection)
        K = _np32(batch["target_K"][0])
        RT = _np32(batch["target_RT"][0])
        n_views = K.shape[0]
        n_local = n_views - view0 if n_local is None else n_local
        if v_embed is None:
            v_embed = viewpoint_embedding(batch)[0]
        ve = _np32(v_embed)
        verts = _np32(batch["vertices"][0])
        coord = np.ascontiguousarray(batch["coord"][0].detach().cpu().numpy().astype(np.int32))
        out_sh = np.ascontiguousarray(batch["out_sh"][0].detach().cpu().numpy().astype(np.int32))
        bounds = _np32(batch["bounds"][0])
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        nat.check(nat.lib.md_bind_sample(self._h, p(K), p(RT), p(ve), p(verts), p(coord), p(out_sh), p(bounds),
                                         verts.shape[0], n_views, view0, n_local, PROJECTIONS[projection],
                                         nat.cur_stream()), "md_bind_sample")
        self.n_views, self.view0, self.n_local = n_views, view0, n_local
        self.bind_count += 1

    # ------------------------------------------------------------------ stages
    def embed_time(self, timestep):
        out = torch.empty(self.cfg.time_embed_dim, device=self.device)
        nat.check(nat.lib.md_embed_time(self._h, float(timestep), out.data_ptr(), nat.cur_stream()), "md_embed_time")
        return out

    def _t_embed(self, t):
        if torch.is_tensor(t) and t.numel() == self.cfg.time_embed_dim:
            return t.to(self.device, torch.float32).reshape(-1).contiguous()
        return self.embed_time(float(t))

    def spatial_volume(self, x_local, t_embed):
        """t_embed: the [256] embedding (embed_time) or a scalar timestep."""
        x = x_local.to(self.device, torch.float32).contiguous()
        te = self._t_embed(t_embed)
        out = torch.empty(1, 64, self.V, self.V, self.V, device=self.device)
        nat.check(nat.lib.md_spatial_volume(self._h, x.data_ptr(), te.data_ptr(), out.data_ptr(), nat.cur_stream()),
                  "md_spatial_volume")
        return out

    def frustum_feats(self, volume, lv0, T, t_embed):
        vol = volume.to(self.device, torch.float32).contiguous()
        te = self._t_embed(t_embed)
        S, D = self.S, self.D
        outs = [torch.empty(T, self.cfg.volume_dims[i], D >> i, S >> i, S >> i, device=self.device) for i in range(4)]
        arr = (C.c_void_p * 4)(*[o.data_ptr() for o in outs])
        nat.check(nat.lib.md_frustum_feats(self._h, vol.data_ptr(), lv0, T, te.data_ptr(), arr, nat.cur_stream()),
                  "md_frustum_feats")
        return {S >> i: outs[i] for i in range(4)}

    def unet_forward(self, x, timesteps, context, source_dict):
        B = x.shape[0]
        S = self.S
        x = x.to(self.device, torch.float32).contiguous()
        ctx = context.to(self.device, torch.float32).reshape(B, -1).contiguous()
        src = [source_dict[S >> i].to(self.device, torch.float32).contiguous() for i in range(4)]
        ts = (C.c_float * B)(*[float(t) for t in timesteps.detach().cpu().tolist()])
        arr = (C.c_void_p * 4)(*[s.data_ptr() for s in src])
        out = torch.empty(B, self.cfg.out_channels, S, S, device=self.device)
        nat.check(nat.lib.md_unet_forward(self._h, x.data_ptr(), ts, ctx.data_ptr(), arr, B, out.data_ptr(),
                                          nat.cur_stream()), "md_unet_forward")
        return out

    def denoise_step(self, x_local, x_input, clip_embed, index, cfg_scale, noise=None, seed=0, want_eps=False):
        """In-place DDIM step on x_local [n_local,4,S,S] (CUDA fp32 contiguous)."""
        assert x_local.is_cuda and x_local.dtype == torch.float32 and x_local.is_contiguous()
        eps = torch.empty_like(x_local) if want_eps else None
        nat.check(nat.lib.md_denoise_step(self._h, x_local.data_ptr(), x_input.data_ptr(), clip_embed.data_ptr(),
                                          int(index), float(cfg_scale), nat.ptr(noise), int(seed), nat.ptr(eps),
                                          nat.cur_stream()), "md_denoise_step")
        return eps

    def has_vae(self):
        """True when the loaded state dict carried first_stage_model.decoder.* / post_quant_conv.*."""
        return bool(nat.lib.md_has_vae(self._h))

    def vae_decode(self, x):
        """decode_first_stage (morphable_diffusion.py:468-471): x [n,4,S,S] scaled latents -> images [n,3,8S,8S]."""
        x = x.to(self.device, torch.float32).contiguous()
        n, _, S, _ = x.shape
        out = torch.empty(n, 3, 8 * S, 8 * S, device=self.device)
        nat.check(nat.lib.md_vae_decode(self._h, x.data_ptr(), out.data_ptr(), n, S, nat.cur_stream()), "md_vae_decode")
        return out

    def has_vae_encoder(self):
        return bool(nat.lib.md_has_vae_encoder(self._h))

    def vae_encode_moments(self, image):
        """AutoencoderKL.encode up to the moments: image [n,3,8S,8S] in [-1,1] -> [n,8,S,S] (mean | logvar)."""
        x = image.to(self.device, torch.float32).contiguous()
        n, _, H, _ = x.shape
        out = torch.empty(n, 8, H // 8, H // 8, device=self.device)
        nat.check(nat.lib.md_vae_encode(self._h, x.data_ptr(), out.data_ptr(), n, H // 8, nat.cur_stream()), "md_vae_encode")
        return out

    def has_clip(self):
        return bool(nat.lib.md_has_clip(self._h))

    def clip_embed(self, image):
        """FrozenCLIPImageEmbedder.encode: image [n,3,H,W] in [-1,1] -> [n,1,768]."""
        x = image.to(self.device, torch.float32).contiguous()
        n, _, H, W = x.shape
        out = torch.empty(n, 768, device=self.device)
        nat.check(nat.lib.md_clip_embed(self._h, x.data_ptr(), out.data_ptr(), n, H, W, nat.cur_stream()), "md_clip_embed")
        return out.unsqueeze(1)

    def set_ddim(self, ddim_steps, ddim_eta=1.0):
        """Schedule of SyncDDIMSampler(model, ddim_steps, ddim_eta=...) (morphable_diffusion.py:649-672)."""
        if (int(ddim_steps), float(ddim_eta)) != self.ddim:
            nat.check(nat.lib.md_set_ddim(self._h, int(ddim_steps), float(ddim_eta)), "md_set_ddim")
            self.ddim = (int(ddim_steps), float(ddim_eta))

    def ddim_timestep(self, index):
        return nat.lib.md_ddim_timestep(self._h, index)

    def workspace_peak(self):
        return nat.lib.md_workspace_peak(self._h)

    # ------------------------------------------------------------------ multi-GPU
    def init_comm(self, rank, world, unique_id_bytes):
        buf = C.create_string_buffer(bytes(unique_id_bytes), 128)
        nat.check(nat.lib.md_comm_init(self._h, rank, world, buf), "md_comm_init")
        self._rank, self._world = rank, world

    def init_peer_exchange(self, dist):
        """NVLink peer exchange for the step's cross-rank sum (md_peer_buffer / md_peer_attach, include/mdiff.h): call on
        every rank after init_comm; `dist` is an initialised torch.distributed (handles travel by all_gather_object,
        a barrier closes the setup).  Afterwards denoise_step makes no NCCL call."""
        rank, world = self._rank, self._world
        if world <= 1:
            return False
        h = C.create_string_buffer(64)
        err = None
        try:
            nat.check(nat.lib.md_peer_buffer(self._h, world, h), "md_peer_buffer")
        except nat.MdiffError as e:
            err = str(e)
        handles = [None] * world
        dist.all_gather_object(handles, (h.raw, err))
        if err is None and all(e is None for _, e in handles):
            buf = C.create_string_buffer(b"".join(raw for raw, _ in handles), 64 * world)
            try:
                if os.environ.get("MD_PEER_TEST_FAIL") == str(rank):  # test hook: this rank pretends it cannot map its peers
                    raise nat.MdiffError("md_peer_attach: simulated failure (MD_PEER_TEST_FAIL)")
                nat.check(nat.lib.md_peer_attach(self._h, rank, world, buf), "md_peer_attach")
            except nat.MdiffError as e:
                err = str(e)
        # the exchange is collective: one rank that cannot map its peers (e.g. a launcher that hides the other devices
        # from each process) sends everyone back to the NCCL all-reduce
        errs = [None] * world
        dist.all_gather_object(errs, err)
        errs = [e for _, e in handles if e is not None] + [e for e in errs if e is not None]
        if errs:
            nat.check(nat.lib.md_peer_detach(self._h), "md_peer_detach")
            if rank == 0:
                import warnings
                warnings.warn("NVLink peer exchange unavailable, using the NCCL all-reduce: " + errs[0])
        dist.barrier()
        return not errs

    def peer_exchange_attached(self):
        return bool(nat.lib.md_peer_attached(self._h))


def comm_unique_id():
    buf = C.create_string_buffer(128)
    nat.check(nat.lib.md_comm_unique_id(buf), "md_comm_unique_id")
    return buf.raw


def viewpoint_embedding(batch):
    """get_viewpoint_embedding (morphable_diffusion.py:383-397): [d_elev, sin d_az, cos d_az, 0] in radians."""
    d_e = torch.deg2rad(batch["target_elevation"]) - torch.deg2rad(batch["input_elevation"])
    d_a = torch.deg2rad(batch["target_azimuth"]) - torch.deg2rad(batch["input_azimuth"])
    return torch.stack([d_e, torch.sin(d_a), torch.cos(d_a), torch.zeros_like(d_a)], -1)


def voxelize(vertices):
    """GPU voxelisation (md_voxelize). vertices [Nv,3] CUDA fp32 -> coord i32 [Nv,3], out_sh i32 [3], bounds [2,3]."""
    v = vertices.contiguous()
    coord = torch.empty(v.shape[0], 3, dtype=torch.int32, device=v.device)
    out_sh = torch.empty(3, dtype=torch.int32, device=v.device)
    bounds = torch.empty(2, 3, dtype=torch.float32, device=v.device)
    nat.check(nat.lib.md_voxelize(v.data_ptr(), v.shape[0], coord.data_ptr(), out_sh.data_ptr(), bounds.data_ptr(),
                                  nat.cur_stream()), "md_voxelize")
    return coord, out_sh, bounds
