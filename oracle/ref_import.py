"""ORACLE tooling (build-container only): import the REAL reference modules from /root/reference.

Nothing here travels to the GPU box (the reference tree does not exist there); it is used by
oracle/make_golden.py to produce tests/golden/* and by tests that are skipped when /root/reference is absent.

The reference needs packages that are not installed (pytorch_lightning, spconv, kornia, omegaconf, trimesh, clip,
skimage, matplotlib).  None of them does arithmetic on the hot path except spconv; they are replaced by inert
stubs in sys.modules.  spconv is replaced by a dense-grid stand-in (oracle/spconv_dense.py) — see the
"parity unpinned" note there.
"""
import os
import sys
import types

REF_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "ldm"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    import torch
    import torch.nn as nn

    sys.dont_write_bytecode = True
    if "matplotlib" not in sys.modules:
        _stub("matplotlib")
        _stub("matplotlib.pyplot")
    _stub("omegaconf", OmegaConf=object)
    _stub("omegaconf.listconfig", ListConfig=type("ListConfig", (list,), {}))

    class LightningModule(nn.Module):
        @property
        def device(self):
            return next(self.parameters()).device

        def log(self, *a, **k):
            pass

    pl = _stub("pytorch_lightning", LightningModule=LightningModule)
    _stub("pytorch_lightning.utilities", rank_zero_only=lambda f: f)
    _stub("pytorch_lightning.utilities.distributed", rank_zero_only=lambda f: f)
    pl.utilities = sys.modules["pytorch_lightning.utilities"]
    _stub("skimage")
    _stub("skimage.io", imsave=lambda *a, **k: None, imread=lambda *a, **k: None)
    sys.modules["skimage"].io = sys.modules["skimage.io"]
    _stub("trimesh")
    _stub("clip")

    def create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
        xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
        ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
        if normalized_coordinates:
            xs = (xs / (width - 1) - 0.5) * 2
            ys = (ys / (height - 1) - 0.5) * 2
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        return torch.stack([gx, gy], -1).unsqueeze(0)

    _stub("kornia", create_meshgrid=create_meshgrid)

    from oracle import spconv_dense as sp
    _stub("spconv")
    _stub("spconv.pytorch")
    _stub("spconv.pytorch.core", SparseConvTensor=sp.SparseConvTensor)
    _stub("spconv.pytorch.conv", SparseConv3d=sp.SparseConv3d, SubMConv3d=sp.SubMConv3d)
    _stub("spconv.pytorch.modules", SparseSequential=sp.SparseSequential)

    # ldm.modules.encoders.modules pulls in clip/kornia/transformers heavy imports: only the class name is needed.
    enc = _stub("ldm.modules.encoders.modules", FrozenCLIPImageEmbedder=object)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return enc


_installed = False


def reference():
    """Returns a namespace with the reference classes on the hot path."""
    global _installed
    if not available():
        raise RuntimeError("/root/reference is not present (GPU box?)")
    if not _installed:
        install_stubs()
        _installed = True
    import importlib
    ns = types.SimpleNamespace()
    ns.md = importlib.import_module("ldm.models.diffusion.morphable_diffusion")
    ns.attention = importlib.import_module("ldm.models.diffusion.attention")
    ns.network = importlib.import_module("ldm.models.diffusion.network")
    ns.utils = importlib.import_module("ldm.models.diffusion.utils")
    ns.openai = importlib.import_module("ldm.modules.diffusionmodules.openaimodel")
    return ns


UNET_PARAMS = dict(volume_dims=[64, 128, 256, 512], image_size=32, in_channels=8, out_channels=4,
                   model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
                   channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True, transformer_depth=1,
                   context_dim=768, use_checkpoint=False, legacy=False)


def build_reference_model(projection="perspective", view_num=16, cfg_scale=2.0, sample_steps=50, latent=32):
    """A SyncMultiviewDiffusion with everything on the per-step path constructed by the reference's own code and
    the frozen side models (VAE, CLIP — outside the step loop, no checkpoint here) skipped."""
    import torch.nn as nn
    ns = reference()
    md = ns.md
    m = md.SyncMultiviewDiffusion.__new__(md.SyncMultiviewDiffusion)
    nn.Module.__init__(m)
    m.view_num = view_num
    m.viewpoint_dim = 4
    m.image_size = latent * 8
    m.cfg_scale = cfg_scale
    m._init_time_step_embedding()
    m._init_schedule()
    # input_image_size is a constructor argument of the reference's SpatialVolumeNet (morphable_diffusion.py:152-157);
    # SyncMultiviewDiffusion.__init__ never passes it (:351), which is why the stock model is tied to 32x32 latents
    m.spatial_volume = md.SpatialVolumeNet(m.time_embed_dim, m.viewpoint_dim, m.view_num, input_image_size=latent * 8,
                                           projection=projection, use_spatial_volume=False)
    m.spatial_volume.smpl_feature_extractor.num_views = view_num
    unet_config = {"target": "ldm.models.diffusion.attention.DepthWiseAttention",
                   "params": dict(UNET_PARAMS, image_size=latent)}
    m.model = md.UNetWrapper(unet_config)
    m.sampler = md.SyncDDIMSampler(m, sample_steps, "uniform", 1.0, latent_size=latent)
    m.eval()
    return m, ns
