"""ldm.models.diffusion.attention: DepthWiseAttention (reference :87-142) is the B200 shell class; DepthAttention /
DepthTransformer fall through to the reference file."""
import os

from morphablediffusion_b200.ldm_api import DepthWiseAttention  # noqa: F401


def __getattr__(name):
    from ldm import _reference_module
    ref = _reference_module(os.path.join("models", "diffusion", "attention.py"),
                            "ldm.models.diffusion._reference_attention")
    if ref is not None and hasattr(ref, name):
        return getattr(ref, name)
    raise AttributeError(f"module 'ldm.models.diffusion.attention' has no attribute {name!r}")
