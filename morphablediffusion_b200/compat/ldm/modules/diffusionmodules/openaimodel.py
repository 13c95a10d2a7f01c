"""ldm.modules.diffusionmodules.openaimodel: UNetModel (reference :414-777) is the B200 shell class; the building
blocks the reference file also defines (ResBlock, Upsample, TimestepEmbedSequential, ...) fall through to it."""
from morphablediffusion_b200.ldm_api import UNetModel  # noqa: F401


def __getattr__(name):
    from ldm import _reference_module
    ref = _reference_module(os.path.join("modules", "diffusionmodules", "openaimodel.py"),
                            "ldm.modules.diffusionmodules._reference_openaimodel")
    if ref is not None and hasattr(ref, name):
        return getattr(ref, name)
    raise AttributeError(f"module 'ldm.modules.diffusionmodules.openaimodel' has no attribute {name!r}")


import os  # noqa: E402
