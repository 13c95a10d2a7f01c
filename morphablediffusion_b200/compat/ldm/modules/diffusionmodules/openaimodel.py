"""ldm.modules.diffusionmodules.openaimodel (reference :414-777) -> B200 implementation."""
from morphablediffusion_b200.ldm_api import UNetModel  # noqa: F401
