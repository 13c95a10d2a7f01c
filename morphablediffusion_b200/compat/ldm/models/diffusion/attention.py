"""ldm.models.diffusion.attention (reference :87-142) -> B200 implementation."""
from morphablediffusion_b200.ldm_api import DepthWiseAttention  # noqa: F401
