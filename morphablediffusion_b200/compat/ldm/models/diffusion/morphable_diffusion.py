"""ldm.models.diffusion.morphable_diffusion (reference :67-776) -> B200 implementations."""
from morphablediffusion_b200.ldm_api import (SpatialVolumeNet, SyncDDIMSampler, SyncMultiviewDiffusion,  # noqa: F401
                                             UNetWrapper)
