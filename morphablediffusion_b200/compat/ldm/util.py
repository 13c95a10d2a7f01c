"""ldm.util (reference ldm/util.py:217-232): the config -> class plugin mechanism."""
from morphablediffusion_b200.ldm_api import get_obj_from_str, instantiate_from_config  # noqa: F401
