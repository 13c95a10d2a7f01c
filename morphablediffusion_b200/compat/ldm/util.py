"""ldm.util: `instantiate_from_config` / `get_obj_from_str` (reference ldm/util.py:217-232) are the B200 shell's (the
hot-path dotted names resolve to the B200 classes even when a config is instantiated before anything was imported);
every other name of the reference module (default, exists, count_params, log_txt_as_img, ...) falls through to the
reference's own file."""
from morphablediffusion_b200.ldm_api import get_obj_from_str, instantiate_from_config  # noqa: F401


def __getattr__(name):
    from ldm import _reference_module
    ref = _reference_module("util.py", "ldm._reference_util")
    if ref is not None and hasattr(ref, name):
        return getattr(ref, name)
    raise AttributeError(f"module 'ldm.util' has no attribute {name!r}")
