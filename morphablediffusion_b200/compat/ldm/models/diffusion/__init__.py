from ldm import _extend_path

_extend_path(__path__, "models/diffusion")
