"""Import-path shim: put `morphablediffusion_b200/compat` AHEAD of the reference checkout on sys.path and the
reference's own scripts (generate_face.py:11, eval/generate_all_facescape.py:14, train_morphable_diffusion.py) import
the B200 classes of the per-step hot path unchanged, while every other `ldm.*` module (ldm.models.autoencoder,
ldm.modules.encoders.modules, ldm.modules.diffusionmodules.model, ldm.base_utils, ldm.data.*, ...) still resolves to
the reference: this package and its sub-packages extend their `__path__` onto the reference's `ldm/` tree, found as
the next `ldm/` directory on sys.path or under $MD_REFERENCE_ROOT.

Shadowed (B200 implementations): ldm.models.diffusion.morphable_diffusion, ldm.models.diffusion.attention,
ldm.modules.diffusionmodules.openaimodel (UNetModel), ldm.util.instantiate_from_config.  Names those reference modules
define beyond the hot-path classes fall through to the reference file (module __getattr__)."""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def _reference_roots():
    """`ldm/` directories of reference checkouts visible to this process, in priority order."""
    roots = []
    env = os.environ.get("MD_REFERENCE_ROOT")
    cands = [os.path.join(env, "ldm")] if env else []
    cands += [os.path.join(p or ".", "ldm") for p in sys.path]
    for d in cands:
        d = os.path.abspath(d)
        if d != _HERE and d not in roots and os.path.isfile(os.path.join(d, "util.py")) and \
                os.path.isdir(os.path.join(d, "models")):
            roots.append(d)
    return roots


REFERENCE_LDM = _reference_roots()


def _extend_path(pkg_path, rel):
    """Append <reference ldm>/<rel> to a sub-package's __path__ (rel = '' for this package)."""
    for r in REFERENCE_LDM:
        d = os.path.join(r, rel) if rel else r
        if os.path.isdir(d) and d not in pkg_path:
            pkg_path.append(d)


def _reference_module(rel_file, private_name):
    """Execute the reference's own copy of a module this shim shadows (e.g. 'util.py') under a private module name and
    return it; None when no reference checkout is visible."""
    if private_name in sys.modules:
        return sys.modules[private_name]
    for r in REFERENCE_LDM:
        path = os.path.join(r, rel_file)
        if os.path.isfile(path):
            spec = importlib.util.spec_from_file_location(private_name, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[private_name] = mod
            try:
                spec.loader.exec_module(mod)
            except BaseException:
                sys.modules.pop(private_name, None)
                raise
            return mod
    return None


_extend_path(__path__, "")
