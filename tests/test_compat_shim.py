"""The drop-in import shim (morphablediffusion_b200/compat) with the reference checkout BEHIND it on sys.path — the
arrangement INTEGRATION.md §1 documents for running generate_face.py / train_morphable_diffusion.py unchanged.

Runs in a subprocess (it installs a package named `ldm` into sys.modules).  Needs /root/reference, so it is skipped
on the GPU box; nothing here touches a GPU."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

SCRIPT = textwrap.dedent(r"""
    import json, os, sys, types
    ROOT, REF = sys.argv[1], sys.argv[2]
    sys.dont_write_bytecode = True
    sys.path[:0] = [os.path.join(ROOT, "morphablediffusion_b200", "compat"), ROOT, REF]
    import torch, torch.nn as nn, yaml

    # third-party packages the reference's NON-hot-path modules import and this image lacks: inert stand-ins
    def stub(name, **attrs):
        m = types.ModuleType(name); m.__dict__.update(attrs); sys.modules[name] = m; return m
    class LightningModule(nn.Module):
        pass
    stub("pytorch_lightning", LightningModule=LightningModule)
    stub("taming"); stub("taming.modules"); stub("taming.modules.vqvae")
    stub("taming.modules.vqvae.quantize", VectorQuantizer2=object)
    for n in ("matplotlib", "matplotlib.pyplot", "kornia", "clip", "skimage", "skimage.io"):
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                stub(n, imread=None)

    out = {}
    import ldm
    out["ldm_file"] = ldm.__file__
    out["ref_roots"] = ldm.REFERENCE_LDM
    from ldm.models.diffusion.morphable_diffusion import SyncMultiviewDiffusion, SyncDDIMSampler
    from ldm.util import instantiate_from_config
    import morphablediffusion_b200.ldm_api as api
    out["shadowed"] = SyncMultiviewDiffusion is api.SyncMultiviewDiffusion and SyncDDIMSampler is api.SyncDDIMSampler
    # modules that are NOT on the hot path still come from the reference
    import ldm.modules.diffusionmodules.model as vae_model
    import ldm.modules.distributions.distributions as dists
    import ldm.lr_scheduler as lrs
    import ldm.models.autoencoder as ae
    out["vae_from_reference"] = vae_model.__file__.startswith(REF) and ae.__file__.startswith(REF)
    out["misc_from_reference"] = dists.__file__.startswith(REF) and lrs.__file__.startswith(REF)
    # names of shadowed modules beyond the hot-path classes fall through to the reference's own file
    from ldm.util import default, exists
    out["util_fallthrough"] = default(None, 3) == 3 and exists(1)
    from ldm.modules.diffusionmodules.openaimodel import ResBlock, UNetModel
    out["openai_fallthrough"] = ResBlock.__module__.endswith("_reference_openaimodel") and UNetModel is api.UNetModel

    # generate_face.py:71-78: config -> model -> load_state_dict(strict=False)
    cfg = yaml.safe_load(open(os.path.join(REF, "configs", "facescape.yaml")))
    model = instantiate_from_config(cfg["model"])
    out["model_class"] = type(model).__module__ + "." + type(model).__name__
    out["first_stage"] = type(model.first_stage_model).__module__ + "." + type(model.first_stage_model).__name__
    out["clip_error"] = model._clip_error
    spec = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_state_dict_spec.json")))
    sd = {k: torch.zeros(shape) for k, shape in spec.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    out["unexpected"] = list(unexpected)
    out["missing_hot_path"] = [k for k in missing if k.startswith(("model.", "spatial_volume.", "time_embed."))]
    out["n_first_stage_keys"] = sum(k.startswith("first_stage_model.") for k in model.state_dict())
    sampler = SyncDDIMSampler(model, 20)     # generate_face.py:243
    out["sampler_steps"] = int(len(sampler.ddim_timesteps))
    print("RESULT " + json.dumps(out))
""")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ldm")), reason="needs the reference checkout")
def test_compat_shim_with_reference_behind_it(tmp_path):
    script = tmp_path / "compat_probe.py"
    script.write_text(SCRIPT)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    env.pop("PYTHONPATH", None)
    p = subprocess.run([sys.executable, str(script), ROOT, REF], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    out = json.loads(line[len("RESULT "):])
    assert out["ldm_file"].startswith(os.path.join(ROOT, "morphablediffusion_b200", "compat"))
    assert out["ref_roots"] and out["ref_roots"][0].startswith(REF)
    assert out["shadowed"] and out["vae_from_reference"] and out["misc_from_reference"]
    assert out["util_fallthrough"] and out["openai_fallthrough"]
    assert out["model_class"] == "morphablediffusion_b200.ldm_api.SyncMultiviewDiffusion"
    # the frozen VAE is the reference's own class, built by the shell like morphable_diffusion.py:399-425 does
    assert out["first_stage"] == "ldm.models.autoencoder.AutoencoderKL" and out["n_first_stage_keys"] > 100
    assert out["unexpected"] == [] and out["missing_hot_path"] == []
    assert out["sampler_steps"] == 20


def test_shim_alone_still_serves_the_hot_path_classes(tmp_path):
    """Without any reference checkout the shim still resolves the four dotted names the configs use."""
    code = textwrap.dedent(f"""
        import sys
        sys.path[:0] = [{os.path.join(ROOT, 'morphablediffusion_b200', 'compat')!r}, {ROOT!r}]
        sys.path = [p for p in sys.path if not p.rstrip('/').endswith('reference')]
        from ldm.util import instantiate_from_config, get_obj_from_str
        from ldm.models.diffusion.attention import DepthWiseAttention
        from ldm.modules.diffusionmodules.openaimodel import UNetModel
        import ldm
        assert ldm.REFERENCE_LDM == [], ldm.REFERENCE_LDM
        assert get_obj_from_str("ldm.models.diffusion.attention.DepthWiseAttention") is DepthWiseAttention
        try:
            from ldm.util import default
        except ImportError:
            print("OK")
    """)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    env.pop("PYTHONPATH", None)
    env.pop("MD_REFERENCE_ROOT", None)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd=str(tmp_path))
    assert p.returncode == 0 and "OK" in p.stdout, p.stderr[-2000:]
