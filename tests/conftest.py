import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


@pytest.fixture(scope="session")
def built_lib():
    from morphablediffusion_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def state_dict():
    from morphablediffusion_b200 import synth
    return synth.make_state_dict()


@pytest.fixture(scope="session")
def vae_state_dict():
    from morphablediffusion_b200 import synth
    return synth.make_vae_state_dict()


@pytest.fixture(scope="session")
def vae_encoder_state_dict():
    from morphablediffusion_b200 import synth
    return synth.make_vae_encoder_state_dict()


@pytest.fixture(scope="session")
def clip_state_dict():
    from morphablediffusion_b200 import synth
    return synth.make_clip_state_dict()
