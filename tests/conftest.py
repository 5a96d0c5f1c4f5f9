import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
CONFIGS = {"flowers_sd": 102, "midi_vqgan": 0, "stl_sd": 10}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200, sm_100a) device")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="session")
def goldens():
    return {name: torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False) for name in CONFIGS}


def seeded_state_dict(n_classes, seed=1234):
    """Random-init weights of the named U-Net, regenerated from the seed by OUR module (whose leaf
    construction order mirrors the reference's, so the RNG stream lines up)."""
    from flocoder_b200.unet import Unet
    torch.manual_seed(seed)
    m = Unet(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=n_classes)
    return m, {k: v.detach().clone() for k, v in m.state_dict().items()}
