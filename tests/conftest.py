import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOOD_MEAN = [0.14245495, 0.13921481, 0.12434631, 0.31420089, 0.20743526, 0.12046503]
FLOOD_STD = [0.04036231, 0.04186983, 0.05267646, 0.0822221, 0.06834774, 0.05294205]
CROP_MEAN = [494.905781, 815.239594, 924.335066, 2968.881459, 2634.621962, 1739.579917]
CROP_STD = [284.925432, 357.84876, 575.566823, 896.601013, 951.900334, 921.407808]
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests must never silently pass on a box without a GPU: they fail there unless deselected."""
    return


@pytest.fixture(scope="session")
def cuda_dev():
    import torch

    assert torch.cuda.is_available(), "GPU test collected on a machine without CUDA (use -m 'not gpu')"
    import instageo_b200

    instageo_b200._lib.load()  # fail loudly if the extension is missing
    return torch.device("cuda:0")
