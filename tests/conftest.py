import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def calibrated_weights():
    """Seeded DenseNet-121 U-Net weights with BN statistics calibrated by one oracle pass (patch 256)."""
    import numpy as np
    from digipathai_b200.models.densenet import init_densenet_weights
    from oracle import densenet_ref
    rng = np.random.default_rng(1)
    tiles = rng.integers(0, 256, (2, 256, 256, 3)).astype(np.uint8)
    w = init_densenet_weights(0)
    densenet_ref.calibrate_bn(w, (tiles.astype(np.float32) - 128.0) / 128.0)
    return w, tiles
