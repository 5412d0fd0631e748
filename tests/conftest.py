import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def model_nograv():
    from helpers import load_model_asset
    return load_model_asset("final_humanoid_no_gravity")


@pytest.fixture(scope="session")
def model_grav():
    from helpers import load_model_asset
    return load_model_asset("final_humanoid")
