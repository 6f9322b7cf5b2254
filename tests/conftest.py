import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: larger CPU-side checks")


@pytest.fixture(scope="session")
def golden():
    import json
    return json.loads((Path(__file__).parent / "golden" / "reference_outputs.json").read_text())
