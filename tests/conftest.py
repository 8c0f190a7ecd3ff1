import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionfinish(session, exitstatus):
    """Writes the parity figures the GPU tests measured (tests/parity.py record()) next to the other GPU-run artefacts."""
    import json
    from tests import parity
    out = os.path.join(ROOT, "gpurun_out")
    tag = os.environ.get("PLTPT_PARITY_TAG", "")  # development: keeps the figures of a library variant (PLTPT_LIB) apart
    for name, data in ((tag + "parity_metrics.json", parity.METRICS), (tag + "image_parity.json", parity.IMAGE_PARITY)):
        if not data:
            continue
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, name)
        merged = {}
        if os.path.exists(path):
            try:
                with open(path) as f:
                    merged = json.load(f)
            except ValueError:
                merged = {}
        merged.update(data)
        with open(path, "w") as f:
            json.dump(merged, f, indent=1, sort_keys=True)


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def cornell():
    from pltfalcor_b200 import scene
    return scene.cornell_box()


@pytest.fixture(scope="session")
def sweep():
    from pltfalcor_b200 import scene
    return scene.sweep_scene()


@pytest.fixture(scope="session")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch
