"""pytest configuration: markers, import path, and a one-time in-tree build of the native library."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")
    # Build the C-ABI library (nvcc cross-compiles without a GPU) and the C oracle if they are stale,
    # so both test tiers see the current sources.  Building the checker is not using it.
    import importlib.util

    spec = importlib.util.spec_from_file_location("_msda_native_build", os.path.join(ROOT, "co-detr-tensorrt_b200", "_native.py"))
    native = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(native)
    native.build_native()
    import oracle

    oracle.build()


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
