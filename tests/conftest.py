import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the shared libraries are build artefacts (git-ignored): build them when a fresh checkout runs the tests
    import dasp_b200
    from dasp_b200 import synth  # noqa: F401

    if not (os.path.exists(dasp_b200.library_path())
            and os.path.exists(os.path.join(os.path.dirname(dasp_b200.library_path()), "libdasp_synth.so"))):
        dasp_b200.build()


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("test marked gpu but no CUDA device is visible (there is no CPU fallback)")
    torch.cuda.init()
    return torch.device("cuda:0")
