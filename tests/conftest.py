import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.fixture(scope="session")
def ctx(torch_cuda):
    """one library context for the session, on torch's current stream"""
    import min_llm_inference_b200 as mli
    torch = torch_cuda
    c = mli.Context(0, torch.cuda.current_stream().cuda_stream)
    yield c
    c.close()


@pytest.fixture(scope="session")
def ref(torch_cuda):
    import harness
    if not harness.ref_available():
        pytest.skip("oracle/_ref/libmli_ref.so not built")
    return harness.load_ref()
