import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cuda_ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from marie_icr_b200._lib import Context
    return Context.get(0)


@pytest.fixture(params=["fp16", "bf16"])
def dtype16(request, cuda_ctx):
    """Runs the test once per 16-bit element type of the library; restores the fp16 default afterwards."""
    import torch
    cuda_ctx.set_dtype(request.param)
    yield torch.float16 if request.param == "fp16" else torch.bfloat16
    cuda_ctx.set_dtype("fp16")
