import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        # the GPU tests go through libdml_b200.so: if the snapshot came without the built library, build it (nvcc is in
        # the image) rather than fail every test at load time - there is still no fallback path
        from dml_b200 import _lib
        if not os.path.exists(_lib.LIB_PATH):
            import __graft_entry__
            __graft_entry__.build()
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(autouse=True)
def _exact_fp32_reference_maths():
    """torch's own GPU convolutions / matmuls default to TF32; the comparison maths in the tests must be
    exact fp32 (the product path runs its contractions on its own pair GEMM, not on torch's)."""
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    # oneDNN's fp32 depthwise-conv weight gradient is wrong by 5-8 % for the [1, 8, 256, d] res_conv shape
    # (torch 2.11 CPU); the goldens were made with it off (oracle/make_goldens.py) and so is the CPU oracle here.
    with torch.backends.mkldnn.flags(enabled=False):
        yield
