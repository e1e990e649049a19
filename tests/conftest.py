import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def port():
    from oracle import cpu

    if not cpu.available("port"):
        cpu.build("port")
    return cpu.CpuLib("port")


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference CPU library (prebuilt oracle/_ref travels to the GPU box)."""
    from oracle import cpu

    if not cpu.available("reference"):
        try:
            cpu.build("reference")
        except Exception:
            pass
    if not cpu.available("reference"):
        pytest.skip("oracle/_ref/libdeepmd_ref.so not built (needs /root/reference)")
    return cpu.CpuLib("reference")


@pytest.fixture(scope="session", params=["port", "reference"])
def olib(request):
    """CPU checker of the GPU operator tests, both flavours: our restatement (`port`, oracle/dp_oracle.c) and the
    unmodified reference CPU library compiled from /root/reference (`reference`, oracle/_ref; prebuilt, travels to
    the GPU box).  Every GPU parity test runs against both."""
    from oracle import cpu

    kind = request.param
    if not cpu.available(kind):
        try:
            cpu.build(kind)
        except Exception:
            pass
    if not cpu.available(kind):
        pytest.skip(f"CPU checker `{kind}` is not built")
    return cpu.CpuLib(kind)
