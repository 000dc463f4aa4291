import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _gpu_available():
    try:
        import nqcdynamics_jl_b200 as nq
        return nq._abi.load_engine_library().nqcb200_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Plain `pytest tests` on a CPU-only host: skip (not fail) the gpu-marked tests.  An explicit `-m gpu` run keeps
    them, so a GPU box without the library or without a device fails loudly."""
    if "gpu" in (config.getoption("-m") or "") or _gpu_available():
        return
    skip = pytest.mark.skip(reason="needs a B200 and nqcdynamics.jl_b200/csrc/libnqcb200.so")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
