import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_addoption(parser):
    parser.addoption("--checked", action="store_true", default=False,
                     help="load libsoap_b200_checked.so (make -C soap_b200/csrc checked: device-side bounds assertions)")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    if config.getoption("--checked"):
        from soap_b200 import _lib

        path = os.path.join(os.path.dirname(_lib.LIB_PATH), "libsoap_b200_checked.so")
        assert os.path.exists(path), f"{path} missing: make -C soap_b200/csrc checked"
        _lib.LIB_PATH = path


def pytest_report_header(config):
    from soap_b200 import _lib

    return f"soap_b200 library: {_lib.LIB_PATH}"


def pytest_collection_modifyitems(config, items):
    # GPU tests must never silently pass on a CPU box
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
