import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
# several band contexts of one process on one GPU (tests/test_gpu_bands_local.py) wait for each other inside kernels: their
# streams must not share a hardware queue (read by the CUDA runtime when it initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_usable():
    """(True, '') when libvrs.so is built and a CUDA device answers; the product has no CPU fallback to test instead."""
    try:
        import vrs_pkg
        V = vrs_pkg.load()
        R = V.Renderer(8, 8)
        R.destroy()
        return True, ""
    except Exception as e:              # VRS_ERR_NO_DEVICE, missing libvrs.so, ...
        return False, str(e)


def pytest_collection_modifyitems(config, items):
    gpu_items = [i for i in items if i.get_closest_marker("gpu")]
    if not gpu_items:
        return
    ok, why = _cuda_usable()
    if ok:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and a built libvrs.so: " + why)
    for i in gpu_items:
        i.add_marker(skip)


@pytest.fixture(scope="session")
def V():
    import vrs_pkg
    return vrs_pkg.load()


@pytest.fixture(scope="session")
def O():
    import oracle
    oracle.build()
    return oracle
