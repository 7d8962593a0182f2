import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _have_gpu() -> bool:
    try:
        from cpu_raymarcher_b200 import _lib
        return _lib.lib().rm_device_count() > 0
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle as po
    po.lib()
    return po


@pytest.fixture(scope="session")
def val_worker():
    import cpu_raymarcher_b200 as rb
    w = rb.RaymarchWorker(device=0, validate_fp64=True)
    yield w
    w.close()


@pytest.fixture(scope="session")
def fast_worker():
    import cpu_raymarcher_b200 as rb
    w = rb.RaymarchWorker(device=0, validate_fp64=False)
    yield w
    w.close()


def make_job(W, H, preset=0, accel="None", alg="sphere-tracer", pitch=0.0, yaw=0.0, y0=0, y1=None, synthetic=None,
             step=0.1, over=1.2, time=0.0):
    return dict(width=W, height=H, time=time, yStart=y0, yEnd=H if y1 is None else y1, camera=dict(pitch=pitch, yaw=yaw),
                algorithm=alg, scenePresetIndex=preset, accelerationStructure=accel, overshootFactor=over, stepSize=step,
                synthetic=synthetic)
