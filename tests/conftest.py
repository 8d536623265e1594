import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    """True when the CUDA library sees a device.  On a box that HAS an NVIDIA device node a zero count is retried (the first
    CUDA call of a fresh container can race the driver's initialisation) and reported loudly: a GPU suite that silently
    skips itself proves nothing."""
    import time
    has_node = os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl")
    last = ""
    for attempt in range(5 if has_node else 1):
        try:
            from bpvo_b200 import _capi
            lib = _capi.lib()
            if lib.bpvo_b200_device_count() > 0:
                return True
            last = lib.bpvo_b200_last_error().decode()
        except Exception as e:      # noqa: BLE001
            last = repr(e)
        if has_node:
            time.sleep(2.0)
    if has_node:
        sys.stderr.write(f"\n[conftest] NVIDIA device node present but bpvo_b200 sees no CUDA device: {last}\n")
    return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


def make_params(descriptor="intensity", levels=3, loss="tukey", **kw):
    from bpvo_b200.types import AlgorithmParameters, DescriptorType, LossFunctionType, VerbosityType
    d = {"intensity": DescriptorType.kIntensity, "bitplanes": DescriptorType.kBitPlanes,
         "gradient": DescriptorType.kIntensityAndGradient, "dfields": DescriptorType.kDescriptorFieldsFirstOrder}[descriptor]
    l = {"tukey": LossFunctionType.kTukey, "huber": LossFunctionType.kHuber, "l2": LossFunctionType.kL2}[loss]
    return AlgorithmParameters(descriptor=d, numPyramidLevels=levels, lossFunction=l, verbosity=VerbosityType.kSilent, **kw)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
