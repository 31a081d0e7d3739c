import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _have_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def port():
    import oracle

    oracle.build(ref=False)
    return oracle.Port()


@pytest.fixture(scope="session")
def ref():
    import oracle

    if not oracle.have_ref():
        pytest.skip("reference .so not available (no /root/reference and no prebuilt oracle/_ref)")
    return oracle.Ref()


@pytest.fixture(scope="session")
def pair_0600():
    z = np.load(GOLDEN / "pair_0600_320x180.npz")
    return z["left"], z["right"]


@pytest.fixture(scope="session")
def golden_0600():
    return np.load(GOLDEN / "ref_0600_320x180_d48.npz")


@pytest.fixture(scope="session")
def golden_synth():
    return np.load(GOLDEN / "ref_synth_96x128_d24.npz")


@pytest.fixture(scope="session")
def native_lib():
    import tea_stereo_matching_b200 as t

    if not t.LIB_PATH.exists():
        t.build_native()
    return t.lib()
