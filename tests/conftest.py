import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def detect_kwargs(d):
    kw = {k[3:]: d[k] for k in d.files if k.startswith("kw_")}
    return {k: (v if v.ndim else v.item()) for k, v in kw.items()}


@pytest.fixture(scope="session")
def fits5():
    return golden("fits5_seed0.npz")


@pytest.fixture(scope="session")
def frame0():
    from fluorosequencingimageanalysis_b200 import synth
    import hashlib
    img = synth.synth_frame(0)
    sha = hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest()
    assert sha == str(golden("fits5_seed0.npz")["img_sha"]), "synthetic generator drifted from the golden frame"
    return img


def relerr(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
