import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["tiny_default", "tiny_custom", "tiny_half_receivers", "openfwi", "marmousi"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """One fixture written by tests/golden/make_golden.py (outputs of the reference's own pde.py)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.name = name
        self.ctx = json.loads(str(z["ctx"]))
        for k in ("n_grid", "nt", "nbc", "ng", "ns"):
            self.ctx[k] = int(self.ctx[k])
        self.v = z["v"]
        self.normalize = bool(z["normalize"])
        self.sample_temporal = int(z["sample_temporal"])
        self.sample_spatial = float(z["sample_spatial"])
        self.cot_seed = int(z["cot_seed"])
        self.cot_checksum = float(z["cot_checksum"])
        self.seis_stride = int(z["seis_stride"])
        self.seis_f32 = z["seis_f32"]
        self.seis_sum = z["seis_sum"]
        self.seis_sumsq = z["seis_sumsq"]
        self.grad_f32 = z["grad_f32"]
        self.grad_f64 = z["grad_f64"]

    def fresh_ctx(self):
        return json.loads(json.dumps(self.ctx))

    def v_phys(self, dtype=np.float32):
        v = self.v.astype(dtype)
        if self.normalize:  # v_denormalize, reference utils/data_trans.py:13-15, same op order
            v = (v + dtype(1)) / dtype(2) * dtype(3000) + dtype(1500)
        return v

    def cotangent(self, shape):
        from red_diffeq_b200.utils import synthetic
        cot = synthetic.cotangent(shape, seed=self.cot_seed)
        assert abs(float(cot.astype(np.float64).sum()) - self.cot_checksum) < 1e-6 * max(1.0, abs(self.cot_checksum)), \
            "cotangent stream differs from the one the fixture was generated with"
        return cot


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return Golden(request.param)


@pytest.fixture(scope="session")
def oracle():
    from oracle import build_oracle, fwi_oracle
    build_oracle.build()
    return fwi_oracle


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))
