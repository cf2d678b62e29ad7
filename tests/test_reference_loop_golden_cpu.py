"""CPU: the caller-side pieces built around the hot path (inversion-loop driver, RED regulariser call pattern,
observed-data perturbations) against outputs of the REFERENCE's own modules -- InversionEngine.optimize, RED_DiffEq,
missing_trace, add_noise_to_seismic -- recorded in tests/golden/loop_toy.npz by tests/golden/make_loop_golden.py (toy operator
and tiny diffusion model of tests/toy_models.py; the fixture travels, the reference does not)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

torch = pytest.importorskip("torch")
from toy_models import TinyDiffusion, loop_case  # noqa: E402


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(GOLDEN_DIR, "loop_toy.npz"))


@pytest.mark.parametrize("tag,width,reg", [("tv", 10, "tv"), ("l2", 10, "l2"), ("none", 10, None), ("diffusion", 10, "diffusion"),
                                           ("diffusion_patched", 27, "diffusion")])
def test_inversion_engine_reproduces_the_reference_engine(fx, tag, width, reg):
    from red_diffeq_b200 import InversionEngine
    op, mu0, mu_true, y = loop_case(width)
    eng = InversionEngine(TinyDiffusion(), None, regularization=reg, sigma_x0=1e-4)
    torch.manual_seed(8888)     # the reference's loop draws from the global generator; same order of draws here
    mu, res = eng.optimize(mu0, mu_true, y, op, ts=6, lr=0.03, reg_lambda=0.01, regularization=reg)
    exact = reg != "diffusion" or width <= 12     # batched patches change the convolution batch: last-bit differences
    assert np.allclose(mu.detach().numpy(), fx[f"{tag}/mu"], rtol=0, atol=1e-7 if exact else 1e-5)
    for k in ("total_losses", "obs_losses", "reg_losses", "mae", "rmse"):
        got = np.array([[float(v) for v in r[k]] for r in res])
        assert np.allclose(got, fx[f"{tag}/{k}"], rtol=1e-6 if exact else 1e-4, atol=1e-9), k


@pytest.mark.parametrize("width", [10, 27])
def test_red_regulariser_reproduces_the_reference_class(fx, width):
    from red_diffeq_b200 import REDDiffEq
    mu = torch.tensor(fx[f"red{width}/mu"]).requires_grad_(True)
    loss, t = REDDiffEq(TinyDiffusion(), use_time_weight=True)(mu, generator=torch.Generator().manual_seed(7))
    loss.sum().backward()
    assert np.array_equal(t.numpy(), fx[f"red{width}/t"])
    assert np.allclose(loss.detach().numpy(), fx[f"red{width}/loss"], rtol=1e-5, atol=1e-9)
    assert np.allclose(mu.grad.numpy(), fx[f"red{width}/grad"], rtol=1e-5, atol=1e-10)


def test_data_perturbations_reproduce_the_reference_functions(fx):
    from red_diffeq_b200 import add_noise_to_seismic, missing_trace
    y = torch.tensor(fx["pert/y"])
    ym, mask = missing_trace(y, 4, generator=torch.Generator().manual_seed(2))
    assert np.array_equal(mask.numpy(), fx["pert/mask"]) and np.array_equal(ym.numpy(), fx["pert/missing"])
    g = add_noise_to_seismic(y, 0.3, "gaussian", generator=torch.Generator().manual_seed(3))
    assert np.array_equal(g.numpy(), fx["pert/gauss"])
    lap = add_noise_to_seismic(y, 0.3, "laplace", generator=torch.Generator().manual_seed(3))
    assert np.allclose(lap.numpy(), fx["pert/laplace"], rtol=1e-5, atol=1e-7)


def test_initial_models_reproduce_the_reference_function(fx):
    from red_diffeq_b200.utils.io import prepare_initial_model
    v = torch.tensor(fx["init/v"])
    assert np.array_equal(prepare_initial_model(v, "smoothed", sigma=3.0).numpy(), fx["init/smoothed"])
    assert np.array_equal(prepare_initial_model(v, "homogeneous").numpy(), fx["init/homogeneous"])
    assert np.array_equal(prepare_initial_model(v, "linear").numpy(), fx["init/linear"])


def test_patch_layout_reproduces_the_reference_function(fx):
    from red_diffeq_b200.regularization import calculate_patches
    for row in fx["patches/table"]:
        w, h, k = (int(v) for v in row[:3])
        spans, overlaps = calculate_patches(w, h)
        assert len(spans) == k
        flat = [v for s in spans for v in s] + list(overlaps)
        assert flat == [int(v) for v in row[3:3 + len(flat)]], (w, h)
