"""Writes tests/golden/loop_toy.npz: outputs of the REFERENCE's own caller-side modules -- InversionEngine.optimize
(red_diffeq/core/inversion.py), RED_DiffEq (regularization/diffusion.py), missing_trace / add_noise_to_seismic
(utils/data_trans.py) -- on the toy operator / tiny diffusion model of tests/toy_models.py, run on CPU in the build container.

The reference package itself cannot be imported (its __init__ pulls in ml_collections, accelerate, ...): the needed modules
are plain torch / numpy / scipy and are loaded through stub parent packages.  Run:  python tests/golden/make_loop_golden.py
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference/red_diffeq"


def load_reference():
    for name in ("red_diffeq", "red_diffeq.utils", "red_diffeq.regularization", "red_diffeq.core"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(os.path.dirname(REF), *name.split("."))]
        sys.modules[name] = m
    mods = {n: importlib.import_module(n) for n in ("red_diffeq.utils.data_trans", "red_diffeq.utils.ssim",
                                                    "red_diffeq.regularization.diffusion", "red_diffeq.regularization.base",
                                                    "red_diffeq.core.inversion")}
    return mods


def run_reference_loop(mods, width, regularization, ts=6):
    from toy_models import TinyDiffusion, loop_case
    op, mu0, mu_true, y = loop_case(width)
    dm = TinyDiffusion()
    eng = mods["red_diffeq.core.inversion"].InversionEngine(dm, mods["red_diffeq.utils.ssim"].SSIM(window_size=11),
                                                              regularization=regularization, sigma_x0=1e-4)
    torch.manual_seed(8888)
    mu, res = eng.optimize(mu0, mu_true, y, op, ts=ts, lr=0.03, reg_lambda=0.01, regularization=regularization)
    out = {"mu": mu.detach().numpy()}
    for k in ("total_losses", "obs_losses", "reg_losses", "ssim", "mae", "rmse"):
        out[k] = np.array([[float(v) for v in r[k]] for r in res], dtype=np.float64)
    return out


def main():
    mods = load_reference()
    dt = mods["red_diffeq.utils.data_trans"]
    data = {}
    for tag, width, reg in (("tv", 10, "tv"), ("l2", 10, "l2"), ("none", 10, None), ("diffusion", 10, "diffusion"),
                            ("diffusion_patched", 27, "diffusion")):
        for k, v in run_reference_loop(mods, width, reg).items():
            data[f"{tag}/{k}"] = v
    # the RED regulariser alone, seeded generator
    from toy_models import TinyDiffusion
    red = mods["red_diffeq.regularization.diffusion"].RED_DiffEq(TinyDiffusion(), use_time_weight=True)
    for width in (10, 27):
        g = torch.Generator().manual_seed(5)
        mu = (0.5 * torch.randn(3, 1, 12, width + 2, generator=g)).requires_grad_(True)
        fn = red.get_reg_loss_patched if width > 12 else red.get_reg_loss
        loss, _, t = fn(mu, generator=torch.Generator().manual_seed(7))
        loss.sum().backward()
        data[f"red{width}/mu"], data[f"red{width}/loss"] = mu.detach().numpy(), loss.detach().numpy()
        data[f"red{width}/grad"], data[f"red{width}/t"] = mu.grad.numpy(), t.numpy()
    # data perturbations
    y = torch.randn(3, 2, 5, 9, generator=torch.Generator().manual_seed(1))
    ym, mask = dt.missing_trace(y, 4, generator=torch.Generator().manual_seed(2))
    data["pert/y"], data["pert/missing"], data["pert/mask"] = y.numpy(), ym.numpy(), mask.numpy()
    data["pert/gauss"] = dt.add_noise_to_seismic(y, 0.3, "gaussian", generator=torch.Generator().manual_seed(3)).numpy()
    data["pert/laplace"] = dt.add_noise_to_seismic(y, 0.3, "laplace", generator=torch.Generator().manual_seed(3)).numpy()
    # observation loss (core/losses.py:15-40) and its gradient w.r.t. the predicted data, masked and unmasked
    losses = importlib.import_module("red_diffeq.core.losses")
    calc = losses.LossCalculator(None)
    g = torch.Generator().manual_seed(6)
    pred = (1e-3 * torch.randn(2, 3, 130, 16, generator=g)).requires_grad_(True)
    target = 1e-3 * torch.randn(2, 3, 130, 16, generator=g)
    target[0, 0, :5] = pred.detach()[0, 0, :5]          # exact ties: sign(0) = 0
    mask = (torch.rand(2, 3, 130, 16, generator=g) > 0.3).float()
    weights = torch.tensor([1.0, 1.7])
    data["loss/pred"], data["loss/target"], data["loss/mask"] = pred.detach().numpy(), target.numpy(), mask.numpy()
    for tag, m in (("masked", mask), ("plain", None)):
        pred.grad = None
        loss = calc.observation_loss(pred, target, mask=m)
        (loss * weights).sum().backward()
        data[f"loss/{tag}_loss"], data[f"loss/{tag}_grad"] = loss.detach().numpy(), pred.grad.numpy().copy()
    # patch layout (regularization/diffusion.py:7-27) over a range of widths / heights
    cp = mods["red_diffeq.regularization.diffusion"].calculate_patches
    rows = []
    for h in (10, 70, 72):
        for w in list(range(h, 4 * h + 3, 7)) + [190, 430]:
            spans, overlaps = cp(w, h)
            rows.append([w, h, len(spans)] + [v for s in spans for v in s] + overlaps + [-1] * (3 * 8 - 1 - 3 * len(spans) + 1))
    width = max(len(r) for r in rows)
    data["patches/table"] = np.array([r + [-1] * (width - len(r)) for r in rows], dtype=np.int64)
    # initial models (utils/data_trans.py:66-102)
    v = 1500 + 3000 * torch.rand(1, 1, 14, 18, generator=torch.Generator().manual_seed(4))
    data["init/v"] = v.numpy()
    data["init/smoothed"] = dt.prepare_initial_model(v, "smoothed", sigma=3.0).numpy()
    data["init/homogeneous"] = dt.prepare_initial_model(v, "homogeneous").numpy()
    data["init/linear"] = dt.prepare_initial_model(v, "linear").numpy()
    np.savez_compressed(os.path.join(HERE, "loop_toy.npz"), **data)
    print("wrote loop_toy.npz with", len(data), "arrays")


if __name__ == "__main__":
    main()
