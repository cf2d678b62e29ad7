"""GPU: the operator inside an inversion loop shaped like the reference's InversionEngine.optimize
(core/inversion.py:42-92: Adam on a padded, normalised leaf `mu`, forward on the slice mu[:, :, 1:-1, 1:-1], masked L1
data misfit per model, clamp to [-1, 1], cosine LR) -- without the diffusion regulariser, which stays on the stock
PyTorch path and is out of scope here.  Checks drop-in behaviour, not convergence quality."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_adam_loop_reduces_the_data_misfit():
    from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize
    from red_diffeq_b200.utils import synthetic
    ctx = dict(n_grid=24, nt=160, dx=10.0, dt=0.001, nbc=12, f=25.0, sz=10, gz=10, ng=24, ns=3)
    dev = torch.device("cuda:0")
    op = FWIForward(ctx, dev, normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none).to(dev)
    assert callable(op)
    B, nz, nx = 2, 20, 24
    mu_true = torch.tensor(synthetic.velocity_models(B, nz, nx, seed=21), device=dev)
    with torch.no_grad():
        y = op(mu_true)                                        # observed data (no history kept under no_grad)
    assert not y.requires_grad
    mu0 = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(mu_true, (3, 3, 3, 3), mode="replicate"), 7, stride=1)  # smoothed start
    mu = torch.nn.functional.pad(mu0, (1, 1, 1, 1), value=0.0).clone().requires_grad_(True)   # (B,1,nz+2,nx+2) like scripts/run_inversion.py:156
    opt = torch.optim.Adam([mu], lr=0.03)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=12, eta_min=0.0)
    mask = torch.ones_like(y)
    losses = []
    for _ in range(12):
        x0 = mu + 1e-4 * torch.randn_like(mu)
        pred = op(x0[:, :, 1:-1, 1:-1])
        loss = ((y - pred).abs() * mask).sum(dim=(1, 2, 3)) / mask.sum(dim=(1, 2, 3)).clamp(min=1.0)
        opt.zero_grad(set_to_none=True)
        loss.sum().backward()
        opt.step()
        with torch.no_grad():
            mu.data.clamp_(-1, 1)
        sched.step()
        losses.append(float(loss.detach().sum()))
    assert np.isfinite(losses).all()
    assert losses[-1] < 0.7 * losses[0], losses
    assert float(mu.grad[:, :, 0, :].abs().max()) == 0.0        # the padding ring gets no gradient from the data term


def _setup(B=2, nz=20, nx=24):
    from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize
    from red_diffeq_b200.utils import synthetic
    ctx = dict(n_grid=nx, nt=160, dx=10.0, dt=0.001, nbc=12, f=25.0, sz=10, gz=10, ng=nx, ns=3)
    dev = torch.device("cuda:0")
    op = FWIForward(ctx, dev, normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
    mu_true_n = torch.tensor(synthetic.velocity_models(B, nz, nx, seed=21), device=dev)
    with torch.no_grad():
        y = op(mu_true_n)
    mu_true = v_denormalize(mu_true_n)                      # the reference hands optimize() the model in m/s
    mu0 = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(mu_true_n, (3, 3, 3, 3), mode="replicate"), 7, stride=1)
    mu0 = torch.nn.functional.pad(mu0, (1, 1, 1, 1), value=0.0)
    return op, mu0, mu_true, y


@pytest.mark.parametrize("reg", [None, "tv"])
def test_inversion_engine_variants_agree(reg):
    """InversionEngine (core/inversion.py): operator + torch loss ops, fused misfit, and the CUDA-graphed iteration run the
    same optimisation -- same losses per iteration, same final model (Adam amplifies last-bit differences a little)."""
    from red_diffeq_b200 import InversionEngine
    op, mu0, mu_true, y = _setup()
    ts = 12
    runs = {}
    for name, kw in {"plain": dict(fused_misfit=False, cuda_graph=False), "fused": dict(fused_misfit=True, cuda_graph=False),
                     "graph": dict(fused_misfit=True, cuda_graph=True)}.items():
        eng = InversionEngine(regularization=reg, **kw)
        mu, res = eng.optimize(mu0, mu_true, y, op, ts=ts, lr=0.03, reg_lambda=0.01, regularization=reg)
        assert eng.used_cuda_graph == (name == "graph")
        assert len(res) == 2 and all(len(res[0][k]) == ts for k in ("total_losses", "obs_losses", "reg_losses", "mae", "rmse", "ssim"))
        runs[name] = (mu.detach().cpu().numpy(), np.array([res[i]["obs_losses"] for i in range(2)]),
                      np.array([res[i]["mae"] for i in range(2)]))
    base = runs["plain"]
    assert np.isfinite(base[1]).all() and (base[1][:, -1] < 0.7 * base[1][:, 0]).all()
    for name in ("fused", "graph"):
        np.testing.assert_allclose(runs[name][1], base[1], rtol=2e-3)
        np.testing.assert_allclose(runs[name][0], base[0], atol=5e-3)
        np.testing.assert_allclose(runs[name][2], base[2], rtol=2e-3)


def test_inversion_engine_missing_traces_and_noise():
    from red_diffeq_b200 import InversionEngine
    op, mu0, mu_true, y = _setup()
    torch.manual_seed(8888)
    eng = InversionEngine(regularization="l2")            # cuda_graph automatic: on for the built-in regularisers
    mu, res = eng.optimize(mu0, mu_true, y, op, ts=8, lr=0.03, reg_lambda=0.01, noise_std=1e-4, missing_number=5, regularization="l2")
    assert eng.used_cuda_graph
    assert mu.shape == (2, 1, 20, 24) and float(mu.detach().abs().max()) <= 1.0
    obs = np.array(res[0]["obs_losses"])
    assert np.isfinite(obs).all() and obs[-1] < obs[0]


def test_files_in_files_out(tmp_path):
    """tools/invert_family.py: .npy family (the reference's input format) -> InversionEngine -> per-model .npz with the
    reference's keys; a short OpenFWI-shaped record so that it runs in seconds."""
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "invert_family.py"), "--synthetic", "3", "--batch", "2",
                          "--ts", "6", "--nt", "300", "--out", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    info = json.loads(out.stdout.strip().splitlines()[-1])
    assert info["models"] == 3 and info["cuda_graph"]
    z = np.load(os.path.join(str(tmp_path), "results", "2_results.npz"))
    assert z["result"].shape == (70, 70) and z["ground_truth"].shape == (70, 70) and z["obs_losses"].shape == (6,)
    assert np.isfinite(z["obs_losses"]).all() and z["obs_losses"][-1] < z["obs_losses"][0]
