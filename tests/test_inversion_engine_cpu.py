"""CPU: host logic of the inversion-loop driver (red-diffeq_b200/core/inversion.py) with a toy differentiable operator --
the loop must reproduce the reference's InversionEngine.optimize body (core/inversion.py:69-113: L1 misfit per model,
regulariser, Adam, clamp, CosineAnnealingLR, per-iteration MAE / RMSE) step for step."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


class _ToyOperator(torch.nn.Module):
    """(B, 1, nz, nx) normalised velocity -> (B, 2, 7, nx) 'seismograms' (a fixed linear map plus a quadratic term)."""

    def __init__(self, nz):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        self.w = torch.randn(2, 7, nz, generator=g)
        self.device = torch.device("cpu")

    def forward(self, v):
        return torch.einsum("stz,bczx->bstx", self.w, v) + 0.1 * torch.einsum("stz,bczx->bstx", self.w, v * v)


def _reference_loop(op, mu, mu_true, y, mask, ts, lr, lam, reg):
    from red_diffeq_b200.core.inversion import tikhonov_loss, total_variation_loss
    from red_diffeq_b200 import v_normalize
    mu = mu.float().clone().detach().requires_grad_(True)
    opt = torch.optim.Adam([mu], lr=lr)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=ts, eta_min=0.0)
    rec = {"total_losses": [], "obs_losses": [], "mae": [], "rmse": []}
    for _ in range(ts):
        pred = op(mu[:, :, 1:-1, 1:-1])
        loss = torch.nn.L1Loss(reduction="none")(y.float(), pred.float()) * mask
        loss_obs = loss.sum(dim=(1, 2, 3)) / mask.sum(dim=(1, 2, 3)).clamp(min=1.0)
        r = {"tv": total_variation_loss, "l2": tikhonov_loss}[reg](mu) if reg else torch.zeros(mu.shape[0])
        total = loss_obs + lam * r
        opt.zero_grad(set_to_none=True)
        total.sum().backward()
        opt.step()
        with torch.no_grad():
            mu.data.clamp_(-1, 1)
        sched.step()
        err = mu.detach()[:, :, 1:-1, 1:-1] - v_normalize(mu_true)
        rec["total_losses"].append(total.detach().numpy().copy())
        rec["obs_losses"].append(loss_obs.detach().numpy().copy())
        rec["mae"].append(err.abs().mean(dim=(1, 2, 3)).numpy().copy())
        rec["rmse"].append((err ** 2).mean(dim=(1, 2, 3)).sqrt().numpy().copy())
    return mu.detach()[:, :, 1:-1, 1:-1], rec


@pytest.mark.parametrize("reg", [None, "tv", "l2"])
def test_loop_equals_the_reference_loop_body(reg):
    from red_diffeq_b200 import InversionEngine
    B, nz, nx, ts = 3, 6, 8, 15
    g = torch.Generator().manual_seed(0)
    mu_true = 1500 + 3000 * torch.rand(B, 1, nz, nx, generator=g)
    op = _ToyOperator(nz)
    from red_diffeq_b200 import v_normalize
    y = op(v_normalize(mu_true))
    mu0 = torch.nn.functional.pad(0.3 * torch.randn(B, 1, nz, nx, generator=g), (1, 1, 1, 1))
    mask = torch.ones_like(y)
    mask[:, :, :, 2] = 0
    engine = InversionEngine(regularization=reg, fused_misfit=False)
    mu_a, results = engine.optimize(mu0, mu_true, y * mask, op, ts=ts, lr=0.03, reg_lambda=0.01, regularization=reg, mask=mask)
    mu_b, rec = _reference_loop(op, mu0, mu_true, y * mask, mask, ts, 0.03, 0.01, reg)
    assert not engine.used_cuda_graph
    assert torch.equal(mu_a.detach(), mu_b)
    assert len(results) == B and sorted(results[0]) == ["mae", "obs_losses", "reg_losses", "rmse", "ssim", "total_losses"]
    for i in range(B):
        for k in ("total_losses", "obs_losses", "mae", "rmse"):
            assert len(results[i][k]) == ts
            np.testing.assert_allclose(np.array(results[i][k]), np.array([rec[k][t][i] for t in range(ts)]), rtol=1e-6)
        assert np.isnan(results[i]["ssim"]).all()       # no SSIM callable given


def test_argument_checks_and_data_perturbations():
    from red_diffeq_b200 import InversionEngine, add_noise_to_seismic, missing_trace
    op = _ToyOperator(6)
    mu, y = torch.zeros(2, 1, 8, 10), torch.zeros(3, 2, 7, 8)
    with pytest.raises(ValueError, match="Batch size"):
        InversionEngine().optimize(mu, mu[:, :, 1:-1, 1:-1], y, op)
    with pytest.raises(ValueError, match="Unknown regularization"):
        InversionEngine().optimize(mu, mu[:, :, 1:-1, 1:-1], y[:2], op, regularization="lasso")
    with pytest.raises(ValueError, match="callable"):
        InversionEngine().optimize(mu, mu[:, :, 1:-1, 1:-1], y[:2], None)
    with pytest.raises(ValueError, match="stock PyTorch path"):
        InversionEngine().optimize(mu, mu[:, :, 1:-1, 1:-1], y[:2], op, regularization="diffusion")
    g = torch.Generator().manual_seed(1)
    y = torch.randn(2, 3, 5, 9, generator=g)
    ym, mask = missing_trace(y, 4, generator=g)
    assert mask.shape == y.shape and float(mask.sum()) == 2 * 3 * 5 * 5
    assert torch.equal(mask[:, 0], mask[:, 2])                       # the same receivers are missing for every shot
    assert torch.equal(ym, y * mask)
    assert missing_trace(y, 0)[0] is y and add_noise_to_seismic(y, 0.0) is y
    n = add_noise_to_seismic(torch.zeros(200, 1, 50, 10), 0.5, generator=g)
    assert abs(float(n.std()) - 0.5) < 0.01
    lap = add_noise_to_seismic(torch.zeros(200, 1, 50, 10), 0.5, noise_type="laplace", generator=g)
    assert abs(float(lap.abs().mean()) - 0.5) < 0.01                  # E|X| = b for Laplace(0, b)


def test_user_regulariser_with_time_tensor():
    from red_diffeq_b200 import InversionEngine
    op = _ToyOperator(6)
    mu_true = 1500 + 3000 * torch.rand(2, 1, 6, 8)
    y = op(torch.zeros(2, 1, 6, 8))
    calls = []

    def reg(mu):     # shaped like RegularizationMethod.get_reg_loss: (loss per model, time tensor)
        calls.append(mu.shape)
        return (mu ** 2).mean(dim=(1, 2, 3)), torch.zeros(2, dtype=torch.long)

    engine = InversionEngine(regularization="diffusion", regularizer=reg)   # operator without .misfit: torch loss ops
    mu, res = engine.optimize(torch.zeros(2, 1, 8, 10), mu_true, y, op, ts=4, regularization="diffusion")
    assert len(calls) == 4 and calls[0] == (2, 1, 8, 10) and np.isfinite(res[1]["reg_losses"]).all()
