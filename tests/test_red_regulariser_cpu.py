"""CPU: the RED-DiffEq regulariser's call pattern (red-diffeq_b200/regularization/diffusion.py: denoiser under no_grad, patches
batched into one call) against the reference's pattern restated below (regularization/diffusion.py:50-140: autograd graph
through the denoiser then .detach(), one denoiser call per patch) on a small stock-PyTorch stand-in for the diffusion model."""
from collections import namedtuple

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.nn.functional as F  # noqa: E402

Pred = namedtuple("Pred", ["pred_noise", "pred_x_start"])


class TinyDiffusion(torch.nn.Module):
    """The four members the regulariser uses, on a 2-layer conv net with a timestep embedding (image_size 12)."""

    def __init__(self, image_size=12, steps=40):
        super().__init__()
        torch.manual_seed(0)
        self.image_size, self.num_timesteps = image_size, steps
        self.c1 = torch.nn.Conv2d(1, 6, 3, padding=1)
        self.c2 = torch.nn.Conv2d(6, 1, 3, padding=1)
        self.emb = torch.nn.Embedding(steps, 6)
        betas = torch.linspace(1e-3, 0.05, steps)
        self.register_buffer("alphas_cumprod", torch.cumprod(1 - betas, dim=0))
        self.calls = []

    def q_sample(self, x_start, t, noise):
        a = self.alphas_cumprod.gather(-1, t).reshape(-1, 1, 1, 1)
        return a.sqrt() * x_start + (1 - a).sqrt() * noise

    def model_predictions(self, x, t, x_self_cond=None, clip_x_start=False, rederive_pred_noise=False):
        self.calls.append(tuple(x.shape))
        h = torch.tanh(self.c1(x) + self.emb(t)[:, :, None, None])
        return Pred(self.c2(h), None)


def _reference_pattern(dm, mu, seed, use_time_weight=False):
    """regularization/diffusion.py:50-80 (fits) / :82-140 (patched), restated: graph through the denoiser, per-patch calls."""
    from red_diffeq_b200.regularization import calculate_patches
    gen = torch.Generator().manual_seed(seed)
    patched = mu.shape[3] > dm.image_size or mu.shape[2] > dm.image_size
    x0 = mu[:, :, 1:-1, 1:-1] if patched else mu
    b = x0.shape[0]
    t = torch.randint(0, dm.num_timesteps, (b,), generator=gen, dtype=torch.long)
    noise = torch.randn(x0.shape, generator=gen, dtype=x0.dtype)
    if not patched:
        pred = dm.model_predictions(dm.q_sample(x0, t=t, noise=noise), t=t, x_self_cond=None, clip_x_start=True, rederive_pred_noise=True)
        field = (pred.pred_noise - noise).detach()
    else:
        spans, overlaps = calculate_patches(x0.shape[3], x0.shape[2])
        field, wmap = torch.zeros_like(x0), torch.zeros_like(x0)
        for i, (a, e) in enumerate(spans):
            xp, npad = F.pad(x0[:, :, :, a:e], (1, 1, 1, 1)), F.pad(noise[:, :, :, a:e], (1, 1, 1, 1))
            pred = dm.model_predictions(dm.q_sample(xp, t=t, noise=npad), t=t, x_self_cond=None, clip_x_start=True, rederive_pred_noise=True)
            g = (pred.pred_noise[:, :, 1:-1, 1:-1] - npad[:, :, 1:-1, 1:-1]).detach()
            w = torch.ones(e - a)
            if i > 0:
                w[:overlaps[i - 1]] = 0.5
            if i < len(spans) - 1:
                w[-overlaps[i]:] = 0.5
            field[:, :, :, a:e] += g * w.view(1, 1, 1, -1)
            wmap[:, :, :, a:e] += w.view(1, 1, 1, -1)
        field = field / wmap.clamp(min=1e-8)
    reg = field * x0
    if use_time_weight:
        gam = dm.alphas_cumprod.gather(-1, t).reshape(-1, 1, 1, 1)
        reg = reg * torch.sqrt((1 - gam) / gam)
    return reg.view(b, -1).mean(dim=1), t


@pytest.mark.parametrize("width,time_weight", [(10, False), (10, True), (27, False), (27, True), (20, False)])
def test_same_loss_and_gradient_as_the_reference_pattern(width, time_weight):
    from red_diffeq_b200 import REDDiffEq
    dm = TinyDiffusion()
    torch.manual_seed(1)
    mu = (0.5 * torch.randn(3, 1, 12, width + 2)).requires_grad_(True)     # padded leaf: interior 10 x width
    ref_loss, ref_t = _reference_pattern(dm, mu, seed=7, use_time_weight=time_weight)
    ref_loss.sum().backward()
    ref_grad, mu.grad = mu.grad.clone(), None
    n_ref_calls, dm.calls = len(dm.calls), []

    red = REDDiffEq(dm, use_time_weight=time_weight)
    loss, t = red(mu, generator=torch.Generator().manual_seed(7))
    loss.sum().backward()
    assert torch.equal(t, ref_t)                                          # same random stream
    assert torch.allclose(loss, ref_loss, rtol=1e-5, atol=1e-8)
    assert torch.allclose(mu.grad, ref_grad, rtol=1e-5, atol=1e-9)
    assert len(dm.calls) == 1                                             # one denoiser call, whatever the width
    if width > 10:
        k = n_ref_calls
        assert k == int(np.ceil(width / 10)) and dm.calls[0] == (3 * k, 1, 12, 12)
    loop = REDDiffEq(dm, use_time_weight=time_weight, batch_patches=False)
    mu.grad = None
    loss2, _ = loop(mu, generator=torch.Generator().manual_seed(7))
    assert torch.allclose(loss2, loss, rtol=1e-5, atol=1e-8)


def test_engine_builds_the_regulariser_from_a_diffusion_model():
    from red_diffeq_b200 import InversionEngine

    class Op(torch.nn.Module):
        device = torch.device("cpu")

        def forward(self, v):
            return v.sum(dim=2, keepdim=True).expand(-1, 2, 5, -1) * 1.0

    dm = TinyDiffusion()
    dm.device = torch.device("cpu")
    eng = InversionEngine(dm, None, regularization="diffusion", sigma_x0=1e-4)
    mu0 = torch.zeros(2, 1, 12, 12)
    y = torch.ones(2, 2, 5, 10)
    mu, res = eng.optimize(mu0, 1500 + 3000 * torch.rand(2, 1, 10, 10), y, Op(), ts=3, regularization="diffusion")
    assert len(dm.calls) == 3 and np.isfinite(res[0]["reg_losses"]).all() and not eng.used_cuda_graph
