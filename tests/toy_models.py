"""Small stock-PyTorch stand-ins shared by the CPU tests and tests/golden/make_loop_golden.py: a differentiable toy
'forward operator' and a tiny diffusion model with the four members the RED regulariser uses."""
from collections import namedtuple

import torch

Pred = namedtuple("Pred", ["pred_noise", "pred_x_start"])


class ToyOperator(torch.nn.Module):
    """(B, 1, nz, nx) normalised velocity -> (B, 2, 7, nx) 'seismograms' (a fixed linear map plus a quadratic term)."""

    def __init__(self, nz):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        self.w = torch.randn(2, 7, nz, generator=g)
        self.device = torch.device("cpu")

    def forward(self, v):
        return torch.einsum("stz,bczx->bstx", self.w, v) + 0.1 * torch.einsum("stz,bczx->bstx", self.w, v * v)


class TinyDiffusion(torch.nn.Module):
    """q_sample / model_predictions / num_timesteps / alphas_cumprod on a 2-layer conv net with a timestep embedding."""

    def __init__(self, image_size=12, steps=40):
        super().__init__()
        gen = torch.Generator().manual_seed(0)
        self.image_size, self.num_timesteps = image_size, steps
        self.device = torch.device("cpu")
        self.c1 = torch.nn.Conv2d(1, 6, 3, padding=1)
        self.c2 = torch.nn.Conv2d(6, 1, 3, padding=1)
        self.emb = torch.nn.Embedding(steps, 6)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(0.3 * torch.randn(p.shape, generator=gen))
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


def loop_case(width):
    """Inputs of the inversion-loop fixtures: B = 2 models of 10 x width cells (padded leaf 12 x width+2)."""
    g = torch.Generator().manual_seed(11)
    nz = 10
    mu_true = 1500 + 3000 * torch.rand(2, 1, nz, width, generator=g)
    op = ToyOperator(nz)
    y = op((mu_true - 1500) / 3000 * 2 - 1)
    mu0 = torch.nn.functional.pad(0.2 * torch.randn(2, 1, nz, width, generator=g), (1, 1, 1, 1))
    return op, mu0, mu_true, y
