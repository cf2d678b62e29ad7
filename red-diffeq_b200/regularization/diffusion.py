"""Call pattern of the RED-DiffEq regulariser around a stock-PyTorch diffusion model.

The denoiser itself -- the reference's U-Net / GaussianDiffusion (red_diffeq/models/diffusion.py) -- is NOT rebuilt here: it
stays on the stock PyTorch path and is handed in as an object with the four members the reference's regulariser uses
(regularization/diffusion.py:69-72): ``q_sample(x0, t, noise)``, ``model_predictions(x_t, t, x_self_cond, clip_x_start,
rederive_pred_noise).pred_noise``, ``num_timesteps``, ``alphas_cumprod`` (+ optional ``image_size``).  What this module changes
is how it is called once the PDE solve costs milliseconds and the regulariser becomes the critical path:

  * the reference builds the autograd graph of q_sample + U-Net for every iteration and then throws it away with
    ``.detach()`` (regularization/diffusion.py:74-75, :130) -- the regulariser's gradient w.r.t. mu is just the detached field
    (pred_noise - noise) times the explicit ``* mu``.  Here the denoiser runs under ``torch.no_grad()``: no activations kept;
  * for models wider than the denoiser's input (Marmousi / Overthrust: 70 x 190 -> three 70 x 70 patches,
    regularization/diffusion.py:82-140) the reference calls the U-Net once per patch; here the k patches are stacked on the batch
    axis and denoised in ONE call of batch k * B, then blended with the same 0.5 overlap weights.

Same random stream as the reference (one randint for the timesteps, one randn for the noise, in that order, from the same
generator), so a seeded run draws the same timesteps and noise.
"""
import math

import torch
import torch.nn.functional as F


def calculate_patches(width, height):
    """Patch columns [start, end) of width `height` covering `width`, and the overlaps between consecutive patches
    (regularization/diffusion.py:7-27): 190 x 70 -> [(0, 70), (60, 130), (120, 190)], [10, 10]."""
    k = math.ceil(width / height)
    if k == 1:
        return [(0, width)], []
    stride = (width - height) / (k - 1)
    spans = [(int(i * stride), min(int(i * stride) + height, width)) for i in range(k - 1)] + [(width - height, width)]
    return spans, [spans[i][1] - spans[i + 1][0] for i in range(k - 1)]


def _pad1(x):
    return F.pad(x, (1, 1, 1, 1), mode="constant", value=0)   # utils/diffusion_utils.py:9-11


class REDDiffEq:
    """``reg_loss, time_tensor = REDDiffEq(diffusion_model)(mu)`` with mu the padded (B, 1, nz+2, nx+2) leaf of the inversion loop;
    dispatches like RegularizationMethod.get_reg_loss (regularization/base.py:27-33): patched when mu is larger than the
    denoiser's input."""

    def __init__(self, diffusion_model, use_time_weight=False, sigma_x0=0.0001, fixed_timestep=None, batch_patches=True):
        self.diffusion_model = diffusion_model
        self.use_time_weight = use_time_weight
        self.sigma_x0 = sigma_x0
        self.fixed_timestep = fixed_timestep
        self.batch_patches = batch_patches
        size = getattr(diffusion_model, "image_size", 72)
        self.input_size = size[0] if isinstance(size, (tuple, list)) else size

    # -- pieces -----------------------------------------------------------------------------------
    def _draw(self, shape, batch, device, dtype, generator):
        t_max = self.fixed_timestep if self.fixed_timestep is not None else self.diffusion_model.num_timesteps
        t = torch.randint(0, t_max, (batch,), generator=generator, device=device, dtype=torch.long)
        return t, torch.randn(shape, generator=generator, device=device, dtype=dtype)

    @torch.no_grad()
    def _score(self, x0, t, noise):
        """(pred_noise - noise) for x0 of the denoiser's input size; nothing is recorded for autograd."""
        dm = self.diffusion_model
        x_t = dm.q_sample(x0, t=t, noise=noise)
        pred = dm.model_predictions(x_t, t=t, x_self_cond=None, clip_x_start=True, rederive_pred_noise=True)
        return pred.pred_noise - noise

    def _finish(self, field, x0, t):
        reg = field * x0                                        # the only differentiable use of mu
        if self.use_time_weight:                                # regularization/diffusion.py:42-48
            gamma = self.diffusion_model.alphas_cumprod.gather(-1, t).reshape(-1, 1, 1, 1)
            reg = reg * torch.sqrt((1.0 - gamma) / gamma)
        b = x0.shape[0]
        return reg.reshape(b, -1).mean(dim=1), field.reshape(b, -1).mean(dim=1), t

    # -- the two entry points of the reference ----------------------------------------------------------
    def get_reg_loss(self, mu, generator=None):
        t, noise = self._draw(mu.shape, mu.shape[0], mu.device, mu.dtype, generator)
        return self._finish(self._score(mu.detach(), t, noise), mu, t)

    def get_reg_loss_patched(self, mu, generator=None):
        x0 = mu[:, :, 1:-1, 1:-1]
        b, _, height, width = x0.shape
        spans, overlaps = calculate_patches(width, height)
        t, noise = self._draw(x0.shape, b, x0.device, x0.dtype, generator)
        xd = x0.detach()
        if self.batch_patches:   # one denoiser call for all patches: batch index = patch * B + model
            px = torch.cat([_pad1(xd[:, :, :, a:e]) for a, e in spans], dim=0)
            pn = torch.cat([_pad1(noise[:, :, :, a:e]) for a, e in spans], dim=0)
            scores = self._score(px, t.repeat(len(spans)), pn)[:, :, 1:-1, 1:-1].split(b, dim=0)
        else:
            scores = [self._score(_pad1(xd[:, :, :, a:e]), t, _pad1(noise[:, :, :, a:e]))[:, :, 1:-1, 1:-1] for a, e in spans]
        field = torch.zeros_like(xd)
        weight_map = torch.zeros_like(xd)
        for i, ((a, e), s) in enumerate(zip(spans, scores)):   # halves in the overlaps (:121-131)
            w = torch.ones(e - a, device=xd.device, dtype=xd.dtype)
            if i > 0:
                w[:overlaps[i - 1]] = 0.5
            if i < len(spans) - 1:
                w[-overlaps[i]:] = 0.5
            field[:, :, :, a:e] += s * w
            weight_map[:, :, :, a:e] += w
        field = field / weight_map.clamp(min=1e-8)
        return self._finish(field, x0, t)

    def __call__(self, mu, generator=None):
        patched = mu.shape[3] > self.input_size or mu.shape[2] > self.input_size
        reg, _, t = (self.get_reg_loss_patched if patched else self.get_reg_loss)(mu, generator=generator)
        return reg, t
