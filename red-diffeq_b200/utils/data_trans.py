"""The two normalisation callables the reference's drivers inject into FWIForward
(utils/data_trans.py:8-20 there; scripts/run_inversion.py:79-85), and the two observed-data perturbations the
inversion loop applies before it starts (utils/data_trans.py:33-63, :105-153; core/inversion.py:64-67)."""
import torch


def v_normalize(v):
    """m/s in [1500, 4500] -> [-1, 1]."""
    return (v - 1500) / 3000 * 2 - 1


def v_denormalize(v_norm):
    """[-1, 1] -> m/s in [1500, 4500]."""
    return (v_norm + 1) / 2 * 3000 + 1500


def s_normalize_none(s):
    """Seismograms are used unscaled."""
    return s


def _laplace_like(y, scale, generator):
    # inverse-CDF sampling of Laplace(0, scale) from U ~ Uniform(-1/2, 1/2):  x = -scale * sign(U) * ln(1 - 2|U|)
    u = torch.rand(y.shape, generator=generator, device=y.device, dtype=y.dtype).sub_(0.5)
    return torch.log1p(-2.0 * u.abs()).mul_(torch.sign(u)).mul_(-scale)


def add_noise_to_seismic(y, std, noise_type="gaussian", generator=None):
    """Observed data plus Gaussian noise of standard deviation `std`, or Laplace noise of scale `std`, drawn on the data's
    own device (same contract as the reference's utils/data_trans.py:33-63; std == 0 returns the input itself)."""
    if std < 0:
        raise AssertionError("The standard deviation/scale of the noise must be greater than 0")
    samplers = {"gaussian": lambda: torch.randn(y.shape, generator=generator, device=y.device, dtype=y.dtype).mul_(std),
                "laplace": lambda: _laplace_like(y, std, generator)}
    if noise_type not in samplers:
        raise AssertionError(f"Unknown noise type: {noise_type}")
    return y if std == 0 else y + samplers[noise_type]()


def missing_trace(y, num_missing, return_mask=True, generator=None):
    """Drops `num_missing` receivers per model -- the same receivers for every shot and time sample of that model, as a
    dead receiver would be -- and returns the (B, ns, nt, n_rec) observed-data mask (1 = observed) that the masked L1
    misfit takes (contract of the reference's utils/data_trans.py:105-153, including one randperm per model so that a
    seeded generator picks the same receivers).  The mask is built once per (model, receiver) and broadcast."""
    if num_missing < 0:
        raise AssertionError("The number of missing traces must be >= 0")
    if num_missing == 0:
        return (y, torch.ones_like(y)) if return_mask else y
    n_models, n_rec = y.shape[0], y.shape[-1]
    keep = torch.ones((n_models, n_rec), dtype=y.dtype, device=y.device)
    for b in range(n_models):
        keep[b, torch.randperm(n_rec, generator=generator, device=y.device)[:num_missing]] = 0
    mask = keep[:, None, None, :].expand_as(y).contiguous()
    return (y * mask, mask) if return_mask else y * mask
