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


def add_noise_to_seismic(y, std, noise_type="gaussian", generator=None):
    """Gaussian (std) or Laplace (scale) noise on the observed data, on the data's device (utils/data_trans.py:33-63)."""
    assert std >= 0, "The standard deviation/scale of the noise must be greater than 0"
    assert noise_type in ["gaussian", "laplace"], f"Unknown noise type: {noise_type}"
    if std == 0:
        return y
    if noise_type == "gaussian":
        noise = torch.randn(y.shape, generator=generator, device=y.device, dtype=y.dtype) * std
    else:   # inverse-transform sampling: X = -b sign(U) log(1 - 2|U|), U ~ Uniform(-0.5, 0.5)
        u = torch.rand(y.shape, generator=generator, device=y.device, dtype=y.dtype) - 0.5
        noise = -std * torch.sign(u) * torch.log(1 - 2 * torch.abs(u))
    return y + noise


def missing_trace(y, num_missing, return_mask=True, generator=None):
    """Zero `num_missing` random receivers of every model -- the same receivers for all its shots -- and return the
    observed-data mask (1 = observed) the masked L1 misfit uses (utils/data_trans.py:105-153)."""
    assert num_missing >= 0, "The number of missing traces must be >= 0"
    mask = torch.ones_like(y)
    if num_missing == 0:
        return (y, mask) if return_mask else y
    y_missing = y.clone()
    for b in range(y.shape[0]):
        idx = torch.randperm(y.shape[3], generator=generator, device=y.device)[:num_missing]
        y_missing[b, :, :, idx] = 0
        mask[b, :, :, idx] = 0
    return (y_missing, mask) if return_mask else y_missing
