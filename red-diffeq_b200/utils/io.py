"""Data formats either side of the hot path: the reference's inputs are ``.npy`` families read through
``np.load(mmap_mode="r")`` -- seismic data (N, ns, nt, n_rec), velocity models (N, 1, nz, nx) in m/s
(scripts/run_inversion.py:282-283) -- and its outputs one ``<index>_results.npz`` per model
(scripts/run_inversion.py:185-216).  Same arrays, same keys, so files written by either side are interchangeable.

Host -> device staging differs from the reference (``torch.from_numpy(mmap[a:b].copy()).float().to(device)``): a batch is
copied from the memory map straight into a reusable PINNED buffer and sent with one asynchronous copy.
"""
import os

import numpy as np
import torch

from .data_trans import v_normalize

RESULT_KEYS = ("result", "initial_velocity", "ground_truth", "total_losses", "obs_losses", "reg_losses", "ssim", "mae", "rmse")


class Family:
    """One seismic / velocity file pair of a dataset family, memory-mapped."""

    def __init__(self, seismic_path, velocity_path):
        self.seismic = np.load(seismic_path, mmap_mode="r")
        self.velocity = np.load(velocity_path, mmap_mode="r")
        if self.seismic.ndim != 4 or self.velocity.ndim != 4 or self.velocity.shape[1] != 1:
            raise ValueError(f"expected seismic (N, ns, nt, n_rec) and velocity (N, 1, nz, nx); got {self.seismic.shape}, {self.velocity.shape}")
        if self.seismic.shape[0] != self.velocity.shape[0]:
            raise ValueError(f"{self.seismic.shape[0]} seismic records for {self.velocity.shape[0]} velocity models")
        self._pinned = None

    def __len__(self):
        return int(self.seismic.shape[0])

    def check_against(self, ctx):
        """The record must have the shape the survey produces (ns shots, nt levels, one trace per receiver)."""
        n_rec = len(ctx["gx"]) if "gx" in ctx and not np.isscalar(ctx["gx"]) else int(ctx["ng"])
        want = (int(ctx["ns"]) if "sx" not in ctx else len(ctx["sx"]), int(ctx["nt"]), n_rec)
        if tuple(self.seismic.shape[1:]) != want:
            raise ValueError(f"seismic records are {tuple(self.seismic.shape[1:])}, the survey produces {want}")

    def batches(self, batch_size, sample_index=None):
        """(start, end) index pairs like scripts/run_inversion.py:285-300: one sample, or consecutive batches."""
        n = len(self)
        if sample_index is not None:
            if sample_index < 0 or sample_index >= n:
                raise IndexError(f"sample_index {sample_index} is out of range [0, {n - 1}]")
            return [(sample_index, sample_index + 1)]
        return [(a, min(a + batch_size, n)) for a in range(0, n, batch_size)]

    def load_batch(self, start, end, device):
        """seismic batch on `device` (fp32), velocity batch on the host (fp32, m/s) -- scripts/run_inversion.py:143-144."""
        shape = (end - start,) + tuple(self.seismic.shape[1:])
        dev = torch.device(device)
        vel = torch.from_numpy(np.array(self.velocity[start:end], dtype=np.float32))
        if dev.type != "cuda":
            return torch.from_numpy(np.array(self.seismic[start:end], dtype=np.float32)).to(dev), vel
        if self._pinned is None or self._pinned.shape[0] < shape[0] or tuple(self._pinned.shape[1:]) != shape[1:]:
            self._pinned = torch.empty(shape, dtype=torch.float32).pin_memory()
        stage = self._pinned[:shape[0]]
        np.copyto(stage.numpy(), self.seismic[start:end], casting="same_kind")      # mmap -> pinned, one pass
        seis = stage.to(dev, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()                               # the pinned buffer is reused
        return seis, vel


def prepare_initial_model(v_true, initial_type=None, sigma=None, linear_coeff=1.0):
    """Initial model in NORMALISED units from the true model in m/s (utils/data_trans.py:66-102): gaussian-smoothed
    (scipy, like the reference, for identical numbers), homogeneous (minimum of the top row) or linear in depth."""
    from scipy.ndimage import gaussian_filter
    assert initial_type in ["smoothed", "homogeneous", "linear"], "please choose from 'smoothed', 'homogeneous', and 'linear'"
    v_np = v_normalize(v_true.detach().cpu().numpy())
    if initial_type == "smoothed":
        out = gaussian_filter(v_np, sigma=sigma)
    elif initial_type == "homogeneous":
        out = np.full_like(v_np, np.min(v_np[0, 0, 0, :]))
    else:
        height = v_np.shape[2]
        ramp = np.linspace(np.min(v_np), np.max(v_np), height).reshape(-1, 1)
        out = np.tile(ramp, (1, v_np.shape[3])).reshape(1, 1, height, -1)
    return torch.tensor(out, dtype=torch.float32, device=v_true.device)


def initial_batch(vel_batch, initial_type, sigma):
    """Per-model initial models, zero-padded by one cell: 70 x 70 -> 72 x 72 (scripts/run_inversion.py:146-159)."""
    models = [torch.nn.functional.pad(prepare_initial_model(vel_batch[i:i + 1], initial_type, sigma=sigma), (1, 1, 1, 1), "constant", 0)
              for i in range(vel_batch.shape[0])]
    return torch.cat(models, dim=0)


def save_batch_results(batch_start, batch_end, mu_batch, results_per_model, initial_model_batch, vel_batch, output_dir):
    """One ``<model index>_results.npz`` per model with the reference's keys (scripts/run_inversion.py:185-216)."""
    mu = mu_batch.detach().cpu().numpy()
    vel = vel_batch.cpu().numpy()
    init = initial_model_batch[:, :, 1:-1, 1:-1].detach().cpu().numpy()
    os.makedirs(output_dir, exist_ok=True)
    paths = []
    for i, model_idx in enumerate(range(batch_start, batch_end)):
        m = results_per_model[i]
        data = {"result": mu[i, 0], "initial_velocity": init[i, 0], "ground_truth": vel[i, 0]}
        for k in RESULT_KEYS[3:]:
            data[k] = np.array(m[k])
        path = os.path.join(os.path.abspath(output_dir), f"{model_idx}_results.npz")
        np.savez(path, **data)
        paths.append(path)
    return paths
