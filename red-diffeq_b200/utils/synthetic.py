"""Deterministic synthetic velocity models of the OpenFWI / Marmousi / Overthrust shapes.

numpy only (PCG64 streams are reproducible across machines), used by tests, golden fixtures and bench.py.
"""
import numpy as np

SEED = 8888  # the reference's default random_seed (configs/default.yaml)

PDE_OPENFWI = dict(n_grid=70, nt=1000, dx=10.0, dt=0.001, nbc=120, f=15.0, sz=10, gz=10, ng=70, ns=5)
PDE_MARMOUSI = dict(n_grid=190, nt=1000, dx=10.0, dt=0.001, nbc=120, f=15.0, sz=10, gz=10, ng=190, ns=5)
PDE_OVERTHRUST = dict(PDE_MARMOUSI)


def velocity_models(B, nz, nx, seed=SEED, noise=1e-4):
    """(B, 1, nz, nx) float32, normalised to [-1, 1]: depth trend + lateral sinusoid + fault step + noise."""
    rng = np.random.default_rng(seed)
    z = np.linspace(0.0, 1.0, nz)[None, :, None]
    x = np.linspace(0.0, 1.0, nx)[None, None, :]
    amp = rng.uniform(0.03, 0.12, size=(B, 1, 1))
    freq = rng.uniform(3.0, 9.0, size=(B, 1, 1))
    phase = rng.uniform(0.0, 2 * np.pi, size=(B, 1, 1))
    fault_x = rng.uniform(0.3, 0.7, size=(B, 1, 1))
    fault_h = rng.uniform(-0.12, 0.12, size=(B, 1, 1))
    depth = z + amp * np.sin(freq * x + phase) + fault_h * (x > fault_x)
    v = np.clip(1500.0 + 3000.0 * depth, 1500.0, 4500.0)
    vn = (v - 1500.0) / 3000.0 * 2.0 - 1.0
    vn = vn + noise * rng.standard_normal(vn.shape)
    return np.clip(vn, -1.0, 1.0).astype(np.float32)[:, None]


def cotangent(shape, seed=SEED + 1):
    """Fixed standard-normal cotangent for VJP comparisons (float32)."""
    return np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
