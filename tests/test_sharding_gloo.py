"""CPU, 2 processes over gloo: the shot/model partitioner + single gradient all-reduce reproduce the
single-process gradient.  The per-rank operator is a stand-in built on the CPU oracle (the CUDA operator
needs a GPU); what is under test is the host logic of ShardedFWIForward."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from conftest import ROOT  # noqa: E402

CTX = dict(n_grid=16, nt=130, dx=10.0, dt=0.001, nbc=8, f=25.0, sz=10, gz=10, ng=16, ns=4)
NZ, NX = 12, 16


class _OracleOp(torch.nn.Module):
    """FWIForward-shaped stand-in: forward + adjoint through the CPU oracle (physical velocities in, no normalisation)."""

    def __init__(self, ctx, device, shot_subset=None, normalize=False):
        super().__init__()
        self.normalize, self.device = normalize, torch.device("cpu")
        from oracle import fwi_oracle
        ctx = dict(ctx)
        if shot_subset is not None:  # sources of this shard, in grid units like a user-supplied ctx['sx']
            all_sx = np.linspace(0, ctx["n_grid"] - 1, num=ctx["ns"])
            ctx["sx"] = list(all_sx[list(shot_subset)])
        self.oracle, self.survey = fwi_oracle, fwi_oracle.Survey(ctx, NZ, NX)

    def forward(self, v):
        oracle, survey = self.oracle, self.survey
        if self.normalize:   # [-1, 1] -> m/s like the product operator with v_denormalize
            v = (v + 1) / 2 * 3000 + 1500

        class Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, v):
                ctx.save_for_backward(v)
                return torch.from_numpy(oracle.forward(survey, v.detach().numpy()))

            @staticmethod
            def backward(ctx, g):
                (v,) = ctx.saved_tensors
                _, grad = oracle.gradient(survey, v.numpy(), g.contiguous().numpy())
                return torch.from_numpy(grad)

        return Fn.apply(v)

    def misfit_stats(self, v, y, mask=None):
        """Same contract as FWIForward.misfit_stats, composed from torch ops (the fused CUDA kernel needs a GPU)."""
        seis = self.forward(v)
        m = torch.ones_like(seis) if mask is None else mask.float()
        dims = (1, 2, 3)
        return torch.stack([((y - seis).abs() * m).sum(dim=dims).double(), m.sum(dim=dims).double()], dim=1)


def _worker(rank, world, port, mode, B, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from red_diffeq_b200.solvers.sharding import ShardedFWIForward
    rng = np.random.default_rng(5)
    v_np = (1500 + 3000 * rng.random((B, 1, NZ, NX))).astype(np.float32)
    y_np = rng.standard_normal((B, CTX["ns"], CTX["nt"], CTX["ng"])).astype(np.float32)
    op = ShardedFWIForward(dict(CTX), "cpu", mode=mode, operator_factory=lambda ctx, dev, shot_subset=None: _OracleOp(ctx, dev, shot_subset))
    v = torch.tensor(v_np, requires_grad=True)
    seis = op(v)
    y = op.local_slice(torch.tensor(y_np))
    loss = ((seis - y) ** 2).sum() / y_np.size        # global normaliser
    loss.backward()
    tot = loss.detach().clone()
    dist.all_reduce(tot)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), grad=v.grad.numpy(), loss=tot.numpy(), shape=np.array(seis.shape))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,B", [("models", 2), ("shots", 1), ("auto", 3)])
def test_sharded_gradient_equals_single_process(tmp_path, mode, B, oracle):
    port = 29500 + (os.getpid() + hash(mode)) % 2000
    mp.spawn(_worker, args=(2, port, mode, B, str(tmp_path)), nprocs=2, join=True)
    rng = np.random.default_rng(5)
    v_np = (1500 + 3000 * rng.random((B, 1, NZ, NX))).astype(np.float32)
    y_np = rng.standard_normal((B, CTX["ns"], CTX["nt"], CTX["ng"])).astype(np.float32)
    survey = oracle.Survey(dict(CTX), NZ, NX)
    seis = oracle.forward(survey, v_np)
    cot = (2.0 * (seis - y_np) / y_np.size).astype(np.float32)
    _, grad = oracle.gradient(survey, v_np, cot)
    loss = ((seis - y_np) ** 2).sum() / y_np.size
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["grad"], r1["grad"])                       # every rank holds the full gradient
    assert np.allclose(r0["loss"], loss, rtol=1e-5)
    rel = np.linalg.norm(r0["grad"] - grad) / np.linalg.norm(grad)
    assert rel < 1e-5, rel
    assert int(r0["shape"][0]) * int(r0["shape"][1]) + int(r1["shape"][0]) * int(r1["shape"][1]) == B * CTX["ns"]


def _misfit_worker(rank, world, port, mode, B, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from red_diffeq_b200.solvers.sharding import ShardedFWIForward
    rng = np.random.default_rng(6)
    v_np = (1500 + 3000 * rng.random((B, 1, NZ, NX))).astype(np.float32)
    y_np = (1e-3 * rng.standard_normal((B, CTX["ns"], CTX["nt"], CTX["ng"]))).astype(np.float32)
    mask_np = (rng.random(y_np.shape) > 0.3).astype(np.float32)
    op = ShardedFWIForward(dict(CTX), "cpu", mode=mode, operator_factory=lambda ctx, dev, shot_subset=None: _OracleOp(ctx, dev, shot_subset))
    v = torch.tensor(v_np, requires_grad=True)
    loss = op.misfit(v, torch.tensor(y_np), torch.tensor(mask_np))
    loss.sum().backward()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), grad=v.grad.numpy(), loss=loss.detach().numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,B", [("models", 3), ("shots", 1), ("shots", 2)])
def test_sharded_misfit_equals_single_process(tmp_path, mode, B, oracle):
    """Shard-local masked L1 misfit with the global normaliser: the (B,) loss and the gradient on every rank equal the
    single-process LossCalculator.observation_loss semantics (reference core/losses.py:27-36) on the full data."""
    port = 31500 + (os.getpid() + hash((mode, B))) % 2000
    mp.spawn(_misfit_worker, args=(2, port, mode, B, str(tmp_path)), nprocs=2, join=True)
    rng = np.random.default_rng(6)
    v_np = (1500 + 3000 * rng.random((B, 1, NZ, NX))).astype(np.float32)
    y_np = (1e-3 * rng.standard_normal((B, CTX["ns"], CTX["nt"], CTX["ng"]))).astype(np.float32)
    mask_np = (rng.random(y_np.shape) > 0.3).astype(np.float32)
    survey = oracle.Survey(dict(CTX), NZ, NX)
    seis = oracle.forward(survey, v_np)
    count = np.maximum(mask_np.sum(axis=(1, 2, 3)), 1.0)
    loss = (np.abs(y_np - seis) * mask_np).sum(axis=(1, 2, 3), dtype=np.float64) / count
    cot = (np.sign(seis - y_np) * mask_np / count[:, None, None, None]).astype(np.float32)
    _, grad = oracle.gradient(survey, v_np, cot)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["grad"], r1["grad"]) and np.array_equal(r0["loss"], r1["loss"])
    assert r0["loss"].shape == (B,) and np.allclose(r0["loss"], loss, rtol=1e-5)
    rel = np.linalg.norm(r0["grad"] - grad) / np.linalg.norm(grad)
    assert rel < 1e-5, rel


def _engine_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mu, res = _run_engine()
    np.savez(os.path.join(out_dir, f"engine{rank}.npz"), mu=mu, obs=res)
    dist.destroy_process_group()


def _run_engine():
    """InversionEngine around a ShardedFWIForward (oracle-backed stand-in operators): shard-local misfit, one gradient
    all-reduce per iteration, Adam on the replicated leaf."""
    from red_diffeq_b200 import InversionEngine
    from red_diffeq_b200.solvers.sharding import ShardedFWIForward
    rng = np.random.default_rng(9)
    B = 2
    mu_true_n = (0.6 * rng.random((B, 1, NZ, NX)) - 0.3).astype(np.float32)
    op = ShardedFWIForward(dict(CTX), "cpu", mode="shots",
                           operator_factory=lambda ctx, dev, shot_subset=None: _OracleOp(ctx, dev, shot_subset, normalize=True))
    full = _OracleOp(dict(CTX), "cpu", None, normalize=True)
    with torch.no_grad():
        y = full(torch.tensor(mu_true_n))
    mu0 = torch.nn.functional.pad(torch.zeros(B, 1, NZ, NX), (1, 1, 1, 1))
    mu_true = (torch.tensor(mu_true_n) + 1) / 2 * 3000 + 1500
    eng = InversionEngine(regularization="tv")
    mu, res = eng.optimize(mu0, mu_true, y, op, ts=4, lr=0.03, reg_lambda=0.01, regularization="tv")
    assert not eng.used_cuda_graph
    return mu.detach().numpy(), np.array([r["obs_losses"] for r in res])


def test_inversion_engine_on_a_sharded_operator(tmp_path, oracle):
    port = 33500 + os.getpid() % 2000
    mp.spawn(_engine_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "engine0.npz"), np.load(tmp_path / "engine1.npz")
    assert np.array_equal(r0["mu"], r1["mu"]) and np.array_equal(r0["obs"], r1["obs"])      # replicas stay in lock-step
    mu_single, obs_single = _run_engine()                                                   # world size 1, same code
    assert np.allclose(r0["obs"], obs_single, rtol=5e-4)          # Adam amplifies the different summation order a little
    assert np.allclose(r0["mu"], mu_single, atol=2e-3)
    assert (obs_single[:, -1] < obs_single[:, 0]).all()


def _stochastic_engine_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(1000 + 17 * rank)                   # every rank's own stream differs, as in a real launch
    from red_diffeq_b200 import InversionEngine
    from red_diffeq_b200.solvers.sharding import ShardedFWIForward
    from toy_models import TinyDiffusion
    rng = np.random.default_rng(9)
    B = 2
    mu_true_n = (0.6 * rng.random((B, 1, NZ, NX)) - 0.3).astype(np.float32)
    op = ShardedFWIForward(dict(CTX), "cpu", mode="shots",
                           operator_factory=lambda ctx, dev, shot_subset=None: _OracleOp(ctx, dev, shot_subset, normalize=True))
    with torch.no_grad():
        y = _OracleOp(dict(CTX), "cpu", None, normalize=True)(torch.tensor(mu_true_n))
    mu0 = torch.nn.functional.pad(torch.zeros(B, 1, NZ, NX), (1, 1, 1, 1))
    mu_true = (torch.tensor(mu_true_n) + 1) / 2 * 3000 + 1500
    eng = InversionEngine(TinyDiffusion(), regularization="diffusion", sigma_x0=1e-2)
    eng.device = torch.device("cpu")
    mu, res = eng.optimize(mu0, mu_true, y, op, ts=3, lr=0.03, reg_lambda=0.05, noise_std=1e-4, missing_number=3,
                           regularization="diffusion")
    out = {"mu": mu.detach().numpy(), "obs": np.array([r["obs_losses"] for r in res]), "reg": np.array([r["reg_losses"] for r in res])}
    # a user regulariser that cannot be handed a generator is refused on a sharded operator
    try:
        InversionEngine(regularizer=lambda m: torch.rand(m.shape[0]), regularization="tv").optimize(
            mu0, mu_true, y, op, ts=1, regularization="tv")
        out["refused"] = np.array(0)
    except ValueError:
        out["refused"] = np.array(1)
    np.savez(os.path.join(out_dir, f"stoch{rank}.npz"), **out)
    dist.destroy_process_group()


def test_sharded_engine_draws_the_same_random_numbers_on_every_rank(tmp_path, oracle):
    """with noise, missing traces, x0 noise and the diffusion regulariser's timesteps every rank must see
    rank 0's draws -- otherwise the all-reduced sums mix inconsistent data and the replicated mu drifts apart."""
    port = 35500 + os.getpid() % 2000
    mp.spawn(_stochastic_engine_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "stoch0.npz"), np.load(tmp_path / "stoch1.npz")
    assert np.array_equal(r0["mu"], r1["mu"])
    assert np.array_equal(r0["obs"], r1["obs"]) and np.array_equal(r0["reg"], r1["reg"])
    assert np.abs(r0["reg"]).max() > 0 and int(r0["refused"]) == 1 and int(r1["refused"]) == 1
