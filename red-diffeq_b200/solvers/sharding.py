"""Shot / model partitioner: shards the independent PDE solves of one gradient evaluation over the GPUs
of a box, with a single all-reduce of the velocity gradient (NCCL over NVLink on GPUs).

The reference has no multi-GPU path: every (model b, shot s) pair is an independent solve
(solvers/pde.py:75-81, fields are (B, ns, nz, nx) with no coupling across b or s), the only coupling is
the sum over shots inside the gradient of a model.  So
  * with at least as many models as ranks, whole models are dealt out (no shot of a model leaves its GPU);
  * otherwise the shots are dealt out and every rank models all velocity models for its shots;
in both cases every rank holds the full (replicated) velocity batch `v`, produces the seismograms of its
own work items only, evaluates its part of the data misfit on them (with the *global* normaliser), and
the backward pass all-reduces d loss / d v once.
"""
import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn


def split_range(n, parts, index):
    """[lo, hi) of the index-th of `parts` near-equal contiguous pieces of range(n) (first pieces get the extra)."""
    if parts < 1 or not (0 <= index < parts):
        raise ValueError("bad partition request")
    base, extra = divmod(n, parts)
    lo = index * base + min(index, extra)
    return lo, lo + base + (1 if index < extra else 0)


def plan_partition(n_models, n_shots, world_size, rank, mode="auto"):
    """Work items of `rank`: (mode, model slice, shot indices).

    mode "models": contiguous models, all shots.  mode "shots": all models, contiguous shots.
    "auto" picks "models" when every rank can get at least one model.
    """
    if mode == "auto":
        mode = "models" if n_models >= world_size else "shots"
    if mode == "models":
        lo, hi = split_range(n_models, world_size, rank)
        return mode, slice(lo, hi), np.arange(n_shots)
    if mode == "shots":
        lo, hi = split_range(n_shots, world_size, rank)
        return mode, slice(0, n_models), np.arange(lo, hi)
    raise ValueError(f"unknown partition mode {mode!r}")


class _SumGradAcrossRanks(torch.autograd.Function):
    """Identity in the forward pass; the backward pass all-reduces (sum) the incoming gradient."""

    @staticmethod
    def forward(ctx, v, group):
        ctx.group = group
        return v.view_as(v)

    @staticmethod
    def backward(ctx, grad):
        grad = grad.contiguous()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(ctx.group) > 1:
            dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=ctx.group)
        return grad, None


class _AllReduceSum(torch.autograd.Function):
    """Sum over the ranks in the forward pass; identity in the backward pass (every rank holds the same replicated loss and
    back-propagates it into its own work items only -- the per-rank velocity gradients meet in _SumGradAcrossRanks)."""

    @staticmethod
    def forward(ctx, t, group):
        t = t.contiguous().clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t

    @staticmethod
    def backward(ctx, grad):
        return grad, None


class ShardedFWIForward(nn.Module):
    """Multi-GPU front end of FWIForward (one process per GPU, torch.distributed already initialised).

        op = ShardedFWIForward(ctx, device, normalize=True, v_denorm_func=..., s_norm_func=...)
        seis_local = op(v)                       # (B_local, ns_local, nt_out, n_rec): this rank's work items
        y_local = op.local_slice(y)              # the matching part of the observed data
        loss = misfit(seis_local, y_local) / global_count
        loss.backward()                          # v.grad is the full gradient on every rank (one all-reduce)

    `operator_factory(ctx, device, shot_subset=...)` builds the per-rank operator; it defaults to the CUDA
    FWIForward and exists so that the host logic can be exercised with a stand-in on CPU (gloo) in tests.
    """

    def __init__(self, ctx, device, group=None, mode="auto", operator_factory=None, **operator_kwargs):
        super().__init__()
        self.group = group
        self.mode = mode
        self.device = device
        self._ctx = ctx
        self._factory = operator_factory
        self._kwargs = operator_kwargs
        self._ops = {}
        self.world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world_size > 1 else 0
        self.last_partition = None

    def _n_shots(self):
        return len(self._ctx["sx"]) if "sx" in self._ctx else int(self._ctx["ns"])

    def partition(self, n_models):
        return plan_partition(n_models, self._n_shots(), self.world_size, self.rank, self.mode)

    def _operator(self, shots):
        key = tuple(int(s) for s in shots)
        op = self._ops.get(key)
        if op is None:
            if self._factory is not None:
                op = self._factory(dict(self._ctx), self.device, shot_subset=key, **self._kwargs)
            else:
                from .pde import FWIForward
                op = FWIForward(dict(self._ctx), self.device, shot_subset=key, **self._kwargs)
            self._ops[key] = op
        return op

    def forward(self, v):
        mode, models, shots = self.partition(v.shape[0])
        self.last_partition = (mode, models, shots)
        v = _SumGradAcrossRanks.apply(v, self.group)
        if len(shots) == 0 or models.stop <= models.start:
            # nothing to do on this rank: a zero-size result that still lets backward() reach the all-reduce
            return v[:0].sum() * v.new_zeros((0, 0, 0, 0))
        return self._operator(shots)(v[models])

    def misfit(self, v, y, mask=None):
        """Shard-local data misfit: the reference's per-model observation loss
        (core/losses.py:15-40) of the FULL observed data `y` (B, ns, nt, n_rec) [and mask], shape (B,), identical on
        every rank.  Each rank runs its work items through the operator's fused ``misfit_stats`` (seismograms and
        residuals never leave the GPU that modelled them), the (B, 2) partial sums are all-reduced (16 B per model), and
        the normaliser is the global count.  ``loss.sum().backward()`` then leaves the full gradient on every rank after
        the single all-reduce of d loss / d v."""
        B = v.shape[0]
        mode, models, shots = self.partition(B)
        self.last_partition = (mode, models, shots)
        v = _SumGradAcrossRanks.apply(v, self.group)
        stats = torch.zeros((B, 2), dtype=torch.float64, device=v.device)
        if len(shots) == 0 or models.stop <= models.start:
            stats = stats + 0.0 * v.sum().to(torch.float64)  # idle rank: keeps backward() on the way to the all-reduce
        else:
            local = self._operator(shots).misfit_stats(v[models], self.local_slice(y), None if mask is None else self.local_slice(mask))
            stats = torch.cat([stats[:models.start], local, stats[models.stop:]], dim=0)
        stats = _AllReduceSum.apply(stats, self.group)
        return (stats[:, 0] / stats[:, 1].clamp(min=1.0)).to(torch.float32)

    def local_slice(self, y):
        """The part of a (B, ns, nt, n_rec) tensor (observed data, masks) that matches forward()'s output."""
        if self.last_partition is None:
            _, models, shots = self.partition(y.shape[0])
        else:
            _, models, shots = self.last_partition
        return y[models][:, torch.as_tensor(shots, device=y.device, dtype=torch.long)]
