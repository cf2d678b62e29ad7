"""B200-native drop-in for the reference's ``red_diffeq.solvers.pde.FWIForward`` (solvers/pde.py:6-93).

Same constructor, same call contract, same ctx side effects and error behaviour -- but the nt-step
finite-difference loop runs in the sm_100a kernels behind the C ABI of include/rdfwi.h, and the
gradient comes from an explicit reverse-time adjoint inside a ``torch.autograd.Function`` instead of an
autograd tape of ~48 ATen ops per time level.

    op = FWIForward(ctx, device, normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
    seis = op(v)            # v: (B, 1, nz, nx) -> (B, ns, ceil(nt/sample_temporal), n_rec)
    loss(seis).backward()   # d loss / d v, identical (within fp32 tolerance) to the reference's autograd

There is no CPU path: a non-CUDA device or a missing librdfwi.so raises.
"""
import threading

import numpy as np
import torch
import torch.nn as nn

from .. import _cabi
from . import survey as _survey


class _HistoryLease:
    """Exclusive use of one wavefield-history buffer between a forward call and its backward.

    The buffers are owned by the operator and reused from one inversion iteration to the next: the
    history is by far the largest allocation (115 GiB for 64 OpenFWI models) and must not be carved up by
    the caching allocator between iterations.
    """

    def __init__(self, arena, buffer):
        self.arena, self.buffer = arena, buffer

    def release(self):
        if self.buffer is not None:
            self.arena.append(self.buffer)
            self.buffer = None

    def __del__(self):  # graph dropped without backward
        try:
            self.release()
        except Exception:
            pass


def _solve_forward(ctx, v_phys, op, need_grad):
    """Runs rdfwi_forward for a (B, 1, nz, nx) velocity batch in m/s; keeps what the adjoint pass needs on ctx."""
    if v_phys.dim() != 4 or v_phys.shape[1] != 1:
        raise ValueError(f"expected a (B, 1, nz, nx) velocity batch, got {tuple(v_phys.shape)}")
    if not v_phys.is_cuda:
        raise RuntimeError("rdfwi: the velocity batch must live on a CUDA device (there is no CPU fallback)")
    if v_phys.dtype != torch.float32:
        raise TypeError(f"rdfwi computes in fp32 like the reference; got {v_phys.dtype}")
    v = v_phys.detach().contiguous()
    B, _, nz, nx = v.shape
    plan = op._plan_for(nz, nx, v.device)
    with torch.cuda.device(v.device):
        stream = torch.cuda.current_stream().cuda_stream
        hist, hist_bytes, lease, segment = None, 0, None, 0
        policy = None
        if need_grad:
            policy = op._choose_segment(plan, B, v.device)
            segment = policy[0]
            hist_bytes = plan.history_bytes(B, segment)
            lease = op._lease_history(hist_bytes, v.device)
            hist = lease.buffer
        seis = torch.empty((B, plan.ns, plan.nt_out, plan.nrec), dtype=torch.float32, device=v.device)
        ws_bytes = plan.workspace_bytes(B)
        ws = op._workspace(ws_bytes, v.device)
        plan.forward(v.data_ptr(), B, seis.data_ptr(), ws.data_ptr(), ws_bytes,
                     hist.data_ptr() if hist is not None else None, hist_bytes, segment, stream)
        op.last_launches = plan.last_launch_count()
    if need_grad:
        ctx.op, ctx.plan, ctx.hist, ctx.hist_bytes, ctx.segment = op, plan, lease, hist_bytes, segment
        ctx.policy = policy
        ctx.v = v
    return seis, plan, ws, ws_bytes


def _solve_backward(ctx, grad_seis):
    """Runs rdfwi_backward with the cotangent of the seismograms; returns d loss / d v_phys."""
    v = ctx.v
    plan, lease = ctx.plan, ctx.hist
    if lease is None:
        raise RuntimeError("rdfwi: backward called twice or without saved history")
    hist = lease.buffer
    B = v.shape[0]
    g = grad_seis.contiguous()
    if g.dtype != torch.float32:
        g = g.float()
    with torch.cuda.device(v.device):
        stream = torch.cuda.current_stream().cuda_stream
        grad_v = torch.empty_like(v)
        # the plan is shared by every call of the operator: another forward (different batch size) may have moved its
        # policy-owned options since this graph's forward ran -- put back what that forward decided
        ctx.op._apply_policy(plan, ctx.policy)
        ws_bytes = plan.workspace_bytes(B)
        ws = ctx.op._workspace(ws_bytes, v.device)
        plan.backward(v.data_ptr(), B, g.data_ptr(), grad_v.data_ptr(), ws.data_ptr(), ws_bytes,
                      hist.data_ptr() if hist is not None else None, ctx.hist_bytes, ctx.segment, stream)
        ctx.op.last_launches += plan.last_launch_count()
    ctx.hist = None  # first-order only, like every caller in the reference
    ctx.v = None
    lease.release()  # hand the wavefield history back to the operator's arena
    return grad_v


class _WaveSolve(torch.autograd.Function):
    """seismograms = F(v_phys); backward = discrete adjoint."""

    @staticmethod
    def forward(ctx, v_phys, op):
        seis, _plan, _ws, _n = _solve_forward(ctx, v_phys, op, ctx.needs_input_grad[0])
        return seis

    @staticmethod
    def backward(ctx, grad_seis):
        return _solve_backward(ctx, grad_seis), None


class _WaveMisfit(torch.autograd.Function):
    """stats = (sum |y - F(v)| * mask, sum mask) per model, the seismograms never leaving the library as a torch graph.

    Forward: rdfwi_forward, then rdfwi_misfit_l1 turns the seismogram buffer IN PLACE into the sign field
    mask * sign(F(v) - y).  Backward: cotangent = sign field * d loss / d stats[:, 0] (one broadcast multiply),
    rdfwi_backward.  Replaces the ~10 elementwise kernels of core/losses.py:27-40 and their autograd."""

    @staticmethod
    def forward(ctx, v_phys, y, mask, op, keep_seis):
        need_grad = ctx.needs_input_grad[0]
        seis, plan, ws, ws_bytes = _solve_forward(ctx, v_phys, op, need_grad)
        B = seis.shape[0]
        if tuple(y.shape) != tuple(seis.shape):
            raise ValueError(f"observed data {tuple(y.shape)} does not match the modelled seismograms {tuple(seis.shape)}")
        if mask is not None and tuple(mask.shape) != tuple(seis.shape):
            raise ValueError(f"mask {tuple(mask.shape)} does not match the modelled seismograms {tuple(seis.shape)}")
        y = y.to(device=seis.device, dtype=torch.float32).contiguous()
        m = None if mask is None else mask.to(device=seis.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(seis.device):
            stream = torch.cuda.current_stream().cuda_stream
            stats = torch.empty((B, 2), dtype=torch.float64, device=seis.device)
            sign = torch.empty_like(seis) if keep_seis else seis
            plan.misfit_l1(seis.data_ptr(), y.data_ptr(), m.data_ptr() if m is not None else None, B, stats.data_ptr(),
                           sign.data_ptr() if need_grad else None, ws.data_ptr(), ws_bytes, stream)
            op.last_launches += plan.last_launch_count()
        if need_grad:
            ctx.sign = sign
        ctx.mark_non_differentiable(*([seis] if keep_seis else []))
        return (stats, seis) if keep_seis else (stats,)

    @staticmethod
    def backward(ctx, grad_stats, *_unused):
        g = grad_stats[:, 0].to(torch.float32).view(-1, 1, 1, 1)
        cot = ctx.sign.mul_(g)  # in place: the sign field is this Function's own buffer
        ctx.sign = None
        return _solve_backward(ctx, cot), None, None, None, None


class FWIForward(nn.Module):
    """2-D constant-density acoustic forward modelling operator, differentiable w.r.t. the velocity.

    Parameters are those of the reference class (solvers/pde.py:8):
      ctx              dict with n_grid, nt, dx, dt, nbc, f, sz, gz, ng, ns [, sx, gx in grid units]; mutated in place
                       (sx / gx are stored back in metres), exactly like the reference
      device           CUDA device the operator runs on
      sample_temporal  keep every k-th time level
      sample_spatial   fraction of ng receivers when gx is not given
      normalize        if True, v_denorm_func maps the input to m/s and s_norm_func post-processes the output
    """

    def __init__(self, ctx, device, sample_temporal=1, sample_spatial=1.0, normalize=True, v_denorm_func=None,
                 s_norm_func=None, shot_subset=None):
        super().__init__()
        # shot_subset (extension, used by ShardedFWIForward): indices of the shots this instance models
        self._shot_subset = None if shot_subset is None else np.asarray(shot_subset, dtype=np.int64)
        self.device = device
        self.normalize = normalize
        if normalize:
            self.v_denorm_func = v_denorm_func
            self.s_norm_func = s_norm_func
        self.sample_temporal = sample_temporal
        self.ctx = _survey.complete_ctx(ctx, sample_spatial)
        self._plans = {}
        self._history_arena = {}
        self._ws = {}
        self._segment = None
        self._segment_auto = {}
        self._lock = threading.Lock()
        self.last_launches = 0
        self.options = {}
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError(f"rdfwi: FWIForward needs a CUDA device, got {dev}; there is no CPU fallback")
        _cabi.load()  # fail now, loudly, if the extension is not built

    # -- plan management --------------------------------------------------------------------------
    def set_option(self, key, value):
        """Library tunable (see rdfwi_plan_set in include/rdfwi.h); applies to existing and future plans."""
        self.options[key] = int(value)
        self._segment_auto.clear()   # the history policy was decided under the old options
        for plan in self._plans.values():
            plan.set(key, value)

    def _plan_for(self, nz, nx, device):
        key = (int(nz), int(nx), device.index if device.index is not None else torch.cuda.current_device())
        with self._lock:
            plan = self._plans.get(key)
            if plan is None:
                c = self.ctx
                wavelet = _survey.ricker(c["f"], c["dt"], c["nt"])
                isx, isz, igx, igz = _survey.grid_indices(c["sx"], c["sz"], c["gx"], c["gz"], c["dx"], c["nbc"])
                if self._shot_subset is not None:
                    isx = isx[self._shot_subset]
                nzp, nxp = nz + 2 * c["nbc"], nx + 2 * c["nbc"]
                isx = _survey.wrap_indices(isx, nxp, "source column")
                igx = _survey.wrap_indices(igx, nxp, "receiver column")
                isz = int(_survey.wrap_indices(isz, nzp, "source row"))
                igz = int(_survey.wrap_indices(igz, nzp, "receiver row"))
                with torch.cuda.device(key[2]):
                    plan = _cabi.Plan(nz, nx, c["nbc"], c["nt"], self.sample_temporal, isz, igz, isx, igx, c["dx"],
                                      c["dt"], wavelet)
                for k, v in self.options.items():
                    plan.set(k, v)
                self._plans[key] = plan
        return plan

    def set_history_segment(self, segment):
        """Wavefield-history policy.  0 = keep every level (fastest; the default while it fits),
        K >= 3 = keep a pair of levels every K levels and recompute K levels at a time in the backward pass
        (memory / (K/2), one extra forward), None = automatic."""
        self._segment = segment

    # options the history policy owns: every decision starts from the user's values (or the library defaults) for these
    _POLICY_KEYS = ("adj_mode", "u_chunk_shots", "scratch_mb")

    def _apply_policy(self, plan, policy):
        """Puts the plan into the state a (segment, extras) decision stands for -- as a whole: the policy-owned options
        that are not part of the decision go back to the user's values, so nothing leaks from an earlier decision
        (another batch size on the same plan)."""
        seg, extra = policy
        for k in self._POLICY_KEYS:
            plan.set(k, extra.get(k, self.options.get(k, 0)))
        plan.set("history_segment", seg)

    def _choose_segment(self, plan, B, device):
        """History policy of one forward/backward pair: returns (segment, extras) and leaves the plan in that state.
        Automatic mode walks the tiers until the buffers fit the free HBM: (1) every level (the split adjoint, option
        imaging = 1, also needs a scratch history of the adjoint field); (2) no history at all: the backward pass recomputes the forward field
        chunk by chunk on the cluster engine (segment = nt, one extra forward, two scratch histories of one or two waves
        of shots); (3) every level + fused per-level adjoint (no scratch); (4) history checkpointed in time on the
        per-level engine."""
        if self._segment is not None:
            policy = (self._segment, {})
            self._apply_policy(plan, policy)
            return policy
        key = (id(plan), B)
        policy = self._segment_auto.get(key)     # decided once per (plan, batch): cudaMemGetInfo is slow
        if policy is None:
            free, _total = torch.cuda.mem_get_info(device)
            idle = sum(b.numel() for b in self._history_arena.get(str(device), []))
            held = self._ws.get(str(device))
            idle += held.numel() if held is not None else 0   # the workspace is re-used (or replaced) by the next call
            budget = 0.9 * (free + idle + torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device))

            def fits(cand):
                self._apply_policy(plan, cand)
                return plan.history_bytes(B, cand[0]) + plan.workspace_bytes(B) <= budget

            user = self.options
            self._apply_policy(plan, (0, {}))
            clustered = user.get("engine", 0) != 1 and plan.get("cluster_size_used") > 0
            wave = plan.get("cluster_wave") if clustered else 0
            tiers = [(0, {})]
            if clustered and "adj_mode" not in user:
                # a record so long that the scratch history of the split adjoint holds less than a wave of shots leaves
                # most SMs idle in every chunk: recomputing the forward field chunk by chunk is faster than that
                # (only the split adjoint, option imaging = 1, has that scratch history: by default the imaging sums are
                # formed inside the adjoint sweep and a full history that fits is always the fastest tier)
                per_shot = 4.0 * plan.nt * plan.level_floats()
                if user.get("imaging", 0) == 1 and int(40e9 // per_shot) < min(wave, B * plan.ns):
                    tiers = []
                tiers.append((plan.nt, {}))
                if "u_chunk_shots" not in user:
                    tiers.append((plan.nt, {"u_chunk_shots": wave}))
                tiers += [(0, {}), (0, {"adj_mode": 1})]
            elif "adj_mode" not in user:
                # per-level engine: the split adjoint streams the adjoint field of a chunk of shots through a scratch
                # history -- the larger the chunk, the larger (and fewer) its launches: give it the HBM the forward history
                # leaves free; with little room the scratch history shrinks to a few shots (more, smaller chunks) before the
                # fused adjoint -- which needs no scratch history but moves 32 B per adjoint cell-update -- is considered
                if "scratch_mb" not in user and "u_chunk_shots" not in user:
                    per_shot = 4.0 * plan.nt * plan.level_floats()
                    spare = 0.9 * (budget - plan.history_bytes(B, 0)) - 4.0 * plan.level_floats() * B * (2 * plan.ns + 4)
                    if spare > 40e9:
                        tiers.insert(0, (0, {"scratch_mb": int(spare / 1e6)}))
                    elif spare >= per_shot:
                        tiers.append((0, {"scratch_mb": int(spare / 1e6)}))
                tiers.append((0, {"adj_mode": 1}))
            policy = (max(3, int(np.ceil(np.sqrt(2.0 * plan.nt)))), {})   # minimises pairs + segment levels
            for cand in tiers:
                if fits(cand):
                    policy = cand
                    break
            self._segment_auto[key] = policy
        self._apply_policy(plan, policy)
        return policy

    def _lease_history(self, nbytes, device):
        """A history buffer of at least nbytes on `device`, reused across iterations when idle."""
        key = str(device)
        with self._lock:
            arena = self._history_arena.setdefault(key, [])
            for i, buf in enumerate(arena):
                if buf.numel() >= nbytes:
                    return _HistoryLease(arena, arena.pop(i))
            arena.clear()  # idle buffers that are too small: give them back before growing
        free, _total = torch.cuda.mem_get_info(device)
        reserved_free = torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
        if nbytes > free + reserved_free:
            raise torch.cuda.OutOfMemoryError(
                f"rdfwi: the wavefield history needs {nbytes / 2**30:.1f} GiB but only "
                f"{(free + reserved_free) / 2**30:.1f} GiB is available; reduce the batch")
        return _HistoryLease(arena, torch.empty(nbytes, dtype=torch.uint8, device=device))

    def _workspace(self, nbytes, device):
        """The library's scratch for one call (coefficient planes, imaging planes, the scratch histories of the split
        adjoint: up to tens of GB).  Owned by the operator and reused by every call: nothing in it outlives a call (each
        call recomputes its coefficient planes), calls of one operator are ordered on the stream autograd runs them on, and
        a buffer that keeps its address is what CUDA-graph replays need.  Allocating it per call let the caching allocator
        split the freed block (a 33 GB request failed with 33 GB cached but fragmented, round 2)."""
        key = str(device)
        with self._lock:
            buf = self._ws.get(key)
            if buf is None or buf.numel() < nbytes:
                self._ws[key] = None
                del buf
                if not torch.cuda.is_current_stream_capturing():
                    torch.cuda.empty_cache() if nbytes > (1 << 30) else None
                buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
                self._ws[key] = buf
        return buf

    def release_memory(self):
        """Drop the idle wavefield-history buffers and the workspace held for reuse."""
        with self._lock:
            for arena in self._history_arena.values():
                arena.clear()
            self._ws.clear()

    # -- the operator -----------------------------------------------------------------------------
    def forward(self, v):
        if self.normalize:
            v = self.v_denorm_func(v)
        s = _WaveSolve.apply(v, self)
        return self.s_norm_func(s) if self.normalize else s

    def misfit_stats(self, v, y, mask=None, return_seismograms=False):
        """Per-model L1 data-misfit sums without materialising the residual as torch tensors (opt-in extension,
        SURVEY.md 8f-1): returns a (B, 2) float64 tensor [sum |y - F(v)| * mask, sum mask], differentiable w.r.t. v
        through column 0.  Needs an identity s_norm_func (every config of the reference uses s_normalize_none)."""
        if self.normalize:
            probe = torch.zeros(1)
            if self.s_norm_func is not None and self.s_norm_func(probe) is not probe:
                raise ValueError("rdfwi: the fused misfit needs an identity s_norm_func (s_normalize_none)")
            v = self.v_denorm_func(v)
        out = _WaveMisfit.apply(v, y, mask, self, bool(return_seismograms))
        return out if return_seismograms else out[0]

    def misfit(self, v, y, mask=None, return_seismograms=False):
        """The reference's observation loss (core/losses.py:15-40) of the modelled data of `v` against `y`:
        per-model mean |y - F(v)| over the observed samples (mask == 1; all samples when mask is None), shape (B,),
        float32, differentiable w.r.t. v.  Equivalent to ``LossCalculator.observation_loss(op(v), y, mask)``."""
        out = self.misfit_stats(v, y, mask, return_seismograms)
        stats = out[0] if return_seismograms else out
        loss = (stats[:, 0] / stats[:, 1].clamp(min=1.0)).to(torch.float32)
        return (loss, out[1]) if return_seismograms else loss

    def pairs_per_gradient(self, B, nz, nx):
        """forward+adjoint cell-update pairs of one gradient evaluation."""
        c = self.ctx
        return B * len(np.atleast_1d(c["sx"])) * (nz + 2 * c["nbc"]) * (nx + 2 * c["nbc"]) * c["nt"]
