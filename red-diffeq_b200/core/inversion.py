"""Inversion-loop driver around the B200 operator: the loop body of the reference's
``InversionEngine.optimize`` (red_diffeq/core/inversion.py:26-129) with the per-iteration host round trips removed.

Same call shape and same results format as the reference::

    engine = InversionEngine(regularization="tv")                       # or regularizer=<callable>, see below
    mu_result, results_per_model = engine.optimize(mu, mu_true, y, fwi_forward, ts=300, lr=0.03, reg_lambda=0.01,
                                                   noise_std=0.0, missing_number=0, regularization="tv")

What is different from the reference's loop (none of it changes the mathematics):
  * the data term is ``fwi_forward.misfit`` (rdfwi_misfit_l1: loss + cotangent in one pass) instead of the operator
    followed by ~10 elementwise kernels (core/losses.py:27-40);
  * the six ``.cpu().numpy()`` calls per iteration (core/inversion.py:96-101) and the ``.item()`` calls of the progress bar
    (:103-113) are gone: losses and metrics are written into (ts, B) device arrays and fetched ONCE after the loop;
  * with ``cuda_graph=True`` one iteration -- noise, forward, misfit, regulariser, adjoint, Adam, clamp, cosine LR,
    metrics -- is captured once in a CUDA graph and replayed ts times: no Python, no launch gaps between the ~60 kernels.
    Adam runs with ``capturable=True`` and the cosine schedule (CosineAnnealingLR with eta_min = 0 in closed form,
    lr_k = lr/2 (1 + cos(pi k / ts))) lives in a device scalar.

The denoiser of the diffusion regulariser (a stock-PyTorch U-Net) is out of this repository's scope: hand the reference's
GaussianDiffusion object to the constructor (``InversionEngine(diffusion_model, ...)``, as the reference does) and
regularization='diffusion' calls it through ``regularization.diffusion.REDDiffEq`` (no_grad, batched patches); or pass any
callable mu -> reg_loss (B,) or (reg_loss, time_tensor) as ``regularizer``.  'tv' and 'l2' (regularization/benchmark.py:4-37)
are restated here because they are five lines.
"""
import math
import time

import torch

from ..utils.data_trans import add_noise_to_seismic, missing_trace, v_normalize


def total_variation_loss(mu):
    """regularization/benchmark.py:4-19"""
    dx = (mu[:, :, :, 1:] - mu[:, :, :, :-1]).abs()
    dy = (mu[:, :, 1:, :] - mu[:, :, :-1, :]).abs()
    return dx.reshape(dx.shape[0], -1).mean(dim=1) + dy.reshape(dy.shape[0], -1).mean(dim=1)


def tikhonov_loss(mu):
    """regularization/benchmark.py:22-37"""
    dx = mu[:, :, :, 1:] - mu[:, :, :, :-1]
    dy = mu[:, :, 1:, :] - mu[:, :, :-1, :]
    return (dx ** 2).reshape(dx.shape[0], -1).mean(dim=1) + (dy ** 2).reshape(dy.shape[0], -1).mean(dim=1)


class InversionEngine:
    """regularization: None | 'tv' | 'l2' | 'diffusion' | 'hybrid' (the last two need `regularizer`).
    regularizer:    callable mu -> reg_loss (B,) or (reg_loss, time_tensor); overrides the built-in ones.
    ssim_loss:      optional callable (pred01, true01) -> scalar, evaluated per model like core/metrics.py:41-44
                    (None: the 'ssim' history is NaN -- the SSIM module is not part of the hot path).
    fused_misfit:   use fwi_forward.misfit (default) or the operator + torch loss ops.
    cuda_graph:     None = automatic (on for the built-in regularisers and for the diffusion regulariser built from
                    `diffusion_model`; off for a user callable and on a sharded operator), True / False.  A graphed run
                    draws its x0 noise / timesteps from the same seeded generator but at other Philox offsets than an eager
                    run (the two warm-up iterations draw too): seeded runs reproduce themselves, not the eager stream --
                    pass cuda_graph=False to draw exactly what the reference's loop draws.
    overlap_regularizer: evaluate the regulariser on a second CUDA stream while the operator's forward kernel runs
                    (SURVEY.md 5 / 8e "overlap with the U-Net regulariser"): the two depend only on x0.  One model of the
                    reference's configs leaves 68 of 148 SMs idle during the solve -- room for the U-Net's kernels.  Pays
                    only inside a CUDA graph (eager, the one Python thread issues the ~600 U-Net launches back to back and
                    the solver's launches queue behind them: measured 12.7 vs 12.2 ms per Marmousi iteration eager, 8.7 vs
                    10.7 ms graphed).  None = automatic (on when the iteration is graphed and the regulariser is the
                    diffusion one), True / False."""

    def __init__(self, diffusion_model=None, ssim_loss=None, regularization=None, use_time_weight=False,
                 sigma_x0=0.0001, fixed_timestep=None, *, regularizer=None, fused_misfit=True, cuda_graph=None,
                 overlap_regularizer=None):
        self.diffusion_model = diffusion_model
        self.ssim_loss = ssim_loss
        self.regularization = regularization
        if regularizer is None and diffusion_model is not None:   # like the reference's RegularizationMethod('diffusion', model)
            from ..regularization.diffusion import REDDiffEq
            self._red = REDDiffEq(diffusion_model, use_time_weight=use_time_weight, sigma_x0=sigma_x0, fixed_timestep=fixed_timestep)
        else:
            self._red = None
        self.regularizer = regularizer
        self.sigma_x0 = sigma_x0
        self.fused_misfit = fused_misfit
        self.cuda_graph = cuda_graph
        self.overlap_regularizer = overlap_regularizer
        self.used_cuda_graph = False
        self.last_loop_seconds = None
        self.device = getattr(diffusion_model, "device", None)

    @staticmethod
    def _tick(device):
        if device.type == "cuda":
            torch.cuda.synchronize(device)
        return time.perf_counter()

    # ------------------------------------------------------------------------------------------
    def _resolve_regularizer(self, regularization):
        if regularization not in ["diffusion", "l2", "tv", "hybrid", None]:
            raise ValueError(f"Unknown regularization: {regularization}")          # core/inversion.py:32-33
        if self.regularizer is not None:
            fn, builtin = self.regularizer, False
        elif regularization == "diffusion" and self._red is not None:
            fn, builtin = self._red, False   # stock-PyTorch denoiser, called under no_grad with batched patches
        elif regularization == "tv":
            fn, builtin = total_variation_loss, True
        elif regularization == "l2":
            fn, builtin = tikhonov_loss, True
        elif regularization is None:
            fn, builtin = (lambda mu: torch.zeros(mu.shape[0], device=mu.device, dtype=mu.dtype)), True
        else:
            raise ValueError(f"regularization={regularization!r} runs on the stock PyTorch path: pass the reference's "
                             "RegularizationMethod(...).get_reg_loss as `regularizer`")

        takes_generator = fn is self._red and fn is not None   # REDDiffEq draws its timesteps / noise from a generator

        def call(mu, generator=None):
            out = fn(mu, generator=generator) if takes_generator else fn(mu)
            return out[0] if isinstance(out, tuple) else out
        call.takes_generator = takes_generator
        return call, builtin

    def optimize(self, mu, mu_true, y, fwi_forward, ts=300, lr=0.03, reg_lambda=0.01, noise_std=0.0,
                 noise_type="gaussian", missing_number=0, regularization=None, mask=None):
        if mu.shape[0] != y.shape[0]:
            raise ValueError("Batch size mismatch between velocity and seismic data")
        # the reference perturbs x0 only when optimize() ITSELF is passed regularization='diffusion' (core/inversion.py:71);
        # the constructor's regularization still selects the regulariser when the argument is None (:38-44)
        noise_x0 = regularization == "diffusion"
        if regularization is None:
            regularization = self.regularization
        reg_fn, builtin = self._resolve_regularizer(regularization)
        if fwi_forward is None or not callable(fwi_forward):
            raise ValueError("fwi_forward must be a callable forward modeling function")
        device = torch.device(self.device if self.device is not None else fwi_forward.device)
        fwi_forward = fwi_forward.to(device)
        B = mu.shape[0]
        mu = mu.float().clone().detach().to(device).requires_grad_(True)
        mu_true_n = v_normalize(mu_true.float().to(device))                           # core/metrics.py:29
        multi_rank = getattr(fwi_forward, "world_size", 1) > 1            # sharded operator: collectives stay out of graphs
        y = add_noise_to_seismic(y.to(device), noise_std, noise_type=noise_type)      # core/inversion.py:64-67
        y, trace_mask = missing_trace(y, missing_number, return_mask=True)
        mask = trace_mask if mask is None else mask.to(device) * trace_mask
        y, mask = y.contiguous().float(), mask.contiguous().float()
        gen = None
        if multi_rank:
            # Every rank holds the replicated model and must see the SAME perturbed data, trace mask, x0 noise and
            # diffusion timesteps -- each rank's own random stream would make the all-reduced misfit sums and gradients
            # mix inconsistent data and let the replicated mu drift apart.  Rank 0's draws are broadcast; the per-iteration
            # randomness comes from a generator every rank seeds with the same (rank-0) seed.
            import torch.distributed as dist
            group = getattr(fwi_forward, "group", None)
            src = dist.get_global_rank(group, 0) if group is not None else 0
            if noise_std > 0 or missing_number > 0:
                dist.broadcast(y, src=src, group=group)
                dist.broadcast(mask, src=src, group=group)
            stochastic = noise_x0 or not builtin
            if stochastic:
                if not builtin and not reg_fn.takes_generator:
                    raise ValueError("a sharded operator needs identical random draws on every rank: a user regulariser that "
                                     "draws random numbers cannot be synchronised -- use regularization='diffusion' with the "
                                     "engine's diffusion_model (REDDiffEq takes a generator), or a deterministic regulariser")
                seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=device)
                dist.broadcast(seed, src=src, group=group)
                gen = torch.Generator(device=device)
                gen.manual_seed(int(seed.item()))
        fused = self.fused_misfit and hasattr(fwi_forward, "misfit")   # any other callable operator: torch loss ops
        graphable = builtin or reg_fn.takes_generator     # stock regularisers and REDDiffEq (denoiser under no_grad)
        use_graph = (graphable and not multi_rank if self.cuda_graph is None else bool(self.cuda_graph)) and device.type == "cuda"

        overlap_reg = ((reg_fn.takes_generator and use_graph) if self.overlap_regularizer is None else bool(self.overlap_regularizer)) and device.type == "cuda"
        reg_stream = torch.cuda.Stream(device=device) if overlap_reg else None

        lr_t = torch.tensor(float(lr), device=device)
        step_t = torch.zeros(1, dtype=torch.long, device=device)
        optimizer = torch.optim.Adam([mu], lr=lr_t if use_graph else lr, capturable=use_graph)
        scheduler = None if use_graph else torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=ts, eta_min=0.0)
        names = ["total_losses", "obs_losses", "reg_losses", "ssim", "mae", "rmse"]
        hist = {k: torch.full((ts, B), float("nan"), device=device) for k in names}

        def iteration():
            if noise_x0:                                                              # :71-76
                x0_pred = mu + self.sigma_x0 * torch.randn(mu.shape, device=mu.device, dtype=mu.dtype, generator=gen)
            else:
                x0_pred = mu
            if overlap_reg:   # the regulariser needs x0 only: it runs beside the operator's forward kernel (its backward is
                cur = torch.cuda.current_stream(device)      # run by autograd on the same side stream and joined there)
                reg_stream.wait_stream(cur)
                x0_pred.record_stream(reg_stream)
                with torch.cuda.stream(reg_stream):
                    reg_loss = reg_fn(x0_pred, generator=gen) if gen is not None else reg_fn(x0_pred)
            if fused:
                loss_obs = fwi_forward.misfit(x0_pred[:, :, 1:-1, 1:-1], y, mask)
            else:
                pred = fwi_forward(x0_pred[:, :, 1:-1, 1:-1])                         # :78-79, losses.py:27-36
                loss_obs = ((y - pred).abs() * mask).sum(dim=(1, 2, 3)) / mask.sum(dim=(1, 2, 3)).clamp(min=1.0)
            if overlap_reg:
                cur.wait_stream(reg_stream)
                reg_loss.record_stream(cur)
            else:
                reg_loss = reg_fn(x0_pred, generator=gen) if gen is not None else reg_fn(x0_pred)
            total = loss_obs + reg_lambda * reg_loss                                  # losses.py:54-66
            optimizer.zero_grad(set_to_none=not use_graph)
            total.sum().backward()                                                    # :85-87
            optimizer.step()
            with torch.no_grad():
                mu.clamp_(-1, 1)                                                      # :89-90
                if use_graph:     # CosineAnnealingLR(T_max=ts, eta_min=0) in closed form, on the device
                    nxt = (step_t + 1).to(torch.float32)
                    lr_t.copy_((0.5 * lr * (1.0 + torch.cos(math.pi * nxt / ts))).squeeze(0))
                else:
                    scheduler.step()
                err = mu[:, :, 1:-1, 1:-1] - mu_true_n                                 # core/metrics.py:32-34
                rows = {"total_losses": total.detach(), "obs_losses": loss_obs.detach(), "reg_losses": reg_loss.detach(),
                        "mae": err.abs().mean(dim=(1, 2, 3)), "rmse": (err ** 2).mean(dim=(1, 2, 3)).sqrt()}
                if self.ssim_loss is not None:                                        # core/metrics.py:37-44
                    a, b = (mu[:, :, 1:-1, 1:-1] + 1) / 2, (mu_true_n + 1) / 2
                    rows["ssim"] = torch.stack([self.ssim_loss(a[i:i + 1], b[i:i + 1]).reshape(()) for i in range(B)])
                for k, row in rows.items():
                    hist[k].index_copy_(0, step_t, row.to(torch.float32).unsqueeze(0))
                step_t.add_(1)

        self.used_cuda_graph = use_graph
        if use_graph:
            # Warm-up on a side stream (allocations, lazy optimizer state, the library's per-thread launch configuration),
            # then rewind the state so that the captured iteration is iteration 0.
            mu0 = mu.detach().clone()
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for _ in range(min(2, ts)):   # (a third row of a (1, B) history would be out of bounds)
                    iteration()
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            with torch.no_grad():
                mu.copy_(mu0)
                mu.grad.zero_()
                for st in optimizer.state.values():
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
                lr_t.fill_(float(lr))
                step_t.zero_()
                for h in hist.values():
                    h.fill_(float("nan"))
            # (the operator's wavefield-history arena is left alone: its buffer keeps its address across replays)
            torch.cuda.empty_cache()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):   # records one iteration; nothing runs yet
                iteration()
            t0 = self._tick(device)
            for _ in range(ts):
                graph.replay()
        else:
            t0 = self._tick(device)
            for _ in range(ts):
                iteration()
        self.last_loop_seconds = self._tick(device) - t0   # the ts iterations alone (no warm-up, capture or result fetch)

        fetched = {k: h.cpu().numpy() for k, h in hist.items()}                       # the ONE device -> host transfer
        final_results_per_model = [{k: [fetched[k][t][i] for t in range(ts)] for k in names} for i in range(B)]   # :115-126
        return mu[:, :, 1:-1, 1:-1], final_results_per_model
