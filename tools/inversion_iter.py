"""Seconds per inversion iteration with the B200 operator inside the reference's loop body (BASELINE configs[1]).

    python tools/inversion_iter.py [--batch 64] [--iters 10] [--reg none|tv|l2] [--workload openfwi|marmousi]

The loop body restates the reference's InversionEngine.optimize (core/inversion.py:69-113): forward on the slice
mu[:, :, 1:-1, 1:-1] of the padded leaf, masked L1 data misfit per model (core/losses.py:27-41), a regulariser,
`total.sum().backward()`, Adam, clamp to [-1, 1], cosine LR, and the per-iteration metrics with their host syncs
(MAE / RMSE per model, `.cpu().numpy()` of the three losses).  Regularisers: the reference's non-learned ones
(regularization/benchmark.py: total variation, Tikhonov) or none; the diffusion regulariser is a stock-PyTorch U-Net
whose weights are not part of the repository -- it stays on the PyTorch path and is not timed here.

Prints one JSON line: s/iteration (CUDA events around the whole loop, after 2 warm-up iterations), the solver's share
(forward + backward of the operator alone, measured separately on the same inputs) and pairs/s through the loop.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize  # noqa: E402
from red_diffeq_b200.utils import synthetic  # noqa: E402


def reg_loss(kind, mu):
    if kind == "none":
        return torch.zeros(mu.shape[0], device=mu.device)
    dx = mu[:, :, :, 1:] - mu[:, :, :, :-1]
    dy = mu[:, :, 1:, :] - mu[:, :, :-1, :]
    if kind == "tv":      # regularization/benchmark.py:4-19
        return dx.abs().flatten(1).mean(dim=1) + dy.abs().flatten(1).mean(dim=1)
    return (dx ** 2).flatten(1).mean(dim=1) + (dy ** 2).flatten(1).mean(dim=1)   # Tikhonov, :22-37


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--reg", default="tv", choices=["none", "tv", "l2"])
    ap.add_argument("--workload", default="openfwi", choices=["openfwi", "marmousi"])
    ap.add_argument("--fused-misfit", action="store_true", help="FWIForward.misfit instead of op(v) + torch loss ops")
    ap.add_argument("--driver", default="loop", choices=["loop", "engine", "engine-graph"],
                    help="loop = the reference's loop body restated here (host syncs included); engine = red_diffeq_b200."
                         "InversionEngine (fused misfit, metrics fetched once); engine-graph = the same with the iteration "
                         "captured in a CUDA graph")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    ctx = dict(synthetic.PDE_OPENFWI if args.workload == "openfwi" else synthetic.PDE_MARMOUSI)
    nz, nx = (70, 70) if args.workload == "openfwi" else (70, 190)
    B, ts = args.batch, 300
    op = FWIForward(dict(ctx), dev, normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none).to(dev)
    mu_true = torch.tensor(synthetic.velocity_models(B, nz, nx, seed=synthetic.SEED), device=dev)
    with torch.no_grad():
        y = op(mu_true)                                               # observed data
    mask = torch.ones_like(y)
    mu0 = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(mu_true, (5, 5, 5, 5), mode="replicate"), 11, stride=1)
    mu = torch.nn.functional.pad(mu0, (1, 1, 1, 1), value=0.0).clone().requires_grad_(True)   # scripts/run_inversion.py:156
    if args.driver != "loop":
        import time
        from red_diffeq_b200 import InversionEngine
        reg = None if args.reg == "none" else args.reg
        eng = InversionEngine(regularization=reg, fused_misfit=True, cuda_graph=args.driver == "engine-graph")
        mu_true_phys = v_denormalize(mu_true)

        def run(ts):   # wall clock around optimize(); fixed costs (warm-up, capture, the final fetch) cancel in the difference
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _, res = eng.optimize(mu.detach(), mu_true_phys, y, op, ts=ts, lr=0.03, reg_lambda=0.01, regularization=reg)
            torch.cuda.synchronize()
            return time.perf_counter() - t0, res
        run(2)
        _, res = run(3 * args.iters)
        s_iter = eng.last_loop_seconds / (3 * args.iters)   # the iterations alone: no warm-up, capture or result fetch
        pairs = op.pairs_per_gradient(B, nz, nx)
        print(json.dumps({"metric": "s / inversion iteration", "value": s_iter, "workload": args.workload, "models": B,
                          "regulariser": args.reg, "driver": args.driver, "cuda_graph": eng.used_cuda_graph,
                          "iterations_timed": 3 * args.iters, "pairs_per_s_through_the_loop": pairs / s_iter,
                          "misfit_first_last": [float(res[0]["obs_losses"][0]), float(res[0]["obs_losses"][-1])]}), flush=True)
        return
    opt = torch.optim.Adam([mu], lr=0.03)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=ts, eta_min=0.0)
    hist = {"total": [], "obs": [], "reg": [], "mae": [], "rmse": []}

    def iteration():
        x0 = mu + 1e-4 * torch.randn_like(mu)                                            # :73-74
        if args.fused_misfit:
            loss_obs = op.misfit(x0[:, :, 1:-1, 1:-1], y, mask)
        else:
            pred = op(x0[:, :, 1:-1, 1:-1])                                              # :78
            loss_obs = ((y - pred).abs() * mask).sum(dim=(1, 2, 3)) / mask.sum(dim=(1, 2, 3)).clamp(min=1.0)
        r = reg_loss(args.reg, x0)
        total = loss_obs + 0.01 * r
        opt.zero_grad(set_to_none=True)
        total.sum().backward()                                                           # :86
        opt.step()
        with torch.no_grad():
            mu.data.clamp_(-1, 1)
        sched.step()
        with torch.no_grad():                                                            # :94-107, host syncs included
            err = mu[:, :, 1:-1, 1:-1] - mu_true
            hist["mae"].append(err.abs().flatten(1).mean(dim=1).cpu().numpy())
            hist["rmse"].append(err.pow(2).flatten(1).mean(dim=1).sqrt().cpu().numpy())
            hist["total"].append(total.detach().cpu().numpy())
            hist["obs"].append(loss_obs.detach().cpu().numpy())
            hist["reg"].append(r.detach().cpu().numpy())

    for _ in range(2):
        iteration()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        iteration()
    e1.record()
    torch.cuda.synchronize()
    s_iter = e0.elapsed_time(e1) * 1e-3 / args.iters

    # the operator alone on the same inputs
    v = mu.detach()[:, :, 1:-1, 1:-1].contiguous()
    cot = torch.sign(torch.randn_like(y)) / y[0].numel()
    for _ in range(2):
        vv = v.clone().requires_grad_(True)
        op(vv).backward(cot)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.iters):
        vv = v.clone().requires_grad_(True)
        op(vv).backward(cot)
    e1.record()
    torch.cuda.synchronize()
    s_solver = e0.elapsed_time(e1) * 1e-3 / args.iters
    pairs = op.pairs_per_gradient(B, nz, nx)
    print(json.dumps({"metric": "s / inversion iteration", "value": s_iter, "workload": args.workload, "models": B,
                      "regulariser": args.reg, "fused_misfit": args.fused_misfit, "iterations_timed": args.iters, "solver_s_per_iter": s_solver,
                      "solver_share": s_solver / s_iter, "pairs_per_s_through_the_loop": pairs / s_iter,
                      "misfit_first_last": [float(hist["obs"][0].mean()), float(hist["obs"][-1].mean())],
                      "mae_first_last": [float(hist["mae"][0].mean()), float(hist["mae"][-1].mean())]}), flush=True)


if __name__ == "__main__":
    main()
