"""End-to-end run on files in the reference's formats: .npy family in -> inversion -> per-model .npz out,
the data path of scripts/run_inversion.py:130-216 with the B200 operator and loop driver.

    python tools/invert_family.py --seismic seis.npy --velocity vel.npy --out results/ [--workload openfwi|marmousi]
                                  [--batch 25] [--ts 300] [--reg tv|l2|none] [--sigma 10] [--sample-index i]
    python tools/invert_family.py --synthetic 4 --out results/      # writes a synthetic family first (observed data = our forward)

The diffusion regulariser is not available here (stock-PyTorch U-Net, weights not in the repository): 'tv', 'l2' or none.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from red_diffeq_b200 import FWIForward, InversionEngine, s_normalize_none, v_denormalize  # noqa: E402
from red_diffeq_b200.utils import io, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seismic")
    ap.add_argument("--velocity")
    ap.add_argument("--synthetic", type=int, default=0, help="write a synthetic family of this many models and invert it")
    ap.add_argument("--out", required=True)
    ap.add_argument("--workload", default="openfwi", choices=["openfwi", "marmousi"])
    ap.add_argument("--batch", type=int, default=25)
    ap.add_argument("--ts", type=int, default=300)
    ap.add_argument("--nt", type=int, default=0, help="override the record length (smoke runs)")
    ap.add_argument("--reg", default="tv", choices=["tv", "l2", "none"])
    ap.add_argument("--sigma", type=float, default=None)
    ap.add_argument("--sample-index", type=int, default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    ctx = dict(synthetic.PDE_OPENFWI if args.workload == "openfwi" else synthetic.PDE_MARMOUSI)
    if args.nt:
        ctx["nt"] = args.nt
    nz, nx = (70, 70) if args.workload == "openfwi" else (70, 190)
    sigma = args.sigma if args.sigma is not None else (10.0 if args.workload == "openfwi" else 20.0)   # configs/*: optimization.sigma
    op = FWIForward(dict(ctx), dev, normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
    os.makedirs(args.out, exist_ok=True)
    if args.synthetic:
        vn = torch.tensor(synthetic.velocity_models(args.synthetic, nz, nx), device=dev)
        with torch.no_grad():
            seis = op(vn)
        args.seismic, args.velocity = os.path.join(args.out, "seismic.npy"), os.path.join(args.out, "velocity.npy")
        np.save(args.seismic, seis.cpu().numpy())
        np.save(args.velocity, v_denormalize(vn).cpu().numpy())
    fam = io.Family(args.seismic, args.velocity)
    fam.check_against(op.ctx)
    reg = None if args.reg == "none" else args.reg
    engine = InversionEngine(regularization=reg)
    t0 = time.perf_counter()
    written = []
    for a, b in fam.batches(args.batch, args.sample_index):
        seis, vel = fam.load_batch(a, b, dev)
        init = io.initial_batch(vel, "smoothed", sigma)
        mu, results = engine.optimize(init, vel, seis, op, ts=args.ts, lr=0.03, reg_lambda=0.01, regularization=reg)
        written += io.save_batch_results(a, b, mu, results, init, vel, os.path.join(args.out, "results"))
    torch.cuda.synchronize()
    z = np.load(written[0])
    print(json.dumps({"models": len(written), "iterations": args.ts, "seconds": time.perf_counter() - t0,
                      "cuda_graph": engine.used_cuda_graph, "first_file": written[0],
                      "obs_loss_first_last": [float(z["obs_losses"][0]), float(z["obs_losses"][-1])],
                      "mae_first_last": [float(z["mae"][0]), float(z["mae"][-1])]}), flush=True)


if __name__ == "__main__":
    main()
