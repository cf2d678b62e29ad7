"""Dev check of the resident (tensor-memory) adjoint against the split adjoint and the reference fixtures, then timing.

usage: python tools/check_resident.py [--time] [--cases a,b,...]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Golden, rel_l2  # noqa: E402
from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize  # noqa: E402


def make_op(g, **opts):
    op = FWIForward(g.fresh_ctx(), "cuda:0", sample_temporal=g.sample_temporal, sample_spatial=g.sample_spatial,
                    normalize=g.normalize, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
    for k, v in opts.items():
        op.set_option(k, v)
    return op


def pget(op, key):
    return next(iter(op._plans.values())).get(key)


def run(op, v_np, cot_np):
    v = torch.tensor(v_np, device="cuda:0", requires_grad=True)
    seis = op(v)
    (seis * torch.tensor(cot_np, device="cuda:0")).sum().backward()
    torch.cuda.synchronize()
    return seis.detach().cpu().numpy(), v.grad.cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--cases", default="tiny_default,tiny_custom,tiny_half_receivers,openfwi,marmousi")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--modes", default="1,2")
    ap.add_argument("--opt", action="append", default=[], help="key=value plan option for the timed runs")
    args = ap.parse_args()
    ok = True
    for name in [c for c in args.cases.split(",") if c]:
        g = Golden(name)
        base = make_op(g, engine=2, imaging=1, cluster_rows=13 if os.environ.get("RDFWI_LIB") else 0)
        shape = (g.v.shape[0], len(base.ctx["sx"]), -(-g.ctx["nt"] // g.sample_temporal), len(base.ctx["gx"]))
        cot = g.cotangent(shape)
        s1, g1 = run(base, g.v, cot)
        for extra in ({}, {"cluster_rows": 13}, {"cluster_rows": 7}, {"cluster_rows": 4}):
            try:
                res = make_op(g, engine=2, imaging=2, **extra)
                s2, g2 = run(res, g.v, cot)
            except Exception as e:  # configuration does not fit this grid
                print(f"{name:22s} {extra}: skipped ({str(e)[:60]})")
                continue
            e_fix, e_split = rel_l2(g2, g.grad_f32), rel_l2(g2, g1)
            good = np.array_equal(s1, s2) and e_fix <= 1e-4
            ok &= good
            print(f"{name:22s} {str(extra):22s} resident vs fixture {e_fix:.2e}  vs split {e_split:.2e}  split vs fixture {rel_l2(g1, g.grad_f32):.2e}"
                  f"  C={pget(res, 'cluster_size_last')} R={pget(res, 'cluster_rows_last')} mode={pget(res, 'adj_split')}  {'ok' if good else 'FAIL'}")
            del res
        del base
    if args.time:
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        g = Golden("openfwi")
        from red_diffeq_b200.utils import synthetic
        B = args.batch
        vn = synthetic.velocity_models(B, g.v.shape[2], g.v.shape[3], seed=1)
        extra = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.opt}
        gref = None
        for mode in [int(m) for m in args.modes.split(",")]:
            op = make_op(g, imaging=mode, **extra)
            op.set_option("timing", 1)
            v = torch.tensor(vn, device="cuda:0", requires_grad=True)
            seis = op(v)
            cot = torch.randn(seis.shape, device=seis.device, generator=torch.Generator(device=seis.device).manual_seed(7))
            for it in range(4):
                v.grad = None
                del seis
                seis = op(v)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                (seis * cot).sum().backward()
                torch.cuda.synchronize()
                t1 = time.perf_counter()
            grad = v.grad.clone()
            print(f"   seismogram bit checksum {int(seis.view(torch.int32).to(torch.int64).sum().item())}  gradient sum {float(grad.double().sum()):.9e}")
            print(f"B={B} imaging={mode} {extra}: backward {1e3 * (t1 - t0):.2f} ms  adj_split={pget(op, 'adj_split')}  C={pget(op, 'cluster_size_last')} R={pget(op, 'cluster_rows_last')}"
                  f"  kernel us/launch: " + ", ".join(f"{k} {pget(op, 'us_' + k) / max(1, pget(op, 'n_' + k)):.0f}" for k in ("forward", "adjoint_field", "imaging", "adjoint_resident")))
            if gref is None:
                gref = grad
            else:
                print("   resident vs split gradient rel-L2", float((grad - gref).norm() / gref.norm()))
            del op, v, seis, cot
            torch.cuda.empty_cache()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
