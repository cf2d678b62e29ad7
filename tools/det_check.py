"""Run-to-run determinism of the cluster engine (debug): python tools/det_check.py [iterations]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import Golden, rel_l2
from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for name, optlist in [("openfwi", ({}, {"cluster_rows": 7}, {"cluster_rows": 13}, {"imaging": 1})), ("marmousi", ({}, {"cluster_rows": 13})),
                      ("tiny_default", ({"cluster_rows": 4, "cluster_size": 3}, {"cluster_rows": 4, "cluster_size": 5}))]:
    g = Golden(name)
    for opts in optlist:
        op = FWIForward(g.fresh_ctx(), "cuda:0", sample_temporal=g.sample_temporal, sample_spatial=g.sample_spatial, normalize=g.normalize,
                        v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
        op.set_option("engine", 2)
        for k, v in opts.items():
            op.set_option(k, v)
        shape = (g.v.shape[0], len(op.ctx["sx"]), -(-g.ctx["nt"] // g.sample_temporal), len(op.ctx["gx"]))
        cot = torch.tensor(g.cotangent(shape), device="cuda:0")
        seis0 = grad0 = None
        bad_s = bad_g = 0
        worst = 0.0
        for it in range(N):
            v = torch.tensor(g.v, device="cuda:0", requires_grad=True)
            s = op(v)
            s.backward(cot)
            sn, gn = s.detach().cpu().numpy(), v.grad.cpu().numpy()
            if seis0 is None:
                seis0, grad0 = sn, gn
            else:
                bad_s += not np.array_equal(sn, seis0)
                if not np.array_equal(gn, grad0):
                    bad_g += 1
                    worst = max(worst, rel_l2(gn, grad0))
        plan = next(iter(op._plans.values()))
        print(f"{name:14s} {str(opts):44s} C={plan.get('cluster_size_last')} R={plan.get('cluster_rows_last')} mode={plan.get('adj_split')}"
              f"  seismograms differing {bad_s}/{N - 1}  gradients differing {bad_g}/{N - 1} (worst rel {worst:.2e})"
              f"  vs f32 fixture {rel_l2(grad0, g.grad_f32):.2e}  vs f64 {rel_l2(grad0, g.grad_f64):.2e}")
        op.release_memory()
