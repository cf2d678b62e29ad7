"""Stress the run-to-run determinism of one configuration (debug): python tools/det_stress.py fixture iterations key=value ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import Golden
from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize
name, N = sys.argv[1], int(sys.argv[2])
g = Golden(name)
op = FWIForward(g.fresh_ctx(), "cuda:0", sample_temporal=g.sample_temporal, sample_spatial=g.sample_spatial, normalize=g.normalize,
                v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
op.set_option("engine", 2)
for kv in sys.argv[3:]:
    op.set_option(kv.split("=")[0], int(kv.split("=")[1]))
shape = (g.v.shape[0], len(op.ctx["sx"]), -(-g.ctx["nt"] // g.sample_temporal), len(op.ctx["gx"]))
cot = torch.tensor(g.cotangent(shape), device="cuda:0")
s0 = g0 = None
bad_s = bad_g = 0
first_bad = None
for it in range(N):
    v = torch.tensor(g.v, device="cuda:0", requires_grad=True)
    s = op(v)
    s.backward(cot)
    torch.cuda.synchronize()
    ws = next(iter(op._ws.values()))
    if s0 is None:
        s0, g0 = s.detach().clone(), v.grad.clone()
        ws0 = ws.clone()
    else:
        if not torch.equal(v.grad, g0) and first_bad is None:
            d = (ws != ws0).nonzero().flatten()
            f = d // 4
            print("workspace bytes differing:", d.numel(), "float offsets:", f.unique()[:40].tolist(), "...", f.unique()[-5:].tolist())
            w32 = ws.view(torch.float32); w032 = ws0.view(torch.float32)
            for q in f.unique()[:12].tolist():
                print("   float", q, float(w032[q]), float(w32[q]))
        ds, dg = not torch.equal(s.detach(), s0), not torch.equal(v.grad, g0)
        bad_s += ds; bad_g += dg
        if (ds or dg) and first_bad is None:
            diff = (v.grad != g0).nonzero()
            first_bad = (it, int(ds), int(dg), diff.shape[0], diff[:3].tolist(), float((v.grad - g0).abs().max()), float(g0.abs().max()))
plan = next(iter(op._plans.values()))
print(f"{name} {sys.argv[3:]} C={plan.get('cluster_size_last')} R={plan.get('cluster_rows_last')} mode={plan.get('adj_split')}: "
      f"seismograms differing {bad_s}/{N - 1}, gradients differing {bad_g}/{N - 1}; first: {first_bad}")
