"""Few-shot workloads (one model, 5 shots: BASELINE configs[0] and the configs[2] shape): time one gradient for every
cluster configuration of the cluster-resident time loop (rows marched per thread x CTAs per cluster) and for the
automatic choice.  Usage: python tools/few_shots.py [openfwi|marmousi] [B]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize  # noqa: E402
from red_diffeq_b200.utils import synthetic  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "openfwi"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = dict(synthetic.PDE_OPENFWI if kind == "openfwi" else synthetic.PDE_MARMOUSI)
nz, nx = (70, 70) if kind == "openfwi" else (70, 190)
dev = torch.device("cuda:0")
vn = torch.tensor(synthetic.velocity_models(B, nz, nx), device=dev)
configs = [(0, 0), (13, 0), (13, 8), (7, 0), (7, 16), (4, 0), (4, 16)]
if len(sys.argv) > 3:   # "rows:cluster_size,..." e.g. 13:0,13:5,13:6
    configs = [tuple(int(x) for x in c.split(":")) for c in sys.argv[3].split(",")]
for rows, csize in configs:
    op = FWIForward(dict(ctx), dev, normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
    op.set_option("cluster_rows", rows)
    op.set_option("cluster_size", csize)
    v = vn.clone().requires_grad_(True)
    try:
        seis = op(v)
        cot = torch.randn_like(seis)
        seis.backward(cot)
        torch.cuda.synchronize()
    except Exception as e:  # configuration does not fit this grid
        print(json.dumps({"workload": kind, "B": B, "rows": rows, "cluster_size": csize, "error": str(e)[:120]}), flush=True)
        continue
    plan = op._plan_for(nz, nx, dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fwd = bwd = 0.0
    reps = 5 if B < 16 else 2
    for _ in range(reps):
        v.grad = None
        ev[0].record()
        seis = op(v)
        ev[1].record()
        seis.backward(cot)
        ev[2].record()
        torch.cuda.synchronize()
        fwd += ev[0].elapsed_time(ev[1]) / reps
        bwd += ev[1].elapsed_time(ev[2]) / reps
    print(json.dumps({"workload": kind, "B": B, "rows": rows, "cluster_size": csize,
                      "ran": [plan.get("cluster_size_last"), plan.get("cluster_rows_last")],
                      "forward_ms": round(fwd, 3), "adjoint_ms": round(bwd, 3), "gradient_ms": round(fwd + bwd, 3),
                      "pairs_per_s": op.pairs_per_gradient(B, nz, nx) / ((fwd + bwd) * 1e-3)}), flush=True)
    op.release_memory()
