import sys; sys.path.insert(0,".")
import torch, numpy as np
from red_diffeq_b200 import FWIForward
for n in (300, 330, 360, 380, 392, 400):
    for ns in (2, 12):
        ctx = dict(n_grid=n, nt=150, dx=10.0, dt=0.001, nbc=120, f=15.0, sz=10, gz=10, ng=n, ns=ns)
        op = FWIForward(dict(ctx), "cuda:0", normalize=False)
        v = torch.full((1,1,n,n), 2000.0, device="cuda:0", requires_grad=True)
        try:
            s = op(v); s.sum().backward(); torch.cuda.synchronize()
            plan = op._plan_for(n, n, torch.device("cuda:0"))
            print(n, ns, "C", plan.get("cluster_size_last"), "R", plan.get("cluster_rows_last"), "adj_split", plan.get("adj_split"), "used", plan.get("cluster_size_used"), flush=True)
        except Exception as e:
            plan = op._plan_for(n, n, torch.device("cuda:0"))
            print(n, ns, "FAILED", str(e)[:150], "C", plan.get("cluster_size_last"), "R", plan.get("cluster_rows_last"), flush=True)
        op.release_memory()
