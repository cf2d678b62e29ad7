"""Short forward+adjoint run of the per-level (tiled) engine on a large grid, for ncu:
    python tools/profile_tile.py [n] [ns] [nt] [key=value ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize  # noqa: E402
from red_diffeq_b200.utils import synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 16
nt = int(sys.argv[3]) if len(sys.argv) > 3 else 160
ctx = dict(n_grid=n, nt=nt, dx=10.0, dt=0.001, nbc=120, f=15.0, sz=10, gz=10, ng=n, ns=ns)
op = FWIForward(ctx, "cuda:0", normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
op.set_option("engine", 1)
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    op.set_option(k, int(v))
v = torch.tensor(synthetic.velocity_models(1, n, n), device="cuda:0", requires_grad=True)
seis = op(v)
seis.backward(torch.ones_like(seis))
torch.cuda.synchronize()
print("ok", float(v.grad.abs().sum()))
