"""Short forward+adjoint run for ncu: B models of the OpenFWI shape, few time levels.

    python tools/profile_step.py [B] [nt] [key=value ...]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize  # noqa: E402
from red_diffeq_b200.utils import synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ctx = dict(synthetic.PDE_OPENFWI)
ctx["nt"] = nt
op = FWIForward(ctx, "cuda:0", normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
for kv in [a for a in sys.argv[3:] if "=" in a]:
    k, v = kv.split("=")
    op.set_option(k, int(v))
v = torch.tensor(synthetic.velocity_models(B, 70, 70), device="cuda:0", requires_grad=True)
for _ in range(2):
    v.grad = None
    seis = op(v)
    seis.backward(torch.ones_like(seis))
torch.cuda.synchronize()
print("ok", float(v.grad.abs().sum()))
