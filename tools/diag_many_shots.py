"""Debug aid: gradient error of the engines against the CPU oracle for several shot counts / batch sizes."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize
from red_diffeq_b200.utils import synthetic
from oracle import build_oracle, fwi_oracle
build_oracle.build()

def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))

for B, ns, engine, opts in [(1, 7, 1, {}), (1, 8, 1, {}), (1, 11, 1, {}), (2, 11, 1, {}), (2, 11, 1, {"chunk_models": 1}),
                            (2, 11, 1, {"adj_rows_per_thread": 2}), (2, 11, 2, {}), (2, 11, 2, {"adj_mode": 1}), (2, 11, 2, {"u_chunk_shots": 7})]:
    nz, nx = 20, 28
    ctx = dict(n_grid=nx, nt=130, dx=10.0, dt=0.001, nbc=12, f=25.0, sz=10, gz=10, ng=nx, ns=ns)
    sv = fwi_oracle.Survey(dict(ctx), nz, nx)
    vn = synthetic.velocity_models(B, nz, nx, seed=31)
    cot = synthetic.cotangent((B, sv.ns, sv.nt_out, sv.nrec), seed=32)
    op = FWIForward(dict(ctx), "cuda:0", normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
    op.set_option("engine", engine)
    for k, v in opts.items():
        op.set_option(k, v)
    v = torch.tensor(vn, device="cuda:0", requires_grad=True)
    seis = op(v)
    (seis * torch.tensor(cot, device="cuda:0")).sum().backward()
    v_phys = (vn + np.float32(1)) / np.float32(2) * np.float32(3000) + np.float32(1500)
    seis_o, grad_o = fwi_oracle.gradient(sv, v_phys, cot)
    g = v.grad.cpu().numpy() / 1500.0
    print(B, ns, engine, opts, "seis equal", np.array_equal(seis.detach().cpu().numpy(), seis_o), "grad rel", rel(g, grad_o),
          "per model", [rel(g[b], grad_o[b]) for b in range(B)], flush=True)
