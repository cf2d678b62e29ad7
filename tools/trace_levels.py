"""Per-warp timeline of the cluster-resident forward kernel (debug): python tools/trace_levels.py [B] [fwd|adj] [openfwi|marmousi]"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize
from red_diffeq_b200.utils import synthetic
opts = [a for a in sys.argv[1:] if "=" in a]          # key=value plan options (e.g. imaging=2)
sys.argv = [a for a in sys.argv if "=" not in a]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = sys.argv[2] if len(sys.argv) > 2 else "fwd"
kind = sys.argv[3] if len(sys.argv) > 3 else "openfwi"
ctx = dict(synthetic.PDE_OPENFWI if kind == "openfwi" else synthetic.PDE_MARMOUSI); ctx["nt"] = 300
nz, nx = (70, 70) if kind == "openfwi" else (70, 190)
op = FWIForward(ctx, "cuda:0", normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
trace = torch.zeros((16, 4, 16, 6), dtype=torch.int64, device="cuda:0")   # [cta][level][warp][phase]
op.set_option("trace_ptr", trace.data_ptr())
for kv in opts:
    op.set_option(kv.split("=")[0], int(kv.split("=")[1]))
v = torch.tensor(synthetic.velocity_models(B, nz, nx), device="cuda:0", requires_grad=True)
s = op(v); torch.cuda.synchronize()
C = op._plan_for(nz, nx, torch.device("cuda:0")).get("cluster_size_last")   # what the launch ran (wide clusters for few shots)
if mode == "adj":      # trace the adjoint-field launch (k_fwd_cluster<ADJ>) instead
    trace.zero_()
    s.backward(torch.ones_like(s)); torch.cuda.synchronize()
tr = trace.cpu().numpy().astype(np.int64)
names = ["start", "after halo wait", "after sweep", "after epilogue", "at barrier", "after barrier"]
for cta in range(C):
    t0 = tr[cta, 1, :, 0].min()   # level 101 as origin
    print(f"CTA {cta}: level 101, cycles relative to the earliest warp start; per warp: " + ", ".join(names[1:]))
    for w in range(16):
        print(f"  warp {w:2d} start {tr[cta,1,w,0]-t0:6d} | " + " ".join(f"{tr[cta,1,w,p]-tr[cta,1,w,0]:6d}" for p in range(1, 6)))
    print(f"  level period (warp 0 start-to-start): {tr[cta,2,0,0]-tr[cta,1,0,0]} cycles")
