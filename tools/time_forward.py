"""Times forward-only (no history) and forward-with-history for a batch: python tools/time_forward.py B [key=value ...]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize
from red_diffeq_b200.utils import synthetic
B = int(sys.argv[1])
ctx = dict(synthetic.PDE_OPENFWI)
op = FWIForward(ctx, "cuda:0", normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
for kv in sys.argv[2:]:
    k, v = kv.split("="); op.set_option(k, int(v))
v = torch.tensor(synthetic.velocity_models(B, 70, 70), device="cuda:0")
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def fwd_nograd():
    with torch.no_grad(): op(v)
def fwd_hist():
    vv = v.detach().requires_grad_(True); op(vv)
print(f"B={B} opts={sys.argv[2:]} forward no-history {timed(fwd_nograd):.2f} ms, with history {timed(fwd_hist):.2f} ms")
