"""First-light GPU diagnostic: localises mismatches between the CUDA path and the oracle."""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Golden, rel_l2
from oracle import fwi_oracle as fo
from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize

print(torch.cuda.get_device_name(0), torch.cuda.mem_get_info())
OPTS = dict(kv.split("=") for kv in sys.argv[1:])
for name in ["tiny_default", "tiny_custom", "tiny_half_receivers", "openfwi"]:
    g = Golden(name)
    for rows in (1, 2, 4):
        op = FWIForward(g.fresh_ctx(), "cuda:0", sample_temporal=g.sample_temporal, sample_spatial=g.sample_spatial,
                        normalize=g.normalize, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
        op.set_option("rows_per_thread", rows); op.set_option("adj_rows_per_thread", min(rows, 2))
        for k, val in OPTS.items():
            op.set_option(k, int(val))
        v = torch.tensor(g.v, device="cuda:0", requires_grad=True)
        seis = op(v)
        cot = g.cotangent(tuple(seis.shape))
        (seis * torch.tensor(cot, device="cuda:0")).sum().backward()
        torch.cuda.synchronize()
        nz_, nx_ = g.v.shape[2:]
        msg0 = f"[cluster C={op._plan_for(nz_, nx_, torch.device('cuda:0')).get('cluster_size_used')}] "
        s = seis.detach().cpu().numpy()[:, :, ::g.seis_stride]
        same = np.array_equal(s, g.seis_f32)
        msg = msg0 + f"{name} R={rows}: seis rel {rel_l2(s, g.seis_f32):.3e} bit-identical {same}"
        if not same:
            d = np.abs(s - g.seis_f32); idx = np.unravel_index(np.argmax(d), d.shape)
            first_t = np.nonzero(d.max(axis=(0, 1, 3)))[0]
            msg += f" maxdiff {d.max():.3e} at {idx} first bad t {first_t[:3]}"
        msg += f" | grad vs ref32 {rel_l2(v.grad.cpu().numpy(), g.grad_f32):.3e} vs ref64 {rel_l2(v.grad.cpu().numpy(), g.grad_f64):.3e}"
        print(msg, flush=True)
