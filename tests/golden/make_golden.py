"""Generate the golden fixtures in tests/golden/ by running the *reference's own* solver.

    python tests/golden/make_golden.py            # needs /root/reference (build container only)

The reference ships no golden vectors or tests for this path, so parity is pinned by
importing red_diffeq/solvers/pde.py from the read-only reference checkout (by file path: the package
itself needs ml_collections / accelerate, which are not installed), running it on CPU in fp32 and in
fp64 (torch default dtype switched), and storing inputs + outputs.  The GPU box has no /root/reference;
tests there read only these .npz files.

Per case the file holds: ctx (json), v (the tensor handed to the operator), the operator flags, the
fp32 seismograms (full or sub-sampled + checksums), and d sum(seis*cot)/dv from the reference's autograd
in fp32 and fp64 for the seeded cotangent `cot_seed`.
"""
import importlib.util
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from red_diffeq_b200.utils import synthetic  # noqa: E402
from red_diffeq_b200.utils.data_trans import s_normalize_none, v_denormalize  # noqa: E402

REF = "/root/reference/red_diffeq/solvers/pde.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("reference_pde", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_reference(ref, ctx, v, cot, dtype, normalize, sample_temporal, sample_spatial):
    torch.set_default_dtype(dtype)
    try:
        op = ref.FWIForward(dict(ctx), "cpu", sample_temporal=sample_temporal, sample_spatial=sample_spatial,
                            normalize=normalize, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
        vt = torch.tensor(v, dtype=dtype, requires_grad=True)
        seis = op(vt)
        (seis * torch.tensor(cot, dtype=dtype)).sum().backward()
        return seis.detach().numpy(), vt.grad.numpy()
    finally:
        torch.set_default_dtype(torch.float32)


def make_case(ref, name, ctx, v, normalize=True, sample_temporal=1, sample_spatial=1.0, cot_seed=4242, seis_stride=1):
    t0 = time.time()
    probe = ref.FWIForward(dict(ctx), "cpu", sample_temporal=sample_temporal, sample_spatial=sample_spatial, normalize=False)
    ns, nrec = len(probe.ctx["sx"]), len(probe.ctx["gx"])
    nt_out = (ctx["nt"] + sample_temporal - 1) // sample_temporal
    cot = synthetic.cotangent((v.shape[0], ns, nt_out, nrec), seed=cot_seed)
    seis32, grad32 = run_reference(ref, ctx, v, cot, torch.float32, normalize, sample_temporal, sample_spatial)
    seis64, grad64 = run_reference(ref, ctx, v.astype(np.float64), cot.astype(np.float64), torch.float64, normalize,
                                   sample_temporal, sample_spatial)
    out = dict(
        ctx=json.dumps({k: (list(map(float, val)) if isinstance(val, (list, tuple, np.ndarray)) else val) for k, val in ctx.items()}),
        v=v, normalize=normalize, sample_temporal=sample_temporal, sample_spatial=sample_spatial,
        cot_seed=cot_seed, cot_checksum=np.float64(cot.astype(np.float64).sum()),
        seis_stride=seis_stride, seis_f32=seis32[:, :, ::seis_stride, :],
        seis_sum=seis32.astype(np.float64).sum(axis=(2, 3)), seis_sumsq=(seis32.astype(np.float64) ** 2).sum(axis=(2, 3)),
        seis_f64_rel_err=np.float64(np.linalg.norm(seis32 - seis64) / np.linalg.norm(seis64)),
        grad_f32=grad32, grad_f64=grad64,
    )
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: seis {seis32.shape} grad {grad32.shape} fp32-vs-fp64 seis {out['seis_f64_rel_err']:.2e} "
          f"grad {np.linalg.norm(grad32 - grad64) / np.linalg.norm(grad64):.2e}  ({time.time() - t0:.1f}s)")


def main():
    ref = load_reference()
    rng = np.random.default_rng(synthetic.SEED)

    tiny = dict(n_grid=16, nt=130, dx=10.0, dt=0.001, nbc=8, f=25.0, sz=10, gz=10, ng=16, ns=3)
    v = (1500.0 + 3000.0 * rng.random((2, 1, 12, 16))).astype(np.float32)
    make_case(ref, "tiny_default", tiny, v, normalize=False)

    custom = dict(n_grid=16, nt=131, dx=10.0, dt=0.001, nbc=9, f=25.0, sz=20, gz=10, ng=16, ns=3,
                  sx=[2, 7.5, 13], gx=[0, 1, 1, 5, 9, 15])
    vn = rng.uniform(-1.2, 1.0, size=(3, 1, 10, 16)).astype(np.float32).clip(-1, 1)  # ties at the minimum (1500 m/s)
    make_case(ref, "tiny_custom", custom, vn, normalize=True, sample_temporal=3)

    half = dict(n_grid=20, nt=140, dx=10.0, dt=0.001, nbc=10, f=25.0, sz=10, gz=30, ng=20, ns=4)
    vn = synthetic.velocity_models(2, 14, 20, seed=11)
    make_case(ref, "tiny_half_receivers", half, vn, normalize=True, sample_temporal=2, sample_spatial=0.5)

    vn = synthetic.velocity_models(1, 70, 70)
    make_case(ref, "openfwi", dict(synthetic.PDE_OPENFWI), vn, normalize=True)

    vn = synthetic.velocity_models(1, 70, 190, seed=synthetic.SEED + 7)
    make_case(ref, "marmousi", dict(synthetic.PDE_MARMOUSI), vn, normalize=True, seis_stride=5)


if __name__ == "__main__":
    main()
