"""GPU: schedule-perturbation test of the cluster-resident time loop -- the substitute for `compute-sanitizer --tool
racecheck`, which is closed on this GPU pool.

k_fwd_cluster has no cluster-wide barrier inside its time loop: halo rows travel as st.async stores onto mbarriers that
were armed one level ahead, the write-after-read safety of a halo buffer is an argument about data dependencies
(kernels_cluster.cu, header comment), and phase parities are computed arithmetically.  An unperturbed run only ever sees
one interleaving of the warps and CTAs.  The library's debug option "perturb" injects pseudo-random per-warp delays (up
to ~4 us, a whole level's duration, on a quarter of the (warp, level, site) triples) in front of every synchronisation
point of a level: the halo waits, the sweep with its early halo pushes, the late pushes, the bulk-copy hand-over and the
sampling / cotangent warp.  Every seed gives a different schedule; a missing ordering shows up as a changed bit.

Each configuration (cluster size x rows per thread; forward mode, adjoint mode with the imaging sums formed in the sweep --
accumulators in tensor memory, forward rows staged by cp.async -- and the split adjoint's adjoint-field mode) is run unperturbed once and then
under several seeds: seismograms and gradients must be bit-identical every time.
"""
import numpy as np
import pytest

from conftest import Golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _op(g, rows, csize, imaging=0):
    from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize
    op = FWIForward(g.fresh_ctx(), "cuda:0", sample_temporal=g.sample_temporal, sample_spatial=g.sample_spatial,
                    normalize=g.normalize, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
    op.set_option("engine", 2)
    op.set_option("cluster_rows", rows)
    op.set_option("imaging", imaging)   # 0: imaging sums formed in the adjoint sweep (tensor memory), 1: split adjoint
    if csize:
        op.set_option("cluster_size", csize)
    return op


def _run(op, v_np, cot_np):
    v = torch.tensor(v_np, device="cuda:0", requires_grad=True)
    seis = op(v)
    seis.backward(torch.tensor(cot_np, device="cuda:0"))
    return seis.detach().cpu().numpy(), v.grad.cpu().numpy()


CASES = [  # fixture, rows per thread, cluster size (0 = smallest that fits), seeds, imaging option
    ("tiny_default", 13, 0, 20, 0), ("tiny_default", 4, 3, 20, 0), ("tiny_default", 7, 2, 20, 0), ("tiny_default", 4, 5, 20, 0),
    ("tiny_custom", 7, 2, 20, 0), ("tiny_custom", 13, 4, 20, 0), ("tiny_half_receivers", 4, 5, 20, 0), ("tiny_half_receivers", 7, 3, 20, 0),
    ("tiny_half_receivers", 13, 2, 20, 0),
    ("openfwi", 13, 0, 6, 0), ("openfwi", 7, 0, 6, 0), ("openfwi", 4, 0, 6, 0), ("openfwi", 13, 6, 4, 0),
    ("marmousi", 13, 0, 6, 0), ("marmousi", 7, 0, 6, 0), ("marmousi", 13, 8, 4, 0),
    ("tiny_default", 5, 2, 20, 0), ("openfwi", 5, 0, 6, 0), ("marmousi", 5, 16, 4, 0),
    ("tiny_default", 13, 0, 10, 1), ("tiny_custom", 7, 2, 10, 1), ("openfwi", 13, 0, 4, 1), ("marmousi", 7, 0, 4, 1),
]


@pytest.mark.parametrize("name,rows,csize,seeds,imaging", CASES)
def test_perturbed_schedules_are_bit_identical(name, rows, csize, seeds, imaging):
    g = Golden(name)
    op = _op(g, rows, csize, imaging)
    ns, nrec = len(op.ctx["sx"]), len(op.ctx["gx"])
    cot = g.cotangent((g.v.shape[0], ns, -(-g.ctx["nt"] // g.sample_temporal), nrec))
    s0, g0 = _run(op, g.v, cot)
    plan = op._plan_for(g.v.shape[2], g.v.shape[3], torch.device("cuda:0"))
    assert plan.get("cluster_rows_last") == rows and (csize == 0 or plan.get("cluster_size_last") == csize)
    assert np.array_equal(s0[:, :, ::g.seis_stride, :], g.seis_f32)          # the unperturbed run is the reference's
    for seed in range(1, seeds + 1):
        op.set_option("perturb", 7919 * seed + rows)
        s1, g1 = _run(op, g.v, cot)
        assert np.array_equal(s0, s1), f"seismograms changed under perturbation seed {seed}"
        assert np.array_equal(g0, g1), f"gradient changed under perturbation seed {seed}"
    op.set_option("perturb", 0)
    op.release_memory()


def test_perturbation_really_changes_the_schedule():
    """The option is not a no-op: a perturbed run takes measurably longer than an unperturbed one."""
    g = Golden("openfwi")
    op = _op(g, 13, 0)
    v = torch.tensor(g.v, device="cuda:0")

    def timed():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.no_grad():
            op(v)
            a.record()
            op(v)
            b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b)

    t_plain = timed()
    op.set_option("perturb", 12345)
    t_pert = timed()
    assert t_pert > 1.2 * t_plain, (t_plain, t_pert)
