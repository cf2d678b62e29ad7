"""GPU: the RED-DiffEq regulariser and loop with the REFERENCE's own U-Net.

The denoiser is the unmodified red_diffeq/models/diffusion.py (Unet + GaussianDiffusion at the sizes of
configs/*/red-diffeq.yaml, random-init: the weights are not in the repository), staged under baseline/_ref by
baseline/stage_reference.py and imported through baseline/ref_loader.py.  Checked here, in fp32 (TF32 convolutions off so
that a different batch size cannot pick a differently-rounded algorithm):
  * this repo's REDDiffEq (no_grad, patches batched into one U-Net call) gives the same loss, the same gradient w.r.t. mu
    and the same timesteps as the reference's RED_DiffEq call pattern, from the same seeded generator -- un-patched
    (OpenFWI, 72 x 72) and patched (Marmousi, 72 x 192 -> three patches);
  * the reference's own InversionEngine.optimize (unmodified core/inversion.py, its LossCalculator, MetricsCalculator,
    SSIM) runs with THIS repo's FWIForward as a drop-in for its operator, and this repo's InversionEngine reproduces its
    iterations.
"""
import numpy as np
import pytest

from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from baseline import ref_loader
    if not ref_loader.available():
        pytest.skip("baseline/_ref not staged (python baseline/stage_reference.py where /root/reference exists)")
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    dm = ref_loader.build_diffusion(torch.device("cuda:0"))
    mods = dict(zip(("reg", "base", "inv", "ssim"), ref_loader.load("red_diffeq.regularization.diffusion", "red_diffeq.regularization.base",
                                                                    "red_diffeq.core.inversion", "red_diffeq.utils.ssim")))
    yield dm, mods
    torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("B,width,use_time_weight", [(3, 70, False), (2, 190, False), (2, 190, True), (4, 70, True)])
def test_call_patterns_agree_with_the_reference_unet(ref, B, width, use_time_weight):
    from red_diffeq_b200 import REDDiffEq
    from red_diffeq_b200.utils import synthetic
    dm, mods = ref
    dev = torch.device("cuda:0")
    mu0 = torch.nn.functional.pad(torch.tensor(synthetic.velocity_models(B, 70, width, seed=5), device=dev), (1, 1, 1, 1))
    theirs = mods["base"].RegularizationMethod("diffusion", dm, use_time_weight=use_time_weight)
    ours = REDDiffEq(dm, use_time_weight=use_time_weight)
    out = []
    for fn in (theirs.get_reg_loss, ours):
        mu = mu0.clone().requires_grad_(True)
        gen = torch.Generator(device=dev).manual_seed(123)
        loss, t = fn(mu, generator=gen)
        loss.sum().backward()
        out.append((loss.detach().cpu().numpy(), mu.grad.cpu().numpy(), t.cpu().numpy()))
    (l0, g0, t0), (l1, g1, t1) = out
    assert np.array_equal(t0, t1)
    assert np.isfinite(l0).all() and np.abs(g0).max() > 0
    assert np.allclose(l0, l1, rtol=1e-4, atol=1e-7), (l0, l1)
    assert rel_l2(g1, g0) <= 1e-4


def test_reference_loop_runs_on_the_b200_operator_and_our_engine_reproduces_it(ref):
    """Drop-in check at the boundary the north-star names: the reference's InversionEngine (core/inversion.py:26-129)
    calls fwi_forward(x0_pred[:, :, 1:-1, 1:-1]) and back-propagates through it; here fwi_forward is this repo's operator."""
    from red_diffeq_b200 import FWIForward, InversionEngine, s_normalize_none, v_denormalize
    from red_diffeq_b200.utils import synthetic
    dm, mods = ref
    dev = torch.device("cuda:0")
    B, ts = 2, 3
    ctx = dict(synthetic.PDE_OPENFWI)
    ctx["nt"] = 400
    op = FWIForward(dict(ctx), dev, normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
    mu_true_n = torch.tensor(synthetic.velocity_models(B, 70, 70, seed=11), device=dev)
    with torch.no_grad():
        y = op(mu_true_n)
    mu0 = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(mu_true_n, (5, 5, 5, 5), mode="replicate"), 11, stride=1)
    mu0 = torch.nn.functional.pad(mu0, (1, 1, 1, 1), value=0.0)
    mu_true = v_denormalize(mu_true_n)
    kw = dict(ts=ts, lr=0.03, reg_lambda=0.75, regularization="diffusion")

    torch.manual_seed(8888)
    theirs = mods["inv"].InversionEngine(dm, mods["ssim"].SSIM(window_size=11), regularization="diffusion", sigma_x0=1e-4)
    mu_a, res_a = theirs.optimize(mu0, mu_true, y, op, **kw)
    torch.manual_seed(8888)
    ours = InversionEngine(dm, regularization="diffusion", sigma_x0=1e-4, cuda_graph=False)
    mu_b, res_b = ours.optimize(mu0, mu_true, y, op, **kw)

    for k in ("obs_losses", "reg_losses", "total_losses", "mae", "rmse"):
        a = np.array([[float(x) for x in r[k]] for r in res_a])
        b = np.array([[float(x) for x in r[k]] for r in res_b])
        assert np.isfinite(a).all()
        assert np.allclose(a, b, rtol=2e-3, atol=1e-6), (k, a, b)
    d = (mu_a.detach() - mu_b.detach()).abs()
    # Adam's first steps are lr * sign(g) wherever |g| >> eps: a cell whose tiny gradient changes sign between two fp32
    # evaluation orders moves by up to 2 * lr per step; everywhere else the two runs must coincide
    assert (d <= 1e-3).float().mean().item() >= 0.995, d.max().item()
    assert (mu_a.detach() - mu0[:, :, 1:-1, 1:-1]).abs().max().item() > 0.01        # the loop did move the model
    op.release_memory()


def test_graphed_iteration_with_the_regulariser_on_a_second_stream(ref):
    """Default for the diffusion regulariser on one GPU: the whole iteration -- x0 noise, U-Net under no_grad on a side
    stream beside the operator's forward kernel, fused misfit, adjoint, Adam, metrics -- captured in one CUDA graph.  It
    must optimise like the eager, serial loop: same loss trajectory within the noise the different random offsets cause
    (sigma_x0 = 1e-4, random timesteps), and a misfit that decreases."""
    from red_diffeq_b200 import FWIForward, InversionEngine, s_normalize_none, v_denormalize
    from red_diffeq_b200.utils import synthetic
    dm, _ = ref
    dev = torch.device("cuda:0")
    B, ts = 2, 8
    ctx = dict(synthetic.PDE_MARMOUSI)
    ctx["nt"] = 400
    op = FWIForward(dict(ctx), dev, normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
    mu_true_n = torch.tensor(synthetic.velocity_models(B, 70, 190, seed=12), device=dev)
    with torch.no_grad():
        y = op(mu_true_n)
    mu0 = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(mu_true_n, (5, 5, 5, 5), mode="replicate"), 11, stride=1)
    mu0 = torch.nn.functional.pad(mu0, (1, 1, 1, 1), value=0.0)
    mu_true = v_denormalize(mu_true_n)
    kw = dict(ts=ts, lr=0.03, reg_lambda=0.0, regularization="diffusion")   # lambda = 0: the data term alone decides mu
    eager = InversionEngine(dm, regularization="diffusion", cuda_graph=False, overlap_regularizer=False)
    mu_a, res_a = eager.optimize(mu0, mu_true, y, op, **kw)
    auto = InversionEngine(dm, regularization="diffusion")
    mu_b, res_b = auto.optimize(mu0, mu_true, y, op, **kw)
    assert auto.used_cuda_graph and not eager.used_cuda_graph
    a = np.array([[float(x) for x in r["obs_losses"]] for r in res_a])
    b = np.array([[float(x) for x in r["obs_losses"]] for r in res_b])
    assert (a[:, -1] < 0.7 * a[:, 0]).all() and (b[:, -1] < 0.7 * b[:, 0]).all()
    assert np.allclose(a, b, rtol=5e-2), (a, b)
    rb = np.array([[float(x) for x in r["reg_losses"]] for r in res_b])
    assert np.isfinite(rb).all() and np.abs(rb).max() > 0                    # the U-Net did run inside the graph
    # Adam's early steps are lr * sign(g): cells with tiny gradients take different paths under different x0 noise; the
    # models agree on average far better than the 8 x 0.03 a cell can move
    assert (mu_a.detach() - mu_b.detach()).abs().mean().item() < 3e-2
    op.release_memory()
