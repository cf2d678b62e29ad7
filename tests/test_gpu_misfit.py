"""GPU: the fused data misfit (FWIForward.misfit -> rdfwi_misfit_l1) against the reference's
LossCalculator.observation_loss (core/losses.py:15-40, restated below with the same torch ops) evaluated on the
operator's own seismograms: same loss, same velocity gradient."""
import numpy as np
import pytest

from conftest import Golden, rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _op(g):
    from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize
    return FWIForward(g.fresh_ctx(), "cuda:0", sample_temporal=g.sample_temporal, sample_spatial=g.sample_spatial,
                      normalize=g.normalize, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)


def _observation_loss(predicted, target, mask):
    """reference core/losses.py:27-40, verbatim semantics"""
    loss = torch.nn.L1Loss(reduction="none")(target.float(), predicted.float())
    if mask is not None:
        loss = loss * mask
        num_observed = mask.sum(dim=tuple(range(1, len(mask.shape)))).clamp(min=1.0)
        return loss.sum(dim=tuple(range(1, len(loss.shape)))) / num_observed
    return loss.mean(dim=tuple(range(1, len(loss.shape))))


def _data(g, op, masked):
    rng = np.random.default_rng(11)
    v = torch.tensor(g.v, device="cuda:0")
    with torch.no_grad():
        y = op(torch.tensor((g.v * (1 + 0.02 * rng.standard_normal(g.v.shape))).astype(np.float32), device="cuda:0"))
    mask = None
    if masked:   # missing traces: whole receivers of some shots zeroed, like utils/data_trans.py missing_trace
        m = np.ones(tuple(y.shape), dtype=np.float32)
        m[:, :, :, ::3] = 0.0
        m[0, 0] = 0.0
        mask = torch.tensor(m, device="cuda:0")
        y = y * mask
    return v, y, mask


@pytest.mark.parametrize("name", ["tiny_default", "tiny_custom", "tiny_half_receivers", "openfwi"])
@pytest.mark.parametrize("masked", [False, True])
def test_fused_misfit_equals_reference_loss_on_our_seismograms(name, masked):
    g = Golden(name)
    op = _op(g)
    v, y, mask = _data(g, op, masked)
    weights = torch.linspace(1.0, 2.0, v.shape[0], device="cuda:0")   # per-model upstream gradients (1.0 in the reference's loop)

    va = v.clone().requires_grad_(True)
    loss_a = _observation_loss(op(va), y, mask)
    (loss_a * weights).sum().backward()

    vb = v.clone().requires_grad_(True)
    loss_b, seis_b = op.misfit(vb, y, mask, return_seismograms=True)
    assert loss_b.shape == (v.shape[0],) and loss_b.dtype == torch.float32 and not seis_b.requires_grad
    (loss_b * weights).sum().backward()

    vc = v.clone().requires_grad_(True)
    loss_c = op.misfit(vc, y, mask)        # seismogram buffer turned into the sign field in place
    (loss_c * weights).sum().backward()

    with torch.no_grad():
        assert torch.equal(seis_b, op(v))
    assert torch.equal(loss_b, loss_c) and torch.equal(vb.grad, vc.grad)
    assert torch.allclose(loss_b, loss_a, rtol=2e-6, atol=0.0)
    assert rel_l2(vb.grad.cpu().numpy(), va.grad.cpu().numpy()) <= 1e-6
    with torch.no_grad():
        assert torch.equal(op.misfit(v, y, mask), loss_b)   # no history, no sign field under no_grad


def test_fused_misfit_rejects_what_it_cannot_fuse():
    from red_diffeq_b200 import FWIForward, v_denormalize
    g = Golden("tiny_default")
    op = FWIForward(g.fresh_ctx(), "cuda:0", normalize=True, v_denorm_func=v_denormalize, s_norm_func=lambda s: 2.0 * s)
    v = torch.tensor(g.v, device="cuda:0")
    with pytest.raises(ValueError):
        op.misfit(v, torch.zeros((2, 3, 130, 16), device="cuda:0"))
    op = _op(g)
    with pytest.raises(ValueError):
        op.misfit(v, torch.zeros((2, 3, 130, 15), device="cuda:0"))


def test_all_masked_model_has_zero_loss_and_gradient():
    g = Golden("tiny_default")
    op = _op(g)
    v, y, _ = _data(g, op, False)
    mask = torch.ones_like(y)
    mask[1] = 0.0                       # nothing observed for model 1: count clamps to 1 (losses.py:35), loss 0, no gradient
    vv = v.clone().requires_grad_(True)
    loss = op.misfit(vv, y, mask)
    loss.sum().backward()
    assert float(loss[1].detach()) == 0.0 and float(vv.grad[1].abs().max()) == 0.0 and float(vv.grad[0].abs().max()) > 0.0


@pytest.mark.parametrize("tag", ["masked", "plain"])
def test_misfit_kernel_reproduces_the_reference_loss_calculator(tag):
    """rdfwi_misfit_l1 called through the C ABI on recorded inputs against the outputs of the reference's own
    LossCalculator.observation_loss and its autograd (tests/golden/loop_toy.npz, written by tests/golden/make_loop_golden.py):
    per-model loss, and d(sum_b w_b loss_b)/d predicted = w_b / count_b * sign field, exact ties included."""
    import os
    from conftest import GOLDEN_DIR
    fx = np.load(os.path.join(GOLDEN_DIR, "loop_toy.npz"))
    g = Golden("tiny_default")                       # survey with (ns, nt_out, nrec) = (3, 130, 16)
    op = _op(g)
    plan = op._plan_for(g.v.shape[2], g.v.shape[3], torch.device("cuda:0"))
    seis = torch.tensor(fx["loss/pred"], device="cuda:0")
    obs = torch.tensor(fx["loss/target"], device="cuda:0")
    mask = torch.tensor(fx["loss/mask"], device="cuda:0") if tag == "masked" else None
    B = seis.shape[0]
    stats = torch.empty((B, 2), dtype=torch.float64, device="cuda:0")
    sign = torch.empty_like(seis)
    ws = torch.empty(plan.workspace_bytes(B), dtype=torch.uint8, device="cuda:0")
    plan.misfit_l1(seis.data_ptr(), obs.data_ptr(), mask.data_ptr() if mask is not None else None, B, stats.data_ptr(),
                   sign.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    loss = (stats[:, 0] / stats[:, 1].clamp(min=1.0)).float().cpu().numpy()
    assert np.allclose(loss, fx[f"loss/{tag}_loss"], rtol=2e-6, atol=0)
    w = torch.tensor([1.0, 1.7], device="cuda:0", dtype=torch.float64)
    grad = (sign * (w / stats[:, 1].clamp(min=1.0)).float().view(B, 1, 1, 1)).cpu().numpy()
    assert np.allclose(grad, fx[f"loss/{tag}_grad"], rtol=1e-6, atol=0)
    assert np.array_equal(grad == 0, fx[f"loss/{tag}_grad"] == 0)        # masked samples and exact ties carry no gradient
