"""GPU: the CUDA path (through the C ABI, via the FWIForward operator) against the golden vectors of the
reference and against the pinned CPU oracle.

Tolerances (north-star): seismograms relative L2 <= 1e-5 (we additionally require bit-identity, which the
kernel is designed for), velocity gradient relative L2 <= 1e-4 with the cotangent held fixed.
"""
import numpy as np
import pytest

from conftest import Golden, rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SEIS_TOL = 1e-5
GRAD_TOL = 1e-4


def _op(g, **kw):
    from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize
    return FWIForward(g.fresh_ctx(), "cuda:0", sample_temporal=g.sample_temporal, sample_spatial=g.sample_spatial,
                      normalize=g.normalize, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none, **kw)


def _run(op, v_np, cot_np=None):
    v = torch.tensor(v_np, device="cuda:0", requires_grad=cot_np is not None)
    seis = op(v)
    grad = None
    if cot_np is not None:
        (seis * torch.tensor(cot_np, device="cuda:0")).sum().backward()
        grad = v.grad.cpu().numpy()
    return seis.detach().cpu().numpy(), grad


def test_native_library_loaded():
    from red_diffeq_b200 import _cabi
    assert _cabi.load().rdfwi_version() >= 100
    with open("/proc/self/maps") as f:
        assert "librdfwi.so" in f.read()


def test_seismograms_match_reference(golden):
    seis, _ = _run(_op(golden), golden.v)
    ref = golden.seis_f32
    got = seis[:, :, ::golden.seis_stride, :]
    assert rel_l2(got, ref) <= SEIS_TOL
    assert np.array_equal(got, ref), f"not bit-identical: max abs diff {np.abs(got - ref).max()}"
    np.testing.assert_array_equal(seis.astype(np.float64).sum(axis=(2, 3)), golden.seis_sum)


def test_gradient_matches_reference(golden):
    op = _op(golden)
    shape = (golden.v.shape[0], len(op.ctx["sx"]), -(-golden.ctx["nt"] // golden.sample_temporal), len(op.ctx["gx"]))
    cot = golden.cotangent(shape)
    _, grad = _run(op, golden.v, cot)
    assert rel_l2(grad, golden.grad_f32) <= GRAD_TOL
    # informational bound against the reference run in fp64 (the reference's own fp32 is 4e-5..7e-5 away)
    assert rel_l2(grad, golden.grad_f64) <= 2e-4


@pytest.mark.parametrize("name", ["tiny_default", "tiny_custom", "tiny_half_receivers"])
@pytest.mark.parametrize("rows", [1, 2, 4, 8])
@pytest.mark.parametrize("chunk", [0, 1])
@pytest.mark.parametrize("engine", ["per-level", "per-level-fused", "cluster-resident", "cluster-split", "cluster+per-level-fused"])
def test_kernel_variants_agree_with_oracle(name, rows, chunk, engine, oracle):
    if not engine.startswith("per-level") and (rows != 1 or chunk != 0):
        pytest.skip("rows/chunk only affect the per-level engine")
    g = Golden(name)
    op = _op(g)
    op.set_option("engine", 1 if engine.startswith("per-level") else 2)
    op.set_option("adj_mode", 1 if engine.endswith("fused") else 0)   # split (default) or fused adjoint
    if engine == "per-level":
        op.set_option("u_chunk_shots", 3)                                 # several chunks even on the tiny cases
    if engine == "cluster-split":
        op.set_option("imaging", 1)                                       # adjoint-field history + streaming imaging kernel
        op.set_option("u_chunk_shots", 2)                                 # several chunks even on the tiny cases
    op.set_option("rows_per_thread", rows)
    op.set_option("adj_rows_per_thread", 1 if rows in (1, 8) else 2)
    op.set_option("chunk_models", chunk)
    sv = oracle.Survey(g.fresh_ctx(), g.v.shape[2], g.v.shape[3], g.sample_temporal, g.sample_spatial)
    cot = g.cotangent((g.v.shape[0], sv.ns, sv.nt_out, sv.nrec))
    seis, grad = _run(op, g.v, cot)
    seis_o, grad_o = oracle.gradient(sv, g.v_phys(), cot)
    assert np.array_equal(seis, seis_o)
    scale = 1500.0 if g.normalize else 1.0
    assert rel_l2(grad, grad_o * scale) <= GRAD_TOL


@pytest.mark.parametrize("name", ["tiny_default", "tiny_custom", "tiny_half_receivers"])
@pytest.mark.parametrize("segment", [3, 7, 16, 200])
def test_checkpointed_history_matches_full_history(name, segment):
    """Recompute-from-checkpoints backward == full-history backward (same kernels, same order => bit-identical)."""
    g = Golden(name)
    full = _op(g)
    full.set_option("engine", 1)
    full.set_option("adj_mode", 1)   # the fused per-level adjoint: the kernel the checkpointed path runs
    full.set_history_segment(0)
    ck = _op(g)
    ck.set_option("engine", 1)     # (segment >= nt on the cluster engine is the recompute tier, tested below)
    ck.set_history_segment(segment)
    shape = (g.v.shape[0], len(full.ctx["sx"]), -(-g.ctx["nt"] // g.sample_temporal), len(full.ctx["gx"]))
    cot = g.cotangent(shape)
    s0, g0 = _run(full, g.v, cot)
    s1, g1 = _run(ck, g.v, cot)
    assert np.array_equal(s0, s1)
    assert np.array_equal(g0, g1)
    assert rel_l2(g1, g.grad_f32) <= GRAD_TOL


@pytest.mark.parametrize("name", ["tiny_default", "tiny_custom", "tiny_half_receivers", "openfwi"])
@pytest.mark.parametrize("chunk", [0, 2])
@pytest.mark.parametrize("imaging", [0, 1])
def test_recomputed_history_matches_full_history(name, chunk, imaging):
    """No history kept (segment = nt): the backward pass recomputes the forward field chunk by chunk on the cluster
    engine and runs the adjoint (resident imaging, or the split adjoint) on it -- same kernels, same order =>
    bit-identical to the full-history run."""
    g = Golden(name)
    full = _op(g)
    full.set_option("engine", 2)
    full.set_option("imaging", imaging)
    full.set_history_segment(0)
    rc = _op(g)
    rc.set_option("engine", 2)
    rc.set_option("imaging", imaging)
    if chunk:
        rc.set_option("u_chunk_shots", chunk)
    rc.set_history_segment(g.ctx["nt"])
    shape = (g.v.shape[0], len(full.ctx["sx"]), -(-g.ctx["nt"] // g.sample_temporal), len(full.ctx["gx"]))
    cot = g.cotangent(shape)
    s0, g0 = _run(full, g.v, cot)
    s1, g1 = _run(rc, g.v, cot)
    plan = rc._plan_for(g.v.shape[2], g.v.shape[3], torch.device("cuda:0"))
    assert plan.get("adj_split") == (2 if imaging == 1 else 5) and plan.history_bytes(g.v.shape[0], g.ctx["nt"]) == 0
    assert np.array_equal(s0, s1)
    assert np.array_equal(g0, g1)
    assert rel_l2(g1, g.grad_f32) <= GRAD_TOL


@pytest.mark.parametrize("name,rows,csize", [("tiny_default", 4, 0), ("tiny_default", 4, 3), ("tiny_custom", 7, 2),
                                             ("tiny_half_receivers", 4, 5), ("tiny_half_receivers", 7, 0),
                                             ("openfwi", 7, 0), ("openfwi", 4, 0), ("marmousi", 7, 0), ("marmousi", 13, 8),
                                             ("tiny_default", 5, 2), ("tiny_custom", 5, 0), ("openfwi", 5, 0), ("openfwi", 5, 16),
                                             ("marmousi", 5, 16)])
def test_wide_cluster_configurations(name, rows, csize):
    """Few-shot configurations of the cluster-resident time loop (a shot spread over more CTAs, 7, 5 or 4 rows marched per
    thread instead of 13): same arithmetic per cell => seismograms bit-identical to the reference fixtures and gradients
    bit-identical to the throughput configuration (13 rows, smallest cluster)."""
    g = Golden(name)
    base = _op(g)
    base.set_option("engine", 2)
    base.set_option("cluster_rows", 13)
    wide = _op(g)
    wide.set_option("engine", 2)
    wide.set_option("cluster_rows", rows)
    if csize:
        wide.set_option("cluster_size", csize)
    shape = (g.v.shape[0], len(base.ctx["sx"]), -(-g.ctx["nt"] // g.sample_temporal), len(base.ctx["gx"]))
    cot = g.cotangent(shape)
    s0, g0 = _run(base, g.v, cot)
    s1, g1 = _run(wide, g.v, cot)
    plan = wide._plan_for(g.v.shape[2], g.v.shape[3], torch.device("cuda:0"))
    assert plan.get("cluster_rows_last") == rows and (csize == 0 or plan.get("cluster_size_last") == csize)
    assert np.array_equal(s1[:, :, ::g.seis_stride, :], g.seis_f32)
    assert np.array_equal(s0, s1)
    assert np.array_equal(g0, g1)
    assert rel_l2(g1, g.grad_f32) <= GRAD_TOL


def test_few_shots_pick_a_wide_cluster():
    """One OpenFWI model (5 shots) would occupy 20 SMs with the throughput configuration: the automatic choice is a wider
    cluster with fewer rows per thread; a batch that fills the GPU keeps 13 rows per thread."""
    g = Golden("openfwi")
    op = _op(g)
    s, _ = _run(op, g.v)
    plan = op._plan_for(70, 70, torch.device("cuda:0"))
    assert np.array_equal(s[:, :, ::g.seis_stride, :], g.seis_f32)
    assert plan.get("cluster_rows_last") in (4, 5, 7) and plan.get("cluster_size_last") > plan.get("cluster_size_used")
    many = np.repeat(g.v, 16, axis=0)
    s16, _ = _run(op, many)
    assert plan.get("cluster_rows_last") == 13 and plan.get("cluster_size_last") == plan.get("cluster_size_used")
    assert np.array_equal(s16[3:4], s)


@pytest.mark.parametrize("engine", ["per-level", "per-level-fused", "per-level-checkpointed", "cluster-resident", "cluster-split", "cluster+per-level-fused"])
def test_many_shots_per_model(engine, oracle):
    """More shots than any golden case (ns = 11: the per-level adjoint deals them over several grid.z slices, each with its
    own imaging plane; the cluster engines run more shots than fit one chunk)."""
    from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize
    from red_diffeq_b200.utils import synthetic
    nz, nx, B = 20, 28, 2
    ctx = dict(n_grid=nx, nt=130, dx=10.0, dt=0.001, nbc=12, f=25.0, sz=10, gz=10, ng=nx, ns=11)
    sv = oracle.Survey(dict(ctx), nz, nx)
    vn = synthetic.velocity_models(B, nz, nx, seed=31)
    cot = synthetic.cotangent((B, sv.ns, sv.nt_out, sv.nrec), seed=32)
    op = FWIForward(dict(ctx), "cuda:0", normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
    op.set_option("engine", 1 if engine.startswith("per-level") else 2)
    op.set_option("adj_mode", 1 if engine.endswith("fused") else 0)
    if engine == "cluster-split":
        op.set_option("imaging", 1)
    if engine in ("cluster-split", "per-level"):
        op.set_option("u_chunk_shots", 7)
    if engine == "per-level-checkpointed":
        op.set_history_segment(16)
    seis, grad = _run(op, vn, cot)
    v_phys = (vn + np.float32(1)) / np.float32(2) * np.float32(3000) + np.float32(1500)
    seis_o, grad_o = oracle.gradient(sv, v_phys, cot)
    assert np.array_equal(seis, seis_o)
    assert rel_l2(grad, grad_o * 1500.0) <= GRAD_TOL


def test_coefficients_bit_identical(oracle):
    from red_diffeq_b200 import _cabi
    g = Golden("tiny_custom")
    op = _op(g)
    nz, nx = g.v.shape[2:]
    plan = op._plan_for(nz, nx, torch.device("cuda:0"))
    B = g.v.shape[0]
    v = torch.tensor(g.v_phys(), device="cuda:0").contiguous()
    pitch, nzp, nbc, ns = plan.get("pitch"), plan.get("nzp"), g.ctx["nbc"], plan.ns
    alpha = torch.empty((B, nzp, pitch), device="cuda:0")
    kap = torch.empty((B, nbc + 1), device="cuda:0")
    velmin = torch.empty(B, device="cuda:0")
    argmin = torch.empty(B, dtype=torch.int32, device="cuda:0")
    beta = torch.empty((B, ns), device="cuda:0")
    ws = torch.empty(plan.workspace_bytes(B), dtype=torch.uint8, device="cuda:0")
    plan.coefficients(v.data_ptr(), B, alpha.data_ptr(), kap.data_ptr(), velmin.data_ptr(), argmin.data_ptr(),
                      beta.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    sv = oracle.Survey(g.fresh_ctx(), nz, nx, g.sample_temporal, g.sample_spatial)
    nxp = nx + 2 * nbc
    for b in range(B):
        planes, vmin, amin = oracle.coefficient_planes(sv, g.v_phys()[b:b + 1])
        a = alpha[b].cpu().numpy()
        assert np.array_equal(a[:, :nxp], planes[0])
        assert np.array_equal(a[:, nxp:], planes[0][:, :pitch - nxp])  # periodic image columns
        assert velmin[b].item() == vmin and argmin[b].item() == amin
        k = kap[b].cpu().numpy()
        # kappa plane: column nbc+1 (interior column) top rows walk the profile from the outside in
        col = planes[1][:nbc, nbc + 1]
        assert np.array_equal(k[:nbc][::-1], col)
        assert k[nbc] == 0.0
        src_beta = planes[4][sv.isz, sv.isx % nxp]
        assert np.array_equal(beta[b].cpu().numpy(), src_beta)


def test_no_grad_and_noncontiguous_input():
    g = Golden("tiny_half_receivers")
    op = _op(g)
    B, _, nz, nx = g.v.shape
    leaf = torch.zeros((B, 1, nz + 2, nx + 2), device="cuda:0")
    leaf[:, :, 1:-1, 1:-1] = torch.tensor(g.v, device="cuda:0")
    leaf.requires_grad_(True)
    view = leaf[:, :, 1:-1, 1:-1]  # what core/inversion.py:78 passes
    assert not view.is_contiguous()
    seis = op(view)
    assert np.array_equal(seis.detach().cpu().numpy()[:, :, ::g.seis_stride], g.seis_f32)
    cot = g.cotangent(tuple(seis.shape))
    (seis * torch.tensor(cot, device="cuda:0")).sum().backward()
    grad = leaf.grad.cpu().numpy()
    assert rel_l2(grad[:, :, 1:-1, 1:-1], g.grad_f32) <= GRAD_TOL
    assert np.all(grad[:, :, 0, :] == 0) and np.all(grad[:, :, :, 0] == 0)
    with torch.no_grad():
        seis2 = op(view)
    assert not seis2.requires_grad
    assert torch.equal(seis2, seis.detach())


def test_batch_independence_and_determinism():
    """Size-independent properties: every model of a batch equals its solo run; repeated runs are bit-identical."""
    g = Golden("tiny_default")
    op = _op(g)
    v = np.concatenate([g.v, g.v[::-1].copy(), g.v * 0.9 + 300.0], axis=0).astype(np.float32)
    cot = np.random.default_rng(0).standard_normal((v.shape[0], 3, 130, 16)).astype(np.float32)
    s_all, g_all = _run(op, v, cot)
    s_again, g_again = _run(op, v, cot)
    assert np.array_equal(s_all, s_again) and np.array_equal(g_all, g_again)
    for b in range(v.shape[0]):
        s_b, g_b = _run(op, v[b:b + 1], cot[b:b + 1])
        assert np.array_equal(s_b, s_all[b:b + 1])
        assert np.array_equal(g_b, g_all[b:b + 1])


def test_adjoint_is_linear_in_cotangent():
    g = Golden("tiny_default")
    op = _op(g)
    rng = np.random.default_rng(1)
    c1 = rng.standard_normal((2, 3, 130, 16)).astype(np.float32)
    c2 = rng.standard_normal((2, 3, 130, 16)).astype(np.float32)
    _, g1 = _run(op, g.v, c1)
    _, g2 = _run(op, g.v, c2)
    _, g12 = _run(op, g.v, (c1 + 2 * c2).astype(np.float32))
    assert rel_l2(g12, g1 + 2 * g2) <= 1e-5


def test_gradient_against_finite_differences():
    """Directional derivative of sum(seis*cot) vs the adjoint gradient (fp32 forward, so a loose tolerance)."""
    g = Golden("tiny_default")
    op = _op(g)
    rng = np.random.default_rng(2)
    cot = rng.standard_normal((2, 3, 130, 16)).astype(np.float32)
    _, grad = _run(op, g.v, cot)
    dv = rng.standard_normal(g.v.shape).astype(np.float32)
    eps = 2.0  # m/s
    sp, _ = _run(op, (g.v + eps * dv).astype(np.float32))
    sm, _ = _run(op, (g.v - eps * dv).astype(np.float32))
    fd = ((sp.astype(np.float64) - sm.astype(np.float64)) * cot).sum() / (2 * eps)
    an = (grad.astype(np.float64) * dv).sum()
    assert abs(fd - an) <= 2e-2 * abs(an)


def test_errors():
    from red_diffeq_b200 import FWIForward
    g = Golden("tiny_default")
    with pytest.raises(RuntimeError):
        FWIForward(g.fresh_ctx(), "cpu", normalize=False)
    ctx = g.fresh_ctx()
    ctx["nt"] = 20  # shorter than the wavelet: ValueError like numpy's broadcast error in the reference
    op = FWIForward(ctx, "cuda:0", normalize=False)
    with pytest.raises(ValueError):
        op(torch.tensor(g.v, device="cuda:0"))
    ctx = g.fresh_ctx()
    ctx["sx"] = [500]
    op = FWIForward(ctx, "cuda:0", normalize=False)
    with pytest.raises(IndexError):
        op(torch.tensor(g.v, device="cuda:0"))


@pytest.mark.parametrize("nx,nbc", [(15, 9), (17, 9), (21, 10), (16, 8)])
@pytest.mark.parametrize("engine", ["per-level", "per-level-fused", "cluster-resident", "cluster-split", "cluster+per-level-fused"])
def test_odd_widths_against_oracle(nx, nbc, engine, oracle):
    """Padded widths with nxp % 4 in {1, 3, 0, ...}: the periodic image columns of the pitched layout (1..3 of them)
    must reproduce torch.roll's wrap-around bit for bit; no reference fixture has such a width, so the pinned oracle checks."""
    from red_diffeq_b200 import FWIForward
    ctx = dict(n_grid=nx, nt=125, dx=10.0, dt=0.001, nbc=nbc, f=25.0, sz=20, gz=10, ng=nx, ns=2)
    nz = 11
    rng = np.random.default_rng(nx * 100 + nbc)
    v = (1500 + 3000 * rng.random((2, 1, nz, nx))).astype(np.float32)
    op = FWIForward(dict(ctx), "cuda:0", normalize=False)
    op.set_option("engine", 1 if engine.startswith("per-level") else 2)
    op.set_option("adj_mode", 1 if engine.endswith("fused") else 0)
    sv = oracle.Survey(dict(ctx), nz, nx)
    cot = rng.standard_normal((2, sv.ns, sv.nt_out, sv.nrec)).astype(np.float32)
    seis, grad = _run(op, v, cot)
    seis_o, grad_o = oracle.gradient(sv, v, cot)
    assert (sv.nxp % 4) == (nx + 2 * nbc) % 4
    assert np.array_equal(seis, seis_o)
    assert rel_l2(grad, grad_o) <= GRAD_TOL


def test_tiled_engine_on_a_grid_wider_than_a_tile(oracle):
    """Per-level engine on a grid of several 128-column x 32-row tiles (padded 172 x 330, width not a multiple of 4 or
    128; the engine is forced, this size would fit a cluster): bit-identical seismograms, gradient against the pinned oracle."""
    from red_diffeq_b200 import FWIForward
    nz, nx, nbc = 132, 290, 20
    ctx = dict(n_grid=nx, nt=140, dx=10.0, dt=0.001, nbc=nbc, f=25.0, sz=15, gz=12, ng=nx, ns=3)
    rng = np.random.default_rng(77)
    v = (1500 + 3000 * rng.random((1, 1, nz, nx))).astype(np.float32)
    sv = oracle.Survey(dict(ctx), nz, nx)
    cot = rng.standard_normal((1, sv.ns, sv.nt_out, sv.nrec)).astype(np.float32)
    seis_o, grad_o = oracle.gradient(sv, v, cot)
    for rows in (4, 8):
        op = FWIForward(dict(ctx), "cuda:0", normalize=False)
        op.set_option("engine", 1)
        op.set_option("rows_per_thread", rows)
        op.set_option("u_chunk_shots", 2)
        seis, grad = _run(op, v, cot)
        assert np.array_equal(seis, seis_o)
        assert rel_l2(grad, grad_o) <= GRAD_TOL


def test_full_size_batch_properties():
    """BASELINE configs[1] at full size (64 OpenFWI models x 5 shots x 1000 levels, 124 GB of wavefield history):
    size-independent properties -- every model of the batch is bit-identical to its solo run (seismograms and
    gradient), and a repeated evaluation is bit-identical (no atomics anywhere)."""
    from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize
    from red_diffeq_b200.utils import synthetic
    free, _ = torch.cuda.mem_get_info()
    if free < 150e9:
        pytest.skip("needs ~130 GB of free HBM")
    ctx = dict(synthetic.PDE_OPENFWI)
    op = FWIForward(dict(ctx), "cuda:0", normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
    B = 64
    g = Golden("openfwi")
    vn = synthetic.velocity_models(B, 70, 70)
    vn[0] = g.v[0]                              # model 0 = the reference fixture's model
    cot = synthetic.cotangent((B, 5, 1000, 70), seed=3)
    v = torch.tensor(vn, device="cuda:0", requires_grad=True)
    c = torch.tensor(cot, device="cuda:0")
    seis = op(v)
    seis.backward(c)
    g1 = v.grad.clone()
    v.grad = None
    seis2 = op(v)
    seis2.backward(c)
    assert torch.equal(seis, seis2) and torch.equal(g1, v.grad)
    assert np.array_equal(seis[:1].detach().cpu().numpy(), g.seis_f32)   # bit-identical to the reference inside a batch of 64
    solo = FWIForward(dict(ctx), "cuda:0", normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
    for b in (0, 31, 63):
        vb = torch.tensor(vn[b:b + 1], device="cuda:0", requires_grad=True)
        sb = solo(vb)
        sb.backward(c[b:b + 1])
        assert torch.equal(sb, seis[b:b + 1])
        assert torch.equal(vb.grad, g1[b:b + 1])
    op.release_memory()
    solo.release_memory()


def test_c_abi_from_plain_c(tmp_path):
    """examples/c_abi_example.c: the library used from plain C (no Python, no torch) -- forward + adjoint run and return a
    finite, non-zero gradient."""
    import os
    import re
    import shutil
    import subprocess
    from conftest import ROOT
    from red_diffeq_b200 import _cabi
    cudart = "/usr/local/cuda/lib64"
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(cudart, "libcudart.so")):
        pytest.skip("needs gcc and the CUDA runtime")
    exe = str(tmp_path / "rdfwi_example")
    libdir = os.path.dirname(_cabi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_example.c"),
                           "-o", exe, "-L", libdir, "-lrdfwi", "-L", cudart, "-lcudart", "-lm",
                           "-Wl,-rpath," + libdir, "-Wl,-rpath," + cudart])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    m = re.search(r"= ([0-9.e+-]+), kernel launches of the adjoint pass: (\d+)", out.stdout)
    assert m and np.isfinite(float(m.group(1))) and float(m.group(1)) > 0 and int(m.group(2)) > 3, out.stdout
