"""CPU: the oracle (oracle/) against the golden vectors produced by the reference's own pde.py.

This is what pins the oracle; the GPU parity tests then compare the CUDA path with the pinned oracle
and with the same fixtures.
"""
import numpy as np

from conftest import rel_l2

# fp32 gradient: the reference's own autograd differs from its fp64 run by 4e-5..7e-5 on these cases
# (tests/golden/make_golden.py prints it); the closed-form adjoint in fp32 must stay within the
# north-star tolerance of the reference's fp32 result.
GRAD_TOL_F32 = 1e-4
GRAD_TOL_F64 = 1e-9


def _survey(oracle, g):
    return oracle.Survey(g.fresh_ctx(), g.v.shape[2], g.v.shape[3], g.sample_temporal, g.sample_spatial)


def test_forward_bit_identical(golden, oracle):
    sv = _survey(oracle, golden)
    seis = oracle.forward(sv, golden.v_phys())
    assert seis.dtype == np.float32
    assert np.array_equal(seis[:, :, ::golden.seis_stride, :], golden.seis_f32)
    np.testing.assert_allclose(seis.astype(np.float64).sum(axis=(2, 3)), golden.seis_sum, rtol=0, atol=0)
    np.testing.assert_allclose((seis.astype(np.float64) ** 2).sum(axis=(2, 3)), golden.seis_sumsq, rtol=0, atol=0)


def test_gradient_fp32(golden, oracle):
    sv = _survey(oracle, golden)
    cot = golden.cotangent((golden.v.shape[0], sv.ns, sv.nt_out, sv.nrec))
    _, grad = oracle.gradient(sv, golden.v_phys(), cot)
    scale = np.float32(1500.0) if golden.normalize else np.float32(1.0)  # d v_phys / d v_norm
    assert rel_l2(grad * scale, golden.grad_f32) <= GRAD_TOL_F32


def test_gradient_fp64(golden, oracle):
    sv = _survey(oracle, golden)
    cot = golden.cotangent((golden.v.shape[0], sv.ns, sv.nt_out, sv.nrec)).astype(np.float64)
    _, grad = oracle.gradient(sv, golden.v_phys(np.float64), cot, dtype=np.float64)
    scale = 1500.0 if golden.normalize else 1.0
    assert rel_l2(grad * scale, golden.grad_f64) <= GRAD_TOL_F64
