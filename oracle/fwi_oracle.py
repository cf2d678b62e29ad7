"""numpy/ctypes front end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Restates the host-side pieces of the reference's red_diffeq/solvers/pde.py in numpy
(wavelet :26-36, default geometry :16-23, index mapping :54-59) and drives the C restatement
of the recurrence and its adjoint (oracle/fwi_oracle.c).  Pinned against the reference by
tests/golden/ (see tests/golden/make_golden.py).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Geom(ctypes.Structure):
    _fields_ = [
        ("B", ctypes.c_int), ("nz", ctypes.c_int), ("nx", ctypes.c_int), ("nbc", ctypes.c_int),
        ("nzp", ctypes.c_int), ("nxp", ctypes.c_int), ("ns", ctypes.c_int), ("nrec", ctypes.c_int),
        ("nt", ctypes.c_int), ("st", ctypes.c_int), ("isz", ctypes.c_int), ("igz", ctypes.c_int),
        ("dx", ctypes.c_double), ("dt", ctypes.c_double),
        ("isx", ctypes.c_void_p), ("igx", ctypes.c_void_p), ("wavelet", ctypes.c_void_p),
    ]


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libfwi_oracle.so")
        if not os.path.exists(path):
            from . import build_oracle
            build_oracle.build()
        _LIB = ctypes.CDLL(path)
        for name in ("fwi_oracle_forward_f32", "fwi_oracle_forward_f64", "fwi_oracle_gradient_f32",
                     "fwi_oracle_gradient_f64", "fwi_oracle_coeffs_f32"):
            getattr(_LIB, name).restype = ctypes.c_int
    return _LIB


def threads() -> int:
    return int(_lib().fwi_oracle_threads())


def set_threads(n: int) -> None:
    """OpenMP threads used by the following calls (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    _lib().fwi_oracle_set_threads(int(n))


def ricker_wavelet(f, dt, nt):
    """Zero-phase Ricker, float64, zero-padded to nt (solvers/pde.py:26-36).

    Raises ValueError when the wavelet is longer than the record, as numpy does there (:35).
    """
    n_taps = 2.0 * np.floor(2.2 / f / dt / 2.0) + 1.0
    centre = np.floor(n_taps / 2.0)
    arg = (centre - np.arange(n_taps)) * f * dt * np.pi
    arg2 = arg ** 2
    taps = (1.0 - arg2 * 2.0) * np.exp(-arg2)
    if len(taps) > nt:
        raise ValueError(f"could not broadcast input array from shape ({len(taps)},) into shape ({nt},)")
    out = np.zeros(nt)
    out[:len(taps)] = taps
    return out


class Survey:
    """Acquisition geometry + discretisation of one operator instance.

    ctx keys as in the reference's `pde:` YAML blocks: n_grid, nt, dx, dt, nbc, f, sz, gz, ng, ns
    and optional sx / gx in grid units (solvers/pde.py:16-23).
    """

    def __init__(self, ctx, nz, nx, sample_temporal=1, sample_spatial=1.0):
        dx = ctx["dx"]
        if "sx" in ctx:
            sx = np.array(ctx["sx"]) * dx
        else:
            sx = np.linspace(0, ctx["n_grid"] - 1, num=ctx["ns"]) * dx
        if "gx" in ctx:
            gx = np.array(ctx["gx"]) * dx
        else:
            gx = np.linspace(0, ctx["n_grid"] - 1, num=int(sample_spatial * ctx["ng"])) * dx
        nbc = int(ctx["nbc"])
        # metres -> padded grid index, round-half-to-even (solvers/pde.py:54-59)
        self.isx = np.ascontiguousarray((np.around(sx / dx) + nbc).astype("int").astype(np.int32))
        self.igx = np.ascontiguousarray((np.around(gx / dx) + nbc).astype("int").astype(np.int32))
        self.isz = int(np.around(ctx["sz"] / dx) + nbc)
        self.igz = int(np.around(ctx["gz"] / dx) + nbc)
        self.nz, self.nx, self.nbc = int(nz), int(nx), nbc
        self.nzp, self.nxp = self.nz + 2 * nbc, self.nx + 2 * nbc
        self.ns, self.nrec = len(self.isx), len(self.igx)
        self.nt, self.st = int(ctx["nt"]), int(sample_temporal)
        self.nt_out = (self.nt + self.st - 1) // self.st
        self.dx, self.dt = float(dx), float(ctx["dt"])
        self.wavelet = np.ascontiguousarray(ricker_wavelet(ctx["f"], ctx["dt"], self.nt), dtype=np.float64)
        for name, idx, hi in (("source column", self.isx, self.nxp), ("receiver column", self.igx, self.nxp)):
            if idx.min() < -hi or idx.max() >= hi:
                raise IndexError(f"{name} out of the padded grid")
        for name, idx in (("source row", self.isz), ("receiver row", self.igz)):
            if idx < -self.nzp or idx >= self.nzp:
                raise IndexError(f"{name} out of the padded grid")

    def c_geom(self, B):
        g = _Geom()
        g.B, g.nz, g.nx, g.nbc, g.nzp, g.nxp = B, self.nz, self.nx, self.nbc, self.nzp, self.nxp
        g.ns, g.nrec, g.nt, g.st = self.ns, self.nrec, self.nt, self.st
        # negative python-style indices are wrapped like torch indexing would
        g.isz, g.igz = self.isz % self.nzp, self.igz % self.nzp
        self._isx_w = np.ascontiguousarray(self.isx % self.nxp, dtype=np.int32)
        self._igx_w = np.ascontiguousarray(self.igx % self.nxp, dtype=np.int32)
        g.dx, g.dt = self.dx, self.dt
        g.isx = self._isx_w.ctypes.data
        g.igx = self._igx_w.ctypes.data
        g.wavelet = self.wavelet.ctypes.data
        return g

    def pairs(self, B):
        """fwd+adjoint cell-update pairs of one gradient evaluation."""
        return B * self.ns * self.nzp * self.nxp * self.nt


def _prep(v, dtype):
    v = np.ascontiguousarray(v, dtype=dtype)
    if v.ndim == 4:
        assert v.shape[1] == 1
        v = v[:, 0]
    assert v.ndim == 3
    return v


def forward(survey: Survey, v_phys, dtype=np.float32, return_history=False):
    """v_phys (B,1,nz,nx) or (B,nz,nx) in m/s -> seismograms (B, ns, nt_out, nrec)."""
    v = _prep(v_phys, dtype)
    B = v.shape[0]
    g = survey.c_geom(B)
    seis = np.empty((B, survey.ns, survey.nt_out, survey.nrec), dtype=dtype)
    hist = None
    hist_p = None
    if return_history:
        hist = np.empty((B * survey.ns, survey.nt, survey.nzp + 4, survey.nxp + 4), dtype=dtype)
        hist_p = ctypes.c_void_p(hist.ctypes.data)
    fn = _lib().fwi_oracle_forward_f32 if dtype == np.float32 else _lib().fwi_oracle_forward_f64
    rc = fn(ctypes.byref(g), ctypes.c_void_p(v.ctypes.data), ctypes.c_void_p(seis.ctypes.data), hist_p)
    if rc != 0:
        raise MemoryError("oracle forward failed")
    if return_history:
        return seis, hist[:, :, 2:-2, 2:-2]
    return seis


def gradient(survey: Survey, v_phys, cotangent, dtype=np.float32):
    """Returns (seismograms, d sum(seis*cotangent) / d v_phys) with shapes (B,ns,nt_out,nrec), (B,1,nz,nx)."""
    v = _prep(v_phys, dtype)
    B = v.shape[0]
    g = survey.c_geom(B)
    cot = np.ascontiguousarray(cotangent, dtype=dtype)
    assert cot.shape == (B, survey.ns, survey.nt_out, survey.nrec)
    seis = np.empty_like(cot)
    grad = np.empty((B, 1, survey.nz, survey.nx), dtype=dtype)
    fn = _lib().fwi_oracle_gradient_f32 if dtype == np.float32 else _lib().fwi_oracle_gradient_f64
    rc = fn(ctypes.byref(g), ctypes.c_void_p(v.ctypes.data), ctypes.c_void_p(cot.ctypes.data),
            ctypes.c_void_p(seis.ctypes.data), ctypes.c_void_p(grad.ctypes.data))
    if rc != 0:
        raise MemoryError("oracle gradient failed")
    return seis, grad


def coefficient_planes(survey: Survey, v_phys_one):
    """alpha, kappa, temp1, temp2, beta_dt planes (nzp, nxp) + velmin + argmin of ONE model, fp32."""
    v = _prep(v_phys_one, np.float32)
    assert v.shape[0] == 1
    g = survey.c_geom(1)
    planes = np.empty((5, survey.nzp, survey.nxp), dtype=np.float32)
    velmin = ctypes.c_float()
    argmin = ctypes.c_int()
    _lib().fwi_oracle_coeffs_f32(ctypes.byref(g), ctypes.c_void_p(v.ctypes.data), ctypes.c_void_p(planes.ctypes.data),
                                 ctypes.byref(velmin), ctypes.byref(argmin))
    return planes, velmin.value, argmin.value
