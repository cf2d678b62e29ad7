"""Host-side acquisition geometry and source wavelet of the FWI forward operator.

Mirrors what the reference's FWIForward does on the host in numpy before the time loop:
  - default source / receiver coordinates written back into the caller's ctx (solvers/pde.py:16-23)
  - zero-phase Ricker wavelet, float64, zero-padded to nt                     (solvers/pde.py:26-36)
  - metres -> padded grid indices with round-half-to-even                     (solvers/pde.py:54-59)
"""
import numpy as np

REQUIRED_KEYS = ("n_grid", "nt", "dx", "dt", "nbc", "f", "sz", "gz", "ng", "ns")


def complete_ctx(ctx, sample_spatial=1.0):
    """Fill ctx['sx'] / ctx['gx'] (metres) in place, exactly as the reference constructor does.

    User-supplied sx / gx are in grid units and get scaled by dx; otherwise sources and receivers are
    spread evenly over [0, n_grid-1] cells.
    """
    dx = ctx["dx"]
    last = ctx["n_grid"] - 1
    if "sx" in ctx.keys():
        ctx["sx"] = np.array(ctx["sx"]) * dx
    else:
        ctx["sx"] = np.linspace(0, last, num=ctx["ns"]) * dx
    if "gx" in ctx.keys():
        ctx["gx"] = np.array(ctx["gx"]) * dx
    else:
        ctx["gx"] = np.linspace(0, last, num=int(sample_spatial * ctx["ng"])) * dx
    return ctx


def ricker(f, dt, nt):
    """Ricker wavelet with peak frequency f sampled at dt, as float64 of length nt.

    Tap count 2*floor(2.2/(f*dt)/2)+1, peak in the middle tap.  Like the reference (numpy slice
    assignment, solvers/pde.py:35) a record shorter than the wavelet is a ValueError.
    """
    half = np.floor((2.2 / f / dt) / 2)
    n_taps = 2 * half + 1
    phase = (np.floor(n_taps / 2) - np.arange(n_taps)) * f * dt * np.pi
    sq = phase ** 2
    taps = (1 - sq * 2) * np.exp(-sq)
    if taps.shape[0] > nt:
        raise ValueError(f"could not broadcast input array from shape ({taps.shape[0]},) into shape ({nt},)")
    w = np.zeros(nt)
    w[:taps.shape[0]] = taps
    return w


def grid_indices(sx, sz, gx, gz, dx, nbc):
    """(isx, isz, igx, igz) on the padded grid; np.rint == np.around: halves go to the even cell."""
    isx = (np.rint(np.asarray(sx, dtype=np.float64) / dx) + nbc).astype("int")
    igx = (np.rint(np.asarray(gx, dtype=np.float64) / dx) + nbc).astype("int")
    isz = int(np.rint(sz / dx) + nbc)
    igz = int(np.rint(gz / dx) + nbc)
    return isx, isz, igx, igz


def wrap_indices(idx, n, what):
    """Python/torch-style negative indexing; out-of-range is an IndexError like the reference's tensor indexing."""
    idx = np.asarray(idx)
    if np.any(idx < -n) or np.any(idx >= n):
        raise IndexError(f"{what} index out of range for padded grid of size {n}")
    return np.where(idx < 0, idx + n, idx)
