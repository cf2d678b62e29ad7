"""ctypes binding of the C ABI in include/rdfwi.h (librdfwi.so, built in-tree for sm_100a).

There is no CPU fallback and no alternative backend: if the shared library is missing or a call
fails, an exception is raised.
"""
import ctypes
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
_CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.environ.get("RDFWI_LIB") or os.path.join(_PKG, "librdfwi.so")   # RDFWI_LIB: A/B builds of the same sources
SOURCES = ["rdfwi_api.cu", "kernels_prologue.cu", "kernels_step.cu", "kernels_tile.cu", "kernels_cluster.cu", "kernels_imaging.cu", "kernels_epilogue.cu", "kernels_misfit.cu"]
HEADERS = [os.path.join(_CSRC, "rdfwi_common.cuh"), os.path.join(_CSRC, "cluster_ptx.cuh"), os.path.join(_ROOT, "include", "rdfwi.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--threads", "0",
              "-Xcompiler", "-fPIC", "-shared"]

# every symbol include/rdfwi.h declares
EXPORTS = ["rdfwi_version", "rdfwi_last_error", "rdfwi_plan_create", "rdfwi_plan_destroy", "rdfwi_plan_set",
           "rdfwi_plan_get", "rdfwi_level_floats", "rdfwi_workspace_bytes", "rdfwi_history_bytes", "rdfwi_forward",
           "rdfwi_backward", "rdfwi_coefficients", "rdfwi_misfit_l1", "rdfwi_last_launch_count"]


class RdfwiError(RuntimeError):
    pass


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force=False, verbose=False, defines=(), out=None):
    """Compile csrc/*.cu into librdfwi.so for sm_100a (nvcc cross-compiles without a GPU).  `defines` / `out`: an A/B
    variant of the same sources under another file name (loaded with RDFWI_LIB=<path>)."""
    srcs = [os.path.join(_CSRC, s) for s in SOURCES]
    deps = srcs + HEADERS
    out = out or LIB_PATH
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-I", os.path.join(_ROOT, "include"), "-o", out] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return out


class _Survey(ctypes.Structure):
    _fields_ = [("nz", ctypes.c_int32), ("nx", ctypes.c_int32), ("nbc", ctypes.c_int32), ("ns", ctypes.c_int32),
                ("nrec", ctypes.c_int32), ("nt", ctypes.c_int32), ("sample_temporal", ctypes.c_int32),
                ("isz", ctypes.c_int32), ("igz", ctypes.c_int32), ("dx", ctypes.c_double), ("dt", ctypes.c_double),
                ("isx", ctypes.c_void_p), ("igx", ctypes.c_void_p), ("wavelet", ctypes.c_void_p)]


_lib = None


def load():
    """Load librdfwi.so; raises if it has not been built (run `python __graft_entry__.py` or _cabi.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RdfwiError(f"{LIB_PATH} not found: the CUDA extension is not built. Run __graft_entry__.build(). "
                         "There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t
    lib.rdfwi_version.restype = ctypes.c_int
    lib.rdfwi_last_error.restype = ctypes.c_char_p
    lib.rdfwi_plan_create.argtypes = [ctypes.POINTER(_Survey), ctypes.POINTER(vp)]
    lib.rdfwi_plan_destroy.argtypes = [vp]
    lib.rdfwi_plan_set.argtypes = [vp, ctypes.c_char_p, i64]
    lib.rdfwi_plan_get.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(i64)]
    lib.rdfwi_level_floats.argtypes = [vp]
    lib.rdfwi_level_floats.restype = sz
    lib.rdfwi_workspace_bytes.argtypes = [vp, i32]
    lib.rdfwi_workspace_bytes.restype = sz
    lib.rdfwi_history_bytes.argtypes = [vp, i32, i32]
    lib.rdfwi_history_bytes.restype = sz
    lib.rdfwi_forward.argtypes = [vp, vp, i32, vp, vp, sz, vp, sz, i32, vp]
    lib.rdfwi_backward.argtypes = [vp, vp, i32, vp, vp, vp, sz, vp, sz, i32, vp]
    lib.rdfwi_coefficients.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.rdfwi_misfit_l1.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, sz, vp]
    lib.rdfwi_last_launch_count.restype = i64
    for name in ("rdfwi_plan_create", "rdfwi_plan_destroy", "rdfwi_plan_set", "rdfwi_plan_get", "rdfwi_forward",
                 "rdfwi_backward", "rdfwi_coefficients", "rdfwi_misfit_l1"):
        getattr(lib, name).restype = ctypes.c_int
    _lib = lib
    return lib


def _check(rc, what):
    if rc != 0:
        msg = load().rdfwi_last_error()
        raise RdfwiError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


class Plan:
    """Owns one rdfwi_plan (geometry tables on the current CUDA device)."""

    def __init__(self, nz, nx, nbc, nt, sample_temporal, isz, igz, isx, igx, dx, dt, wavelet):
        lib = load()
        self._isx = np.ascontiguousarray(isx, dtype=np.int32)
        self._igx = np.ascontiguousarray(igx, dtype=np.int32)
        self._wav = np.ascontiguousarray(wavelet, dtype=np.float64)
        s = _Survey()
        s.nz, s.nx, s.nbc, s.ns, s.nrec = int(nz), int(nx), int(nbc), len(self._isx), len(self._igx)
        s.nt, s.sample_temporal, s.isz, s.igz = int(nt), int(sample_temporal), int(isz), int(igz)
        s.dx, s.dt = float(dx), float(dt)
        s.isx, s.igx, s.wavelet = self._isx.ctypes.data, self._igx.ctypes.data, self._wav.ctypes.data
        self.ns, self.nrec, self.nt = s.ns, s.nrec, s.nt
        self.nz, self.nx = s.nz, s.nx
        self.nt_out = (s.nt + s.sample_temporal - 1) // s.sample_temporal
        self._h = ctypes.c_void_p()
        _check(lib.rdfwi_plan_create(ctypes.byref(s), ctypes.byref(self._h)), "rdfwi_plan_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            load().rdfwi_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, key, value):
        _check(load().rdfwi_plan_set(self._h, key.encode(), int(value)), f"rdfwi_plan_set({key})")

    def get(self, key):
        out = ctypes.c_int64()
        _check(load().rdfwi_plan_get(self._h, key.encode(), ctypes.byref(out)), f"rdfwi_plan_get({key})")
        return out.value

    def level_floats(self):
        return load().rdfwi_level_floats(self._h)

    def workspace_bytes(self, B):
        return load().rdfwi_workspace_bytes(self._h, B)

    def history_bytes(self, B, segment=0):
        return load().rdfwi_history_bytes(self._h, B, segment)

    def forward(self, v_ptr, B, seis_ptr, ws_ptr, ws_bytes, hist_ptr, hist_bytes, segment, stream):
        _check(load().rdfwi_forward(self._h, v_ptr, B, seis_ptr, ws_ptr, ws_bytes, hist_ptr, hist_bytes, segment, stream),
               "rdfwi_forward")

    def backward(self, v_ptr, B, cot_ptr, grad_ptr, ws_ptr, ws_bytes, hist_ptr, hist_bytes, segment, stream):
        _check(load().rdfwi_backward(self._h, v_ptr, B, cot_ptr, grad_ptr, ws_ptr, ws_bytes, hist_ptr, hist_bytes,
                                     segment, stream), "rdfwi_backward")

    def coefficients(self, v_ptr, B, alpha_ptr, kap_ptr, velmin_ptr, argmin_ptr, beta_ptr, ws_ptr, ws_bytes, stream):
        _check(load().rdfwi_coefficients(self._h, v_ptr, B, alpha_ptr, kap_ptr, velmin_ptr, argmin_ptr, beta_ptr, ws_ptr,
                                         ws_bytes, stream), "rdfwi_coefficients")

    def misfit_l1(self, seis_ptr, obs_ptr, mask_ptr, B, stats_ptr, sign_ptr, ws_ptr, ws_bytes, stream):
        _check(load().rdfwi_misfit_l1(self._h, seis_ptr, obs_ptr, mask_ptr, B, stats_ptr, sign_ptr, ws_ptr, ws_bytes, stream),
               "rdfwi_misfit_l1")

    @staticmethod
    def last_launch_count():
        return int(load().rdfwi_last_launch_count())
