"""CPU: host-side logic of the operator (geometry, wavelet, ctx contract), the C ABI surface, and the
shot/model partitioner -- everything that needs no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, Golden

torch = pytest.importorskip("torch")


def test_ricker_matches_oracle_and_reference_shape(oracle):
    from red_diffeq_b200.solvers import survey
    w = survey.ricker(15.0, 0.001, 1000)
    assert w.dtype == np.float64 and w.shape == (1000,)
    assert np.array_equal(w, oracle.ricker_wavelet(15.0, 0.001, 1000))
    assert np.count_nonzero(w) == 147 and np.argmax(w) == 73 and w[73] == 1.0
    with pytest.raises(ValueError):
        survey.ricker(15.0, 0.001, 146)


def test_grid_indices_round_half_to_even():
    from red_diffeq_b200.solvers import survey
    sx = np.linspace(0, 69, num=5) * 10.0
    isx, isz, igx, igz = survey.grid_indices(sx, 10, np.linspace(0, 69, num=70) * 10.0, 10, 10.0, 120)
    assert list(isx) == [120, 137, 154, 172, 189]          # 17.25->17, 34.5->34, 51.75->52
    assert isz == 121 and igz == 121 and list(igx) == list(range(120, 190))
    isx, *_ = survey.grid_indices(np.linspace(0, 189, num=5) * 10.0, 10, [0.0], 10, 10.0, 120)
    assert list(isx) == [120, 167, 214, 262, 309]          # 94.5 -> 94


def test_ctx_is_completed_in_place_like_the_reference():
    from red_diffeq_b200.solvers import survey
    ctx = dict(n_grid=70, nt=1000, dx=10.0, dt=0.001, nbc=120, f=15.0, sz=10, gz=10, ng=70, ns=5)
    out = survey.complete_ctx(ctx, sample_spatial=0.5)
    assert out is ctx and len(ctx["sx"]) == 5 and len(ctx["gx"]) == 35
    assert ctx["sx"][-1] == 690.0 and ctx["gx"][0] == 0.0
    ctx2 = dict(ctx, sx=[1, 2.5], gx=[3, 3, 4])
    survey.complete_ctx(ctx2)
    assert list(ctx2["sx"]) == [10.0, 25.0] and list(ctx2["gx"]) == [30.0, 30.0, 40.0]   # grid units * dx
    with pytest.raises(IndexError):
        survey.wrap_indices(np.array([400]), 310, "source column")
    assert list(survey.wrap_indices(np.array([-1, 5]), 310, "x")) == [309, 5]


def test_cabi_exports_every_declared_symbol():
    from red_diffeq_b200 import _cabi
    header = open(os.path.join(ROOT, "include", "rdfwi.h")).read()
    declared = sorted(set(re.findall(r"\b(rdfwi_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(_cabi.EXPORTS)
    lib = ctypes.CDLL(_cabi.LIB_PATH)      # loads on a CPU-only box (static cudart); no compute call is made
    for name in declared:
        assert hasattr(lib, name), name
    assert _cabi.load().rdfwi_version() == 100


def test_no_cpu_fallback():
    from red_diffeq_b200 import FWIForward
    g = Golden("tiny_default")
    with pytest.raises(RuntimeError, match="CUDA"):
        FWIForward(g.fresh_ctx(), "cpu", normalize=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        FWIForward(g.fresh_ctx(), torch.device("cpu"), normalize=False)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "red-diffeq_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("# oracle", ""), f"{f} mentions the oracle"


def test_partition_plans():
    from red_diffeq_b200.solvers.sharding import plan_partition, split_range
    assert [split_range(5, 2, r) for r in range(2)] == [(0, 3), (3, 5)]
    assert [split_range(5, 4, r) for r in range(4)] == [(0, 2), (2, 3), (3, 4), (4, 5)]
    assert [split_range(3, 8, r)[1] - split_range(3, 8, r)[0] for r in range(8)] == [1, 1, 1, 0, 0, 0, 0, 0]
    covered = []
    for r in range(8):
        mode, models, shots = plan_partition(64, 5, 8, r)
        assert mode == "models" and len(shots) == 5
        covered += list(range(models.start, models.stop))
    assert covered == list(range(64))
    shots_seen = []
    for r in range(4):
        mode, models, shots = plan_partition(1, 40, 4, r)
        assert mode == "shots" and (models.start, models.stop) == (0, 1)
        shots_seen += list(shots)
    assert shots_seen == list(range(40))
    with pytest.raises(ValueError):
        plan_partition(1, 5, 2, 0, mode="rows")


def test_c_example_compiles_and_links_against_the_library(tmp_path):
    """include/rdfwi.h is valid C99 and the plain-C example of the ABI (examples/c_abi_example.c: plan, forward, adjoint with
    caller-owned buffers, no Python / torch) links against librdfwi.so; it is run on the GPU by tests/test_gpu_parity.py."""
    import shutil
    import subprocess
    from red_diffeq_b200 import _cabi
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    obj = str(tmp_path / "example.o")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c",
                           os.path.join(ROOT, "examples", "c_abi_example.c"), "-o", obj])
    cudart = "/usr/local/cuda/lib64"
    if not os.path.exists(os.path.join(cudart, "libcudart.so")):
        pytest.skip("no CUDA runtime to link the example's cudaMalloc / cudaMemcpy against")
    subprocess.check_call(["gcc", obj, "-o", str(tmp_path / "example"), "-L", os.path.dirname(_cabi.LIB_PATH), "-lrdfwi",
                           "-L", cudart, "-lcudart", "-lm"])
