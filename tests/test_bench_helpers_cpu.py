"""CPU: the pieces of bench.py and baseline/ that can be checked without a GPU -- the parser that turns a committed ncu launch
list into `roofline.traffic`, the kernel-class naming it relies on, the staged-reference loader and its third-party stubs."""
import os

import numpy as np
import pytest

from conftest import ROOT

torch = pytest.importorskip("torch")


def test_kernel_classes_from_ncu_names():
    import bench
    c = bench.classify_kernel
    assert c("void rdfwi::<unnamed>::k_fwd_cluster<13, 312, 0>(rdfwi::ClusterFwdArgs, rdfwi::Grid)") == "forward"
    assert c("void rdfwi::<unnamed>::k_fwd_cluster<13, 432, 1>(rdfwi::ClusterFwdArgs, rdfwi::Grid)") == "adjoint_field"
    assert c("void rdfwi::<unnamed>::k_fwd_cluster<(int)7, (int)0, (bool)1>(rdfwi::ClusterFwdArgs, rdfwi::Grid)") == "adjoint_field"
    assert c("void rdfwi::<unnamed>::k_fwd_cluster<13, 312, 0, 512>(rdfwi::ClusterFwdArgs, rdfwi::Grid)") == "forward"   # round-1 lists
    assert c("void rdfwi::<unnamed>::k_fwd_cluster<13, 312, 2, 0>(rdfwi::ClusterFwdArgs, rdfwi::Grid)") == "adjoint_resident"
    assert c("void rdfwi::<unnamed>::k_fwd_cluster<(int)7, (int)432, (int)2, (bool)0>(rdfwi::ClusterFwdArgs, rdfwi::Grid)") == "adjoint_resident"
    assert c("void rdfwi::<unnamed>::k_fwd_cluster<(int)7, (int)432, (int)1, (bool)0>(rdfwi::ClusterFwdArgs, rdfwi::Grid)") == "adjoint_field"
    assert c("void rdfwi::<unnamed>::k_step_tile<4, 0>(rdfwi::StepArgs, rdfwi::Grid)") == "forward"
    assert c("void rdfwi::<unnamed>::k_step_tile<4, 1>(rdfwi::StepArgs, rdfwi::Grid)") == "adjoint_field"
    assert c("rdfwi::<unnamed>::k_imaging(const float *, ...)") == "imaging"
    assert c("void at::vectorized_elementwise_kernel<4, ...>") is None


def test_committed_launch_lists_parse_and_hold_every_class():
    """Every launch list bench.py names must exist, parse, and contain the kernel classes of its workload with plausible
    per-launch DRAM bytes -- a stale or renamed profile fails here, on CPU, not silently on the GPU box."""
    import bench
    for workload, path in bench.TRAFFIC_PROFILES.items():
        assert os.path.exists(os.path.join(ROOT, path)), f"{path} (roofline.traffic of {workload}) is not committed"
        prof = bench.load_traffic_profile(path)
        assert "forward" in prof and ("adjoint_resident" in prof or {"adjoint_field", "imaging"} <= set(prof)), (workload, sorted(prof))
        for cls, c in prof.items():
            assert c["bytes_per_launch"] > 1e6 and c["ncu_us_per_launch"] > 1.0, (workload, cls, c)
    r1 = bench.load_traffic_profile("profiles/launches_r1_v2_b64.csv")     # the round-1 list the judge recomputed from
    assert abs(r1["forward"]["bytes_per_launch"] - 123.88e9) < 0.2e9 and r1["imaging"]["launches_in_profile"] == 20


def test_attend_stub_is_scaled_dot_product_attention():
    """The one stub that computes something: denoising_diffusion_pytorch.attend.Attend, restated from its published algorithm."""
    from baseline import ref_loader
    ref_loader._install_stubs()
    from denoising_diffusion_pytorch.attend import Attend
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(2, 4, 81, 32, generator=g) for _ in range(3))
    k, v = torch.cat([torch.randn(2, 4, 4, 32, generator=g), k], dim=2), torch.cat([torch.randn(2, 4, 4, 32, generator=g), v], dim=2)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v)
    assert torch.allclose(Attend(flash=False)(q, k, v), ref, atol=1e-5)
    assert torch.allclose(Attend(flash=True)(q, k, v), ref, atol=1e-5)


def test_staged_reference_is_unmodified_and_imports():
    """baseline/_ref holds byte-identical copies (checked against /root/reference where it exists) and the U-Net of the
    reference's configs can be built from them with the third-party stubs (random init, 35.7 M parameters)."""
    import hashlib
    import json
    from baseline import ref_loader, stage_reference
    if not ref_loader.available():
        if not stage_reference.stage(quiet=True):
            pytest.skip("no staged reference and no /root/reference here")
    with open(os.path.join(ref_loader.REF_ROOT, "MANIFEST.json")) as f:
        manifest = json.load(f)
    for rel, digest in manifest["sha256"].items():
        with open(os.path.join(ref_loader.REF_ROOT, rel), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == digest, rel
        src = os.path.join("/root/reference", rel)
        if os.path.exists(src):
            with open(src, "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == digest, f"{rel} differs from the reference"
    dm = ref_loader.build_diffusion(torch.device("cpu"))
    n = sum(p.numel() for p in dm.parameters())
    assert 35e6 < n < 36.5e6 and tuple(dm.image_size) == (72, 72) and dm.num_timesteps == 1000
    pde = ref_loader.load("red_diffeq.solvers.pde")
    assert hasattr(pde, "FWIForward")
    x = torch.zeros(1, 1, 72, 72)
    with torch.no_grad():
        out = dm.model_predictions(dm.q_sample(x, t=torch.tensor([10]), noise=torch.zeros_like(x)), t=torch.tensor([10]),
                                   x_self_cond=None, clip_x_start=True, rederive_pred_noise=True)
    assert out.pred_noise.shape == x.shape and torch.isfinite(out.pred_noise).all()
