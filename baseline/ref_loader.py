"""Imports the reference modules staged under baseline/_ref (see stage_reference.py) in an image that lacks five of the
third-party packages red_diffeq/models/diffusion.py:14,22-27 imports at module scope.

Only bench.py's reference legs and the tests use this module; the product package never does.

Stubbed (absent here: matplotlib, ema_pytorch, accelerate, denoising_diffusion_pytorch; pinned by the reference's
requirements.txt as matplotlib==3.9.2, ema-pytorch==0.5.3, accelerate==0.33.0, denoising-diffusion-pytorch==2.1.1):
  * matplotlib.pyplot, ema_pytorch.EMA, accelerate.Accelerator, denoising_diffusion_pytorch.fid_evaluation.FIDEvaluation,
    denoising_diffusion_pytorch.version.__version__ -- used only by the reference's Trainer class (training, sampling
    grids, FID), which the inversion path never instantiates: placeholders that raise when called;
  * denoising_diffusion_pytorch.attend.Attend -- used by the U-Net's full-attention blocks (models/diffusion.py:204,216).
    Restated from the published algorithm of denoising-diffusion-pytorch 2.1.1 `attend.py`: with flash=False (every config
    of the reference: configs/*/red-diffeq.yaml `flash_attn: false`) it is softmax(q k^T / sqrt(d)) v evaluated with two
    einsums; with flash=True it is torch's F.scaled_dot_product_attention.  No dropout at inference (dropout=0 default).
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF_ROOT, "red_diffeq", "solvers", "pde.py"))


def _placeholder(name):
    class _Missing:
        def __init__(self, *a, **k):
            raise RuntimeError(f"{name} is a placeholder: the package is not installed in this image and the inversion "
                               "path does not use it")
    _Missing.__name__ = name.rsplit(".", 1)[-1]
    return _Missing


def _install_stubs():
    import torch
    from torch import nn
    import torch.nn.functional as F

    def module(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        try:
            return importlib.import_module(name)
        except Exception:
            pass
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__rdfwi_stub__ = True
        sys.modules[name] = m
        parent, _, child = name.rpartition(".")
        if parent:
            setattr(module(parent), child, m)
        return m

    plt = module("matplotlib.pyplot")
    module("matplotlib").pyplot = plt
    module("ema_pytorch", EMA=_placeholder("ema_pytorch.EMA"))
    module("accelerate", Accelerator=_placeholder("accelerate.Accelerator"))
    ddp = module("denoising_diffusion_pytorch")
    if getattr(ddp, "__rdfwi_stub__", False):
        ddp.__path__ = []

    class Attend(nn.Module):
        """softmax(q k^T * scale) v over (b, h, n, d) tensors (denoising-diffusion-pytorch 2.1.1, attend.py)."""

        def __init__(self, dropout=0.0, flash=False, scale=None):
            super().__init__()
            self.dropout, self.flash, self.scale = dropout, flash, scale
            self.attn_dropout = nn.Dropout(dropout)

        def forward(self, q, k, v):
            if self.flash:
                return F.scaled_dot_product_attention(q, k, v, dropout_p=self.dropout if self.training else 0.0, scale=self.scale)
            scale = self.scale if self.scale is not None else q.shape[-1] ** -0.5
            sim = torch.einsum("b h i d, b h j d -> b h i j", q, k) * scale
            attn = self.attn_dropout(sim.softmax(dim=-1))
            return torch.einsum("b h i j, b h j d -> b h i d", attn, v)

    module("denoising_diffusion_pytorch.attend", Attend=Attend)
    module("denoising_diffusion_pytorch.fid_evaluation", FIDEvaluation=_placeholder("FIDEvaluation"))
    module("denoising_diffusion_pytorch.version", __version__="2.1.1 (stub)")


def load(*names):
    """Returns the staged reference modules, e.g. load("red_diffeq.solvers.pde", "red_diffeq.models.diffusion")."""
    if not available():
        raise FileNotFoundError(f"{REF_ROOT} is empty: run `python baseline/stage_reference.py` where /root/reference exists")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _install_stubs()
    mods = [importlib.import_module(n) for n in names]
    return mods[0] if len(mods) == 1 else mods


def build_diffusion(device, dim=64, dim_mults=(1, 2, 4, 8), channels=1, flash_attn=False, image_size=72, timesteps=1000,
                    sampling_timesteps=250, objective="pred_noise", seed=8888):
    """The reference's U-Net + GaussianDiffusion at the sizes of configs/openfwi/red-diffeq.yaml:12-27 (= marmousi,
    overthrust), built exactly like scripts/run_inversion.py:39-55, RANDOM-INIT (the weights are not in the repository;
    SURVEY.md 8d config 2 allows this), eval mode."""
    import torch
    m = load("red_diffeq.models.diffusion")
    torch.manual_seed(seed)
    unet = m.Unet(dim=dim, dim_mults=tuple(dim_mults), flash_attn=flash_attn, channels=channels)
    diffusion = m.GaussianDiffusion(unet, image_size=image_size, timesteps=timesteps, sampling_timesteps=sampling_timesteps,
                                    objective=objective)
    return diffusion.to(device).eval()
