#!/usr/bin/env python
"""Stages the UNMODIFIED reference modules the benchmark's reference legs need under baseline/_ref/ (git-ignored; it
travels to the GPU box with the working tree, /root/reference does not).

    python baseline/stage_reference.py [--reference /root/reference]

What is staged (byte-for-byte copies; nothing is edited, nothing is committed):
    red_diffeq/solvers/pde.py                     the reference operator, timed on the host CPU beside the GPU path
    red_diffeq/models/diffusion.py                the reference's U-Net + GaussianDiffusion (random-init in the bench)
    red_diffeq/regularization/{base,diffusion,benchmark}.py, red_diffeq/utils/{diffusion_utils,data_trans,ssim}.py,
    red_diffeq/core/{inversion,losses,metrics}.py the reference's regulariser call pattern and loop, for A/B timing
The package __init__ files are written EMPTY here: the reference's own red_diffeq/__init__.py imports its config system
(ml_collections) which is neither on the path nor needed.  The reference has no setup.py / pyproject.toml, so the
`pip install --target baseline/_ref` route of the base contract does not apply.
Third-party modules the staged files import but this image lacks are stubbed by baseline/ref_loader.py at import time.
"""
import argparse
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = [
    "red_diffeq/solvers/pde.py",
    "red_diffeq/models/diffusion.py",
    "red_diffeq/regularization/base.py",
    "red_diffeq/regularization/diffusion.py",
    "red_diffeq/regularization/benchmark.py",
    "red_diffeq/utils/diffusion_utils.py",
    "red_diffeq/utils/data_trans.py",
    "red_diffeq/utils/ssim.py",
    "red_diffeq/core/inversion.py",
    "red_diffeq/core/losses.py",
    "red_diffeq/core/metrics.py",
]
PACKAGES = ["red_diffeq", "red_diffeq/solvers", "red_diffeq/models", "red_diffeq/regularization", "red_diffeq/utils",
            "red_diffeq/core"]


def stage(reference="/root/reference", quiet=False):
    """Copies FILES from `reference` into baseline/_ref; returns True when something was staged, False when the
    reference tree is not there (the GPU box: it uses the files staged in the build container)."""
    if not os.path.isdir(os.path.join(reference, "red_diffeq")):
        return False
    manifest = {}
    for pkg in PACKAGES:
        os.makedirs(os.path.join(DEST, pkg), exist_ok=True)
        with open(os.path.join(DEST, pkg, "__init__.py"), "w") as f:
            f.write("")
    for rel in FILES:
        src, dst = os.path.join(reference, rel), os.path.join(DEST, rel)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": reference, "sha256": manifest}, f, indent=1)
    if not quiet:
        print(f"staged {len(FILES)} reference files under {DEST}")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    if not stage(args.reference):
        raise SystemExit(f"{args.reference}/red_diffeq not found")
