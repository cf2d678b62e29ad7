from .diffusion import REDDiffEq, calculate_patches

__all__ = ["REDDiffEq", "calculate_patches"]
