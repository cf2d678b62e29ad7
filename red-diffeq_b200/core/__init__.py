from .inversion import InversionEngine

__all__ = ["InversionEngine"]
