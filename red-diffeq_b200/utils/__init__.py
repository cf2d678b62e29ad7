from .data_trans import s_normalize_none, v_denormalize, v_normalize

__all__ = ["v_normalize", "v_denormalize", "s_normalize_none"]
