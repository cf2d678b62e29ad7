"""The two normalisation callables the reference's drivers inject into FWIForward
(utils/data_trans.py:8-20 there; scripts/run_inversion.py:79-85)."""


def v_normalize(v):
    """m/s in [1500, 4500] -> [-1, 1]."""
    return (v - 1500) / 3000 * 2 - 1


def v_denormalize(v_norm):
    """[-1, 1] -> m/s in [1500, 4500]."""
    return (v_norm + 1) / 2 * 3000 + 1500


def s_normalize_none(s):
    """Seismograms are used unscaled."""
    return s
