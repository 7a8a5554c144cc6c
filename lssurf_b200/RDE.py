"""Robust dispersion estimate on the GPU (API mirror of LSsurf/RDE.py:10-18)."""
import numpy as np
import torch

from . import _cabi


def RDE(x):
    """Half the width between the 16th and 84th percentiles of the finite values of ``x``.

    Same definition as the reference (linear interpolation at 0.16 n and 0.84 n on abscissae
    0.5, 1.5, ...); computed by ``lss_rde``: exact radix selection of the (at most four) order
    statistics involved, no sort.  Accepts a NumPy array or a CUDA tensor; returns a float
    (NaN when fewer than two finite values).
    """
    t = x if isinstance(x, torch.Tensor) else np.asarray(x, dtype=float).ravel()
    val, _ = _cabi.rde(_cabi.to_dev(t, torch.float64).reshape(-1))
    return np.float64(val)
