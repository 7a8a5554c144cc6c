"""sigma_extra: the error to add in quadrature so that the robust spread of scaled residuals is 1.

API mirror of LSsurf/calc_sigma_extra.py:13-44.  The bounded Brent search (scipy's
``minimize_scalar(method='bounded')`` in the reference) runs in the C++ host code of
``lss_sigma_extra`` with scipy's constants; every function evaluation is an exact GPU quantile
selection over the selected points (no sort), so ~10-30 evaluations cost a few ms at 1e7 points.
"""
import numpy as np
import torch

from . import _cabi


def sigma_extra_device(r, sigma, mask, group_masks=None, comm=None, out=None):
    """Device version: r, sigma fp64 CUDA tensors, mask uint8 CUDA tensor; returns (tensor, {group: value})."""
    if out is None:
        out = torch.zeros_like(r)
    else:
        out.zero_()
    values = {}
    if group_masks is None:
        val, _ = _cabi.sigma_extra(r, sigma, mask, comm)
        values['all'] = val
        out.fill_(val)                      # assigned to every point of the group, edited or not (:41)
        return out, values
    for key, grp in group_masks.items():
        sel = (grp & mask).contiguous()
        val, nfev = _cabi.sigma_extra(r, sigma, sel, comm)
        if nfev == 0:                       # fewer than 10 selected points: group skipped (:34)
            continue
        values[key] = val
        out[grp.bool()] = val
    return out, values


def calc_sigma_extra(r, sigma, mask, sigma_extra_masks=None):
    """NumPy in, NumPy out (reference signature); the work is done on the GPU."""
    r_d = _cabi.to_dev(np.asarray(r, dtype=float), torch.float64)
    s_d = _cabi.to_dev(np.asarray(sigma, dtype=float), torch.float64)
    m_d = _cabi.to_dev(np.asarray(mask).astype(np.uint8), torch.uint8)
    groups = None
    if sigma_extra_masks is not None:
        groups = {k: _cabi.to_dev(np.asarray(v).astype(np.uint8), torch.uint8) for k, v in sigma_extra_masks.items()}
    out, _ = sigma_extra_device(r_d, s_d, m_d, groups)
    return out.cpu().numpy()
