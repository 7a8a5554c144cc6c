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


def sigma_extra_on_grid_device(x, y, r, sigma, mask, group_masks=None, spacing=1.e4, L_avg=None, sigma_extra_max=None,
                               comm=None, dist=None):
    """Spatially varying sigma_extra (calc_sigma_extra.py:47-109) on device arrays.

    One ``calc_sigma_extra`` over the edited-in points of every overlapping L_avg x L_avg bin centred on the occupied
    nodes of a ``spacing`` lattice, blended with separable linear tapers; points no bin reaches keep the all-data value.
    Bins are visited in the order ``np.unique`` gives their complex centres (x first, then y) so the weighted sums
    accumulate in the reference's order.  Every bin works on the compacted subset of its points, so the total quantile
    work is ~(L_avg / spacing)^2 passes over the data.  Returns (tensor, {group: all-data value, 'grid_max': max}).

    With ``comm`` / ``dist`` (points sharded over ranks) the bins are global: lattice origin and the set of occupied nodes
    are agreed over the ranks, every rank visits every bin with its own (possibly empty) subset, the order statistics and the
    10-point rule see all ranks' points (histogram all-reduces inside ``lss_sigma_extra``), the blend is per point.
    """
    multi = dist is not None and dist.world > 1
    if L_avg is None:
        L_avg = 2 * spacing
    half = L_avg / 2
    out, values = sigma_extra_device(r, sigma, mask, group_masks, comm)
    # the reference rounds the COMPLEX quotient (x + 1j * y) / spacing (calc_sigma_extra.py:86); NumPy's complex division by
    # a real divisor multiplies both parts by 1 / spacing, so the bin index is round(x * (1 / spacing)), not round(x / spacing)
    inv_spacing = 1.0 / float(spacing)
    ix, iy = torch.round(x * inv_spacing), torch.round(y * inv_spacing)
    big = torch.full((1,), 1e300, dtype=torch.float64, device=r.device)
    lims = torch.cat([ix.min().reshape(1), iy.min().reshape(1), -iy.max().reshape(1)]) if x.numel() else torch.cat([big, big, big])
    if multi:
        lims = dist.allreduce_min(lims)
    ix0, iy0, iy1 = float(lims[0].item()), float(lims[1].item()), -float(lims[2].item())
    ny = int(iy1 - iy0) + 1
    keys = torch.unique(((ix - ix0) * ny + (iy - iy0)).to(torch.int64)).cpu().numpy()
    if multi:
        keys = np.unique(dist.gather_rows(keys))
    sum_W_sigma, sum_W = torch.zeros_like(r), torch.zeros_like(r)
    m8 = mask.to(torch.uint8)
    for key in keys:
        cx, cy = (ix0 + key // ny) * spacing, (iy0 + key % ny) * spacing
        dx, dy = torch.abs(cx - x), torch.abs(cy - y)
        idx = torch.nonzero((dx < half) & (dy < half)).reshape(-1)
        m_sub = m8[idx].contiguous()
        n_sel = m_sub.sum().to(torch.float64).reshape(1)
        if multi:
            n_sel = dist.allreduce_sum(n_sel)
        if int(n_sel.item()) < 10:
            continue
        g_sub = None if group_masks is None else {k: g[idx].contiguous() for k, g in group_masks.items()}
        this, _ = sigma_extra_device(r[idx].contiguous(), sigma[idx].contiguous(), m_sub, g_sub, comm)
        if sigma_extra_max is not None:
            this = torch.clamp(this, max=sigma_extra_max)
        W = (1 - _cabi.true_div(dx[idx], float(half))) * (1 - _cabi.true_div(dy[idx], float(half)))
        sum_W_sigma[idx] += W * this
        sum_W[idx] += W
    out = torch.where(sum_W > 0, sum_W_sigma / sum_W, out)
    top = out.max().reshape(1) if out.numel() else torch.zeros(1, dtype=torch.float64, device=r.device)
    if multi:
        top = -dist.allreduce_min(-top)
    values = dict(values, grid_max=float(top.item()))
    return out, values


def calc_sigma_extra_on_grid(x, y, r, sigma, in_TSE, sigma_extra_masks=None, spacing=1.e4, L_avg=None, sigma_extra_max=None):
    """NumPy in, NumPy out (reference signature, calc_sigma_extra.py:47-49); the work is done on the GPU."""
    f64 = torch.float64
    dev = [_cabi.to_dev(np.asarray(a, dtype=float), f64) for a in (x, y, r, sigma)]
    m_d = _cabi.to_dev(np.asarray(in_TSE).astype(np.uint8), torch.uint8)
    groups = None
    if sigma_extra_masks is not None:
        groups = {k: _cabi.to_dev(np.asarray(v).astype(np.uint8), torch.uint8) for k, v in sigma_extra_masks.items()}
    out, _ = sigma_extra_on_grid_device(*dev, m_d, groups, spacing=spacing, L_avg=L_avg, sigma_extra_max=sigma_extra_max)
    return _cabi.to_host(out)
