"""Prior constraints from an existing dz model (API mirror of LSsurf/match_priors.py:26-44 and the in-memory branch of
``match_prior_dz``, :46-130).

``make_prior_op`` ties the fitted height change at the prior's points to the prior values: an interpolation row at (y, x, t)
minus the interpolation row at (y, x, ref_time), expected misfit ``sigma_scale * sigma_dz``, prior value ``dz``.  Both
rows come from ``lss_interp_coo`` (lin_op.interp_mtx).  The file-reading parts of the reference (h5 tiles, ``match_tile_edges``)
need pointCollection I/O and are out of scope: priors are handed over as point containers.
"""
import numpy as np

from .lin_op import lin_op
from .point_data import data as _point_data


def make_prior_op(grids, dz, src_name, ref_time=0, sigma_scale=1):
    """dz: point container with fields x, y, t (or time), dz, sigma_dz (match_priors.py:26-44)."""
    if 't' not in dz.fields:
        dz.assign(t=dz.time)
    dz = dz[np.asarray(dz.t) != ref_time]
    y, x, t = (np.asarray(getattr(dz, f), dtype=float) for f in ('y', 'x', 't'))
    m1 = lin_op(grids['dz'], name='prior_dz1_' + src_name).interp_mtx((y, x, t))
    m0 = lin_op(grids['dz'], name='prior_dz0_' + src_name).interp_mtx((y, x, np.zeros_like(t) + ref_time))
    m0.v *= -1
    m1.add(m0)
    m1.expected = sigma_scale * np.asarray(dz.sigma_dz, dtype=float)
    m1.prior = np.asarray(dz.dz, dtype=float)
    return m1


def match_prior_dz(grids, dzs=None, filenames=None, ref_epoch=0, sigma_scale=1, sigma_max=None, **kw):
    """One prior operator per in-memory data set in ``dzs`` (point containers: x, y, t, dz, sigma_dz).  Points outside the
    dz grid and (with ``sigma_max``) points with larger sigma_dz are dropped first, like the gridded branch of the
    reference does after flattening its tiles (match_priors.py:96-128)."""
    if filenames:
        raise NotImplementedError('prior files (h5 tiles) need pointCollection I/O: out of scope; pass point containers in `dzs`')
    ops = []
    for k, dz in enumerate(dzs or []):
        if dz is None:
            continue
        if isinstance(dz, dict):
            dz = _point_data().from_dict({k: np.asarray(v) for k, v in dz.items()})
        if 't' not in dz.fields:
            dz.assign(t=dz.time)
        good = grids['dz'].validate_pts((np.asarray(dz.y), np.asarray(dz.x), np.asarray(dz.t)))
        good &= np.isfinite(np.asarray(dz.dz)) & np.isfinite(np.asarray(dz.sigma_dz))
        if sigma_max is not None:
            good &= np.asarray(dz.sigma_dz) < sigma_max
        if not np.any(good):
            continue
        ops.append(make_prior_op(grids, dz[good], 'prior_%d' % k, ref_time=ref_epoch, sigma_scale=sigma_scale))
    return ops


def match_tile_edges(grids, ref_epoch, prior_dir=None, **kw):
    """Constraints tying the edges of a tile to its already-fitted neighbours (match_priors.py:109-224).  The reference reads
    the neighbouring tiles from h5 files with pointCollection; that I/O is outside this package.  Build the same constraints
    from in-memory models with ``make_prior_op`` (one call per neighbour) and pass them through ``prior_args={'dzs': [...]}``."""
    raise NotImplementedError('match_tile_edges reads neighbouring tiles from h5 files (pointCollection I/O): out of scope; '
                              'use make_prior_op / prior_args with in-memory dz models')
