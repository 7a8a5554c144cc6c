"""Regular N-D node grid with a global column index (host-side geometry object).

Mirrors the hot-path part of the reference class ``LSsurf.fd_grid.fd_grid`` (LSsurf/fd_grid.py:13-145):
same constructor signature and attribute names, same rounding rules for the node count
(float -> int truncation, fd_grid.py:61) and for the bounds (reset to the first / last node centre,
fd_grid.py:63).  The grid itself is a few hundred bytes of host state; the per-point work
(validation, cell lookup, global index, interpolation weights) runs on the GPU inside
``lin_op.interp_mtx`` -- ``as_struct()`` is what crosses the C ABI.  The small NumPy methods below
exist for API compatibility with callers that use them on a handful of points.

Out of scope: geotiff / shapefile mask I/O (needs GDAL).  A mask given as
a NumPy array is supported.
"""
import copy as _copy

import numpy as np

from . import _cabi


class fd_grid(object):
    def __init__(self, bounds, deltas, name='', col_0=0, col_N=None, srs_proj4=None,
                 mask_file=None, mask_data=None, mask_interp_threshold=0.5,
                 erode_source_mask=True, xform=None, coords=['y', 'x', 'time']):
        # node count: ((hi - lo) / delta) + 1 evaluated in floating point, then truncated (fd_grid.py:61)
        self.shape = np.array([((b[1] - b[0]) / d) + 1 for b, d in zip(bounds, deltas)]).astype(int)
        self.ctrs = [b[0] + d * np.arange(n) for b, d, n in zip(bounds, deltas, self.shape)]
        self.bds = [np.array([c[0], c[-1]]) for c in self.ctrs]
        self.delta = np.array(deltas)
        self.N_dims = len(self.shape)
        self.N_nodes = np.prod(self.shape)
        stride = np.ones(self.N_dims, dtype=self.shape.dtype)
        for d in range(self.N_dims - 2, -1, -1):
            stride[d] = stride[d + 1] * self.shape[d + 1]
        self.stride = stride                                                        # fd_grid.py:67
        self.col_0 = col_0
        self.col_N = self.col_0 + self.N_nodes if col_N is None else col_N
        self.srs_proj4 = srs_proj4
        self.user_data = dict()
        self.name = name
        self.coords = coords
        self.mask_interp_threshold = mask_interp_threshold
        self.cell_area = None
        self.xform = xform
        self.mask = np.ones(self.shape[0:2], dtype=bool)
        self.mask_3d = None
        self.erode_source_mask = erode_source_mask
        self.mask_file = None
        self.setup_mask(mask_data=mask_data, mask_file=mask_file, interp_threshold=mask_interp_threshold)

    # ------------------------------------------------------------------ C-ABI view
    def as_struct(self):
        """The lss_grid_t handed to the CUDA kernels (include/lssurf_b200.h)."""
        if self.N_dims > _cabi.LSS_MAX_DIMS:
            raise ValueError('grids of more than 3 dimensions are not supported')
        return _cabi.grid_struct(self.shape, self.stride, [b[0] for b in self.bds], [b[1] for b in self.bds],
                                 self.delta, self.col_0)

    def copy(self):
        return _copy.deepcopy(self)

    # ------------------------------------------------------------------ small host helpers
    def validate_pts(self, pts):
        """Closed-bounds test with a finite first coordinate (fd_grid.py:90-96)."""
        good = np.isfinite(pts[0])
        for dim in range(self.N_dims):
            p = np.asarray(pts[dim])
            with np.errstate(invalid='ignore'):
                good = good & (p >= self.bds[dim][0]) & (p <= self.bds[dim][1])
        return good

    def pos_for_nodes(self, nodes):
        inds = np.unravel_index(nodes, self.shape)
        return [ind * d + bd[0] for d, bd, ind in zip(self.delta, self.bds, inds)]

    def get_extent(self, dims=[1, 0]):
        return np.concatenate([self.bds[dim] + self.delta[dim] * np.array([-0.5, 0.5]) for dim in dims])

    def float_sub(self, pts, good=None):
        """Point position in subscript coordinates, NaN for invalid points (fd_grid.py:111-118)."""
        if good is None:
            good = self.validate_pts(pts)
        out = []
        for dim in range(self.N_dims):
            v = np.full(np.shape(pts[0]), np.nan)
            v[good] = (np.asarray(pts[dim])[good] - self.bds[dim][0]) / self.delta[dim]
            out.append(v)
        return out

    def cell_sub_for_pts(self, pts, good=None):
        """Subscript of the node below each point (fd_grid.py:120-128)."""
        return [np.floor(v) for v in self.float_sub(pts, good)]

    def global_ind(self, cell_sub, return_valid=False):
        """Global column of a node from its subscripts (fd_grid.py:130-145)."""
        sub = [np.asarray(s).astype(int) for s in cell_sub]
        if not return_valid:
            return self.col_0 + np.ravel_multi_index(sub, self.shape)
        valid = np.ones_like(sub[0], dtype=bool)
        for dim in range(self.N_dims):
            valid &= (sub[dim] >= 0) & (sub[dim] < self.shape[dim])
        ind = np.zeros_like(sub[0])
        ind[valid] = self.col_0 + np.ravel_multi_index([s[valid] for s in sub], self.shape)
        return ind, valid

    # ------------------------------------------------------------------ masks
    def setup_mask(self, mask_data=None, mask_file=None, interp_threshold=0.5):
        if mask_data is None and mask_file is None:
            return
        if mask_data is None:
            raise NotImplementedError('mask files (geotiff / shapefile) need GDAL and are out of scope; '
                                      'pass mask_data as a NumPy array')
        if isinstance(mask_data, np.ndarray):
            if mask_data.ndim > 2:                                                  # fd_grid.py:241-246
                self.mask = mask_data.astype(bool)
                self.mask_3d = mask_data.astype(bool)
            else:
                if tuple(mask_data.shape) != tuple(self.shape[0:2]):
                    raise ValueError('mask_data must have the shape of the first two grid dimensions')
                self.mask = mask_data.astype(bool)
            return
        raise NotImplementedError('only NumPy-array masks are supported (pointCollection grids need interpolation '
                                  'code that is out of scope)')
