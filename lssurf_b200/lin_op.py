"""Sparse linear operators on fd_grid nodes, held as device-resident COO triplets.

API mirror of the reference class ``LSsurf.lin_op.lin_op`` (LSsurf/lin_op.py:12-753) for the
smooth_fit hot path: same constructor, attribute names (``r, c, v, N_eq, col_0, col_N, row_0, name,
ind0, TOC, grid, dst_grid, dst_ind0, expected, prior, shape``), builder methods that mutate and
return ``self``, and ``toCSR()`` returning a ``scipy.sparse.csr_matrix``.

What is different is where the work happens.  The triplets live in GPU memory (int32 / int32 / fp64
torch tensors used purely as buffers); they are produced by the CUDA kernels ``lss_interp_coo`` and
``lss_stencil_coo`` and reduced to canonical CSR by ``lss_coo_to_csr`` (radix sort + segmented
reduce), all through the C ABI in ``include/lssurf_b200.h``.  ``op.r / op.c / op.v`` materialise
NumPy copies on demand (int64 / int64 / float64, the reference's dtypes and 2-D shapes) and can be
assigned to; after such an access the host copy is authoritative and is re-uploaded when a kernel
needs it, so code that edits ``op.v`` in place (the reference notebooks do) keeps working.

There is no CPU implementation of the kernels: without the CUDA library every builder raises.
"""
import numpy as np
import scipy.sparse as sp
import torch

from . import _cabi
from .fd_grid import fd_grid

_FIELDS = ('r', 'c', 'v')
_DEV_DTYPE = {'r': torch.int32, 'c': torch.int32, 'v': torch.float64}
_HOST_DTYPE = {'r': np.int64, 'c': np.int64, 'v': np.float64}


def _index_array(rows):
    """TOC row lists as an int array; ``range`` objects (what diff_op / interp_mtx store) without a Python loop."""
    if isinstance(rows, range):
        return np.arange(rows.start, rows.stop, rows.step, dtype='int')
    return np.array(rows, dtype='int')


def _join_rows(parts):
    """Concatenation of TOC row lists; consecutive ranges merge into one range."""
    if not parts:
        return np.array([], dtype=int)
    if all(isinstance(p, range) and p.step == 1 for p in parts):
        nonempty = [p for p in parts if len(p)]
        if not nonempty:
            return range(0)
        if all(a.stop == b.start for a, b in zip(nonempty[:-1], nonempty[1:])):
            return range(nonempty[0].start, nonempty[-1].stop)
    return np.concatenate([_index_array(p).ravel() for p in parts])


def rows_of(rows):
    """An index usable on NumPy arrays / tensors for a TOC row list: a slice for contiguous ranges."""
    if isinstance(rows, range) and rows.step == 1:
        return slice(rows.start, rows.stop)
    return _index_array(rows)


def _union_of_columns(arrays):
    """np.unique(np.concatenate(arrays)); column lists are almost always contiguous ranges, whose union is merged
    as intervals instead of sorting millions of integers."""
    if not arrays:
        return np.array([], dtype=int)
    spans = []
    for a in arrays:
        a = np.asarray(a)
        if a.size == 0:
            continue
        if a[-1] - a[0] == a.size - 1 and (a.size < 3 or (a[1] - a[0] == 1 and a[a.size // 2] - a[0] == a.size // 2)):
            spans.append((int(a[0]), int(a[-1]) + 1))
        else:
            return np.unique(np.concatenate(arrays))
    spans.sort()
    merged = []
    for lo, hi in spans:
        if merged and lo <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], hi)
        else:
            merged.append([lo, hi])
    return np.concatenate([np.arange(lo, hi) for lo, hi in merged]) if merged else np.array([], dtype=int)


class DeviceCSR(object):
    """Canonical CSR on the GPU (int32 indptr / indices, fp64 data) plus a lazily built transpose."""

    def __init__(self, indptr, indices, data, shape):
        self.indptr, self.indices, self.data = indptr, indices, data
        self.shape = (int(shape[0]), int(shape[1]))
        self._T = None
        self._plan = None

    @property
    def nnz(self):
        return int(self.data.numel())

    def transpose(self):
        if self._T is None:
            t = _cabi.csr_transpose(self.indptr, self.indices, self.data, self.shape[0], self.shape[1])
            self._T = DeviceCSR(t[0], t[1], t[2], (self.shape[1], self.shape[0]))
            self._T._T = self
        return self._T

    def dot(self, x, mode=_cabi.SPMV_PLAIN, s=None, out=None):
        """y = A x on the device (x: CUDA fp64 tensor)."""
        if self._plan is None:
            self._plan = _cabi.SpmvPlan(self.indptr, self.shape[0])
        return _cabi.spmv_planned(self._plan, self.indptr, self.indices, self.data, x, self.shape[0], mode, s, out)

    def rdot(self, u):
        """A^T u without building the transpose (scatter kernel; use transpose().dot for repeated products)."""
        return _cabi.csr_rmatvec(self.indptr, self.indices, self.data, u, self.shape[0], self.shape[1])

    def row_slice(self, lo, hi):
        """Rows [lo, hi) as a new DeviceCSR (used to shard rows over ranks)."""
        b = int(self.indptr[lo].item())
        e = int(self.indptr[hi].item())
        return DeviceCSR((self.indptr[lo:hi + 1] - b).contiguous(), self.indices[b:e].contiguous(),
                         self.data[b:e].contiguous(), (hi - lo, self.shape[1]))

    def to_scipy(self):
        return sp.csr_matrix((self.data.cpu().numpy(), self.indices.cpu().numpy(), self.indptr.cpu().numpy()),
                             shape=self.shape)

    @staticmethod
    def vstack(blocks):
        """Stack CSR blocks with equal column counts (tensor plumbing: concatenation + offset)."""
        n_cols = blocks[0].shape[1]
        ptrs, off, rows = [blocks[0].indptr[:1]], 0, 0
        for b in blocks:
            ptrs.append(b.indptr[1:] + off)
            off += b.nnz
            rows += b.shape[0]
        return DeviceCSR(torch.cat(ptrs).to(torch.int32).contiguous(), torch.cat([b.indices for b in blocks]).contiguous(),
                         torch.cat([b.data for b in blocks]).contiguous(), (rows, n_cols))


class lin_op(object):
    def __init__(self, grid=None, row_0=0, col_N=None, col_0=None, name=None):
        if col_0 is not None:
            self.col_0 = col_0
        elif grid is not None:
            self.col_0 = grid.col_0
        self.col_N = None
        if col_N is not None:
            self.col_N = col_N
        elif grid is not None:
            self.col_N = grid.col_N
        self.row_0 = row_0
        self.N_eq = 0
        self.name = name
        self.id = None
        self._host = {'r': np.array([], dtype=int), 'c': np.array([], dtype=int), 'v': np.array([], dtype=float)}
        self._dev = {'r': None, 'c': None, 'v': None}
        self._csr_cache = {}
        self._blocks = None       # row-regular form [(cols[n, k], vals[n, k]), ...] when every row holds the same entries block by block
        self._ind0 = np.zeros([0], dtype=int)      # or a list of parts: ('arange', n) / ndarray, joined on demand
        self.TOC = {'rows': dict(), 'cols': dict()}
        self.grid = grid
        self.dst_grid = None
        self.dst_ind0 = None
        self.expected = None
        self.prior = None
        self.shape = None
        self.size = None

    # ------------------------------------------------------------------ ind0 (lazy: interpolation ops index 1e7 points)
    @property
    def ind0(self):
        if isinstance(self._ind0, list):
            def part(p):
                if isinstance(p, tuple) and p[0] == 'arange':
                    return np.arange(0, p[1], dtype='int')
                if isinstance(p, tuple) and p[0] == 'dev':        # centre nodes still on the device (diff_op)
                    return p[1].cpu().numpy().astype(int)
                return np.asarray(p)
            self._ind0 = np.concatenate([part(p) for p in self._ind0]) if self._ind0 else np.zeros([0], dtype=int)
        return self._ind0

    @ind0.setter
    def ind0(self, value):
        self._ind0 = value

    def _ind0_parts(self):
        return list(self._ind0) if isinstance(self._ind0, list) else [self._ind0]

    # ------------------------------------------------------------------ triplet storage
    def _get_host(self, f):
        if self._host[f] is None:
            t = self._dev[f]
            self._host[f] = t.cpu().numpy().astype(_HOST_DTYPE[f])
            self._dev[f] = None            # the host copy may now be edited in place: it is authoritative
            self._csr_cache = {}
            self._blocks = None
        return self._host[f]

    def _set_host(self, f, value):
        self._host[f] = value
        self._dev[f] = None
        self._csr_cache = {}
        self._blocks = None

    r = property(lambda self: self._get_host('r'), lambda self, v: self._set_host('r', v))
    c = property(lambda self: self._get_host('c'), lambda self, v: self._set_host('c', v))
    v = property(lambda self: self._get_host('v'), lambda self, v: self._set_host('v', v))

    def _device(self, f):
        """Device tensor of a triplet field (uploads the host copy when that one is authoritative)."""
        if self._dev[f] is not None:
            return self._dev[f]
        h = self._host[f]
        if h is None:
            return torch.empty(0, dtype=_DEV_DTYPE[f], device=_cabi.device())
        return _cabi.to_dev(np.asarray(h), _DEV_DTYPE[f])

    def _set_device(self, r, c, v, blocks=None):
        self._dev = {'r': r, 'c': c, 'v': v}
        self._host = {'r': None, 'c': None, 'v': None}
        self._csr_cache = {}
        self._blocks = blocks

    def _on_device(self):
        return all(self._dev[f] is not None for f in _FIELDS)

    def __update_size_and_shape__(self):
        self.shape = (self.N_eq, self.col_N)

    _update_shape = __update_size_and_shape__

    # ------------------------------------------------------------------ transforms
    def apply_xform(self, pts, xform=None, dims=[0, 1]):
        """Affine change of coordinates before interpolation (lin_op.py:45-72); host side, tiny."""
        if xform is None:
            xform = self.grid.xform
        if xform is None:
            return pts
        nd = len(xform['origin'])
        local = (np.c_[[np.asarray(p).ravel() for p in pts[0:nd]]].T - xform['origin']) @ xform['basis_vectors']
        return [local[:, d].ravel() for d in dims]

    def ravel(self):
        for f in _FIELDS:
            if self._dev[f] is not None:
                self._dev[f] = self._dev[f].reshape(-1)
            else:
                self._host[f] = np.asarray(self._host[f]).ravel()
        return self

    def fix_dtypes(self):
        for f in ('r', 'c'):
            if self._host[f] is not None:
                self._host[f] = np.asarray(self._host[f]).astype(int)

    # ------------------------------------------------------------------ builders (CUDA kernels)
    def interp_mtx(self, pts_in, xform=None, dims=None, bounds_error=True):
        """Bi/tri-linear interpolation rows for a set of points (lin_op.py:163-247).

        Device path: ``lss_interp_coo`` -- one thread per point emits its 2^d (row, col, weight)
        triplets with NumPy-identical rounding.  ``pts_in`` may hold NumPy arrays or CUDA tensors.
        """
        if xform is None and self.grid.xform is not None:
            xform = self.grid.xform
        if dims is None:
            dims = np.arange(self.grid.N_dims, dtype=int)
        if xform is not None:
            pts = self.apply_xform(pts_in, xform, dims)
        else:
            pts = [p.reshape(-1) if isinstance(p, torch.Tensor) else np.asarray(p).ravel() for p in pts_in]
        pts = [_cabi.to_dev(p, torch.float64) for p in pts[:self.grid.N_dims]]
        n_in = int(pts[0].numel())
        rows, cols, vals, valid, n_bad = _cabi.interp_coo(self.grid.as_struct(), pts)
        if n_bad > 0:
            if bounds_error:
                raise ValueError(f'Found {n_bad} points out of bounds for grid {self.grid.name}')
            keep = valid.bool()
            rows, cols, vals = rows[keep].contiguous(), cols[keep].contiguous(), vals[keep].contiguous()
        n_pts = n_in - n_bad
        self._set_device(rows, cols, vals, blocks=[(cols, vals)] if n_bad == 0 and cols.dim() == 2 else None)
        self._valid_pts = valid if n_bad > 0 else None
        self.N_eq = n_pts
        self._ind0 = [('arange', n_pts)]                       # np.arange(n_pts), materialised on access
        self.TOC['rows'] = {self.name: range(self.N_eq)}
        self.TOC['cols'] = {self.grid.name: np.arange(self.grid.col_0, self.grid.col_0 + self.grid.N_nodes)}
        self._update_shape()
        return self

    def diff_op(self, delta_subs, vals, which_nodes=None, valid_equations_only=True):
        """Rows that apply a stencil (offset template + coefficients) at every node where it fits
        (lin_op.py:80-132).  Device path: ``lss_stencil_coo``, one thread per equation."""
        offs = np.array([np.asarray(d).ravel() for d in delta_subs], dtype=np.int64)
        coef = np.array([np.asarray(vals[k]).ravel()[0] for k in range(offs.shape[1])], dtype=float)
        rows, cols, v, ind0 = _cabi.stencil_coo(self.grid.as_struct(), offs, coef, self.row_0, bool(valid_equations_only))
        if which_nodes is None:
            # the centre nodes stay on the device until somebody reads .ind0 (no 16 MB copy and sync per stencil)
            self._set_device(rows, cols, v)
            self.N_eq = int(ind0.numel())
            self._ind0 = [('dev', ind0)]
            self.TOC['rows'] = {self.name: range(self.N_eq)}
            self.TOC['cols'] = {self.grid.name: np.arange(self.grid.col_0, self.grid.col_0 + self.grid.N_nodes)}
            self._update_shape()
            return self
        ind0 = ind0.cpu().numpy()
        if which_nodes is not None:
            # keep only centres listed in which_nodes (lin_op.py:104-106), rows renumbered consecutively
            sel = np.flatnonzero(np.isin(ind0, which_nodes))
            sel_d = _cabi.to_dev(sel, torch.int64)
            cols, v = cols[sel_d].contiguous(), v[sel_d].contiguous()
            rows = (self.row_0 + torch.arange(sel.size, device=cols.device, dtype=torch.int32))[:, None] \
                .expand(-1, cols.shape[1]).contiguous()
            ind0 = ind0[sel]
        self._set_device(rows, cols, v)
        self.N_eq = int(ind0.size)
        self.ind0 = ind0
        self.TOC['rows'] = {self.name: range(self.N_eq)}
        self.TOC['cols'] = {self.grid.name: np.arange(self.grid.col_0, self.grid.col_0 + self.grid.N_nodes)}
        self._update_shape()
        return self

    def data_bias(self, ind=None, val=None, col=None, DOF=None, _toc=None, _regular=False):
        """One entry per listed row: picks (or adds) a single model parameter (lin_op.py:504-522)."""
        if col is None:
            col = self.col_N
            self.col_N += 1
        dev = _cabi.device()
        ind_d = _cabi.to_dev(ind, torch.int32) if not isinstance(ind, torch.Tensor) else ind.to(torch.int32)
        if np.isscalar(col) or (isinstance(col, np.ndarray) and col.ndim == 0):
            col_d = torch.full_like(ind_d, int(col))
        else:
            col_d = _cabi.to_dev(col, torch.int32) if not isinstance(col, torch.Tensor) else col.to(torch.int32)
        if val is None:
            val_d = torch.ones(ind_d.shape, dtype=torch.float64, device=dev)
        else:
            val_d = _cabi.to_dev(np.asarray(val).ravel(), torch.float64) if not isinstance(val, torch.Tensor) else val
        col_d, val_d = col_d.contiguous(), val_d.contiguous()
        # _regular: the caller guarantees ind == arange(n): one entry per row, in row order
        self._set_device(ind_d.contiguous(), col_d, val_d, blocks=[(col_d.reshape(-1, 1), val_d.reshape(-1, 1))] if _regular else None)
        if _toc is not None:       # caller already knows the distinct rows / columns (avoids sorting 1e7 indices)
            self.TOC['rows'], self.TOC['cols'] = {self.name: _toc[0]}, {self.name: _toc[1]}
        else:
            self.TOC['rows'] = {self.name: torch.unique(ind_d).cpu().numpy().astype(int)}
            self.TOC['cols'] = {self.name: torch.unique(col_d).cpu().numpy().astype(int)}
        self.N_eq = int(ind_d.max().item()) + 1 if ind_d.numel() else 0
        self._update_shape()
        return self

    # ------------------------------------------------------------------ stencil families
    def _family(self, name, offsets, coeffs):
        return lin_op(self.grid, name=name).diff_op(offsets, coeffs)

    def one(self, DOF='z', which_nodes=None):
        self.diff_op([[0]] * len(self.grid.shape), np.array([1.]), which_nodes=which_nodes)
        return self

    def grad(self, DOF='z'):
        """First differences in x and y (lin_op.py:255-261); both use delta[0]."""
        k = np.array([-1., 1.]) / (self.grid.delta[0])
        self.vstack((self._family('d' + DOF + '_dx', ([0, 0], [-1, 0]), k),
                     self._family('d' + DOF + '_dy', ([-1, 0], [0, 0]), k)))
        return self

    def grad_dzdt(self, DOF='z', t_lag=1):
        """Space-time mixed first differences (lin_op.py:263-269)."""
        k = np.array([-1., 1., 1., -1.]) / (t_lag * self.grid.delta[0] * self.grid.delta[2])
        lag = [-t_lag, -t_lag, 0, 0]
        self.vstack((self._family('d2' + DOF + '_dxdt', ([0, 0, 0, 0], [-1, 0, -1, 0], lag), k),
                     self._family('d2' + DOF + '_dydt', ([-1, 0, -1, 0], [0, 0, 0, 0], lag), k)))
        return self

    def diff(self, lag=1, dim=0):
        k = np.array([-1., 1.]) / (lag * self.grid.delta[dim])
        offs = [[0, 0] for _ in range(self.grid.N_dims)]
        offs[dim] = [0, lag]
        self.diff_op(offs, k)
        return self

    def dzdt(self, lag=1, DOF='dz'):
        k = np.array([-1., 1.]) / (lag * self.grid.delta[2])
        self.diff_op(([0, 0], [0, 0], [0, lag]), k)
        self.update_dst_grid([0, 0, 0.5 * lag * self.grid.delta[2]], np.array([1, 1, 1]))
        return self

    def d2z_dt2(self, DOF='dz', t_lag=1):
        """Second difference in time (lin_op.py:286-290).  Like the reference, returns a NEW operator
        named 'd2<DOF>_dt2'."""
        k = np.array([-1, 2, -1]) / ((t_lag * self.grid.delta[2]) ** 2)
        return lin_op(self.grid, name='d2' + DOF + '_dt2').diff_op(([0, 0, 0], [0, 0, 0], [-t_lag, 0, t_lag]), k)

    def grad2(self, DOF='z'):
        """Second differences xx, yy, xy (lin_op.py:292-299)."""
        k = np.array([-1., 2., -1.]) / (self.grid.delta[0] ** 2)
        kxy = 0.5 * np.array([-1., 1., 1., -1]) / (self.grid.delta[0] ** 2)
        self.vstack((self._family('d2' + DOF + '_dx2', ([0, 0, 0], [-1, 0, 1]), k),
                     self._family('d2' + DOF + '_dy2', ([-1, 0, 1], [0, 0, 0]), k),
                     self._family('d2' + DOF + '_dxdy', ([-1, -1, 1, 1], [-1, 1, -1, 1]), kxy)))
        return self

    def grad2_dzdt(self, DOF='z', t_lag=1):
        """Time differences of the second spatial differences (lin_op.py:301-309)."""
        k = np.array([-1., 2., -1., 1., -2., 1.]) / (t_lag * self.grid.delta[0] ** 2. * self.grid.delta[2])
        lag6 = [-t_lag, -t_lag, -t_lag, 0, 0, 0]
        kxy = np.array([-1., 1., 1., -1., 1., -1., -1., 1.]) / (self.grid.delta[0] ** 2 * self.grid.delta[2])
        lag8 = [-t_lag, -t_lag, -t_lag, -t_lag, 0, 0, 0, 0]
        self.vstack((self._family('d3' + DOF + '_dx2dt', ([0, 0, 0, 0, 0, 0], [-1, 0, 1, -1, 0, 1], lag6), k),
                     self._family('d3' + DOF + '_dy2dt', ([-1, 0, 1, -1, 0, 1], [0, 0, 0, 0, 0, 0], lag6), k),
                     self._family('d3' + DOF + '_dxdydt', ([-1, 0, -1, 0, -1, 0, -1, 0], [-1, -1, 0, 0, -1, -1, 0, 0], lag8), kxy)))
        return self

    # ------------------------------------------------------------------ combinators
    def add(self, op):
        """Append another operator's triplets on the same rows (lin_op.py:134-161).  Like the
        reference this neither updates N_eq nor the row TOC."""
        if isinstance(op, (list, tuple)):
            for this_op in op:
                self.add(this_op)
            return self
        if self._on_device() and op._on_device():
            blocks = None
            if self._blocks and op._blocks and self._blocks[0][0].shape[0] == op._blocks[0][0].shape[0]:
                blocks = self._blocks + op._blocks
            self._set_device(*[torch.cat((self._dev[f].reshape(-1), op._dev[f].reshape(-1))) for f in _FIELDS], blocks=blocks)
        else:
            for f in _FIELDS:
                self._set_host(f, np.append(getattr(self, f), getattr(op, f)))
        self._ind0 = self._ind0_parts() + op._ind0_parts()
        for key in op.TOC['cols'].keys():
            self.TOC['cols'][key] = op.TOC['cols'][key]
        self.col_N = np.maximum(self.col_N, op.col_N)
        self._update_shape()
        return self

    def vstack(self, ops, order=None, name=None, TOC_cols=None):
        """Stack operators vertically, keeping the table of contents (lin_op.py:555-631)."""
        if isinstance(ops, lin_op):
            ops = (self, ops)
        if order is None:
            order = range(len(ops))
        if name is not None:
            self.name = name
        if TOC_cols is None:
            TOC_cols = dict()
            all_cols = []
            for op in ops:
                for key, val in op.TOC['cols'].items():
                    TOC_cols[key] = val
                    all_cols.append(np.asarray(val).ravel())
            if self.name is not None:
                TOC_cols[self.name] = _union_of_columns(all_cols)
        if self.col_N is None:
            self.col_N = np.max(np.array([op.col_N for op in ops]))
        self.TOC['cols'] = TOC_cols
        on_dev = all(ops[i]._on_device() for i in order)
        parts = {f: [] for f in _FIELDS}
        expected = []
        next_row = 0
        for i in order:
            op = ops[i]
            if on_dev:
                parts['r'].append(op._dev['r'].reshape(-1) + next_row)
                parts['c'].append(op._dev['c'].reshape(-1))
                parts['v'].append(op._dev['v'].reshape(-1))
            else:
                parts['r'].append(np.asarray(op.r).ravel() + next_row)
                parts['c'].append(np.asarray(op.c).ravel())
                parts['v'].append(np.asarray(op.v).ravel())
            if op.expected is not None:
                expected.append(np.asarray(op.expected).ravel())
            label = op.name if op.name is not None else 'eq'
            # a label already present gets a numeric suffix (lin_op.py:599-607)
            k, candidate = 0, label
            while candidate in self.TOC['rows']:
                k += 1
                candidate = '%s_%d' % (label, k)
            label = candidate
            shifted = []
            for key, rows in op.TOC['rows'].items():
                # contiguous row lists stay `range` objects (no 8 bytes per row on the host for 2e6 constraint rows);
                # index arrays with them through rows_of()
                if isinstance(rows, range) and rows.step == 1:
                    these = range(rows.start + next_row, rows.stop + next_row)
                else:
                    these = _index_array(rows) + next_row
                self.TOC['rows'][key] = these
                shifted.append(these)
            if label not in self.TOC['rows']:
                self.TOC['rows'][label] = _join_rows(shifted)
            next_row += op.N_eq
        self.N_eq = next_row
        if on_dev:
            self._set_device(*[torch.cat(parts[f]) if parts[f] else torch.empty(0, dtype=_DEV_DTYPE[f], device=_cabi.device())
                               for f in _FIELDS])
        else:
            for f in _FIELDS:
                self._set_host(f, np.concatenate(parts[f]))
        if expected:
            self.expected = np.concatenate(expected)
        self._ind0 = [p for op in ops for p in op._ind0_parts()]        # joined on demand
        if self.name is not None and len(self.name) > 0:
            self.TOC['rows'][self.name] = range(0, next_row)
        self._update_shape()
        return self

    # ------------------------------------------------------------------ masks / scaling
    def mask_for_ind0(self, mask_scale=None, mask=None):
        """Sample the grid mask (or a scaling map) at each row's centre node (lin_op.py:633-667)."""
        if mask is None:
            mask = self.grid.mask
        if mask is None:
            return np.ones_like(self.ind0, dtype=float)
        if mask.dtype == bool and mask.all():                      # default all-True mask: nothing to sample
            if mask_scale is None:
                return np.ones(int(self.N_eq), dtype=bool)
            val = 0.0
            for key, scale in mask_scale.items():                  # same outcome as the sampling loop below: last matching key wins
                if key == True:                                    # noqa: E712  (1 and True both select the unmasked nodes)
                    val = scale
            return np.full(int(self.N_eq), val, dtype=float)
        local = self.ind0 - self.grid.col_0
        if len(self.grid.shape) > len(mask.shape):
            full = np.unravel_index(local, self.grid.shape)
            subs = tuple(full[d] for d in range(len(mask.shape)))
        else:
            subs = np.unravel_index(local, mask.shape)
        sampled = mask[subs]
        if mask_scale is None:
            return sampled
        out = np.zeros_like(sampled, dtype=float)
        for key, scale in mask_scale.items():
            out[sampled == key] = scale
        return out

    def sum_to_grid3(self, kernel_size, sub0s=None, lag=None, taper=True, valid_equations_only=True, dims=None):
        """Rows that add up the values around output grid points, optionally as a lagged time difference
        (lin_op.py:404-488): a box (or edge-tapered box) template handed to ``diff_op`` at the centre nodes ``sub0s``;
        the destination grid is then the coarse grid of those centres."""
        kernel_size = np.asarray(kernel_size)
        half_kernel = np.floor((kernel_size - 1) / 2).astype(int) if taper else np.floor(kernel_size / 2).astype(int)
        if dims is None:
            dims = range(len(self.grid.shape))
        n_dims = len(dims)
        if sub0s is None:
            if taper:
                if valid_equations_only:
                    sub0s = np.meshgrid(*[np.arange(half_kernel[ii] + 1, self.grid.shape[ii], kernel_size[ii] - 1, dtype=int)
                                          for ii in dims], indexing='ij')
                else:
                    sub0s = np.meshgrid(*[np.arange(0, self.grid.shape[ii] + 1, kernel_size[ii] - 1, dtype=int) for ii in dims],
                                        indexing='ij')
            else:
                sub0s = np.meshgrid(*[np.arange(half_kernel[ii], self.grid.shape[ii], kernel_size[ii], dtype=int) for ii in dims],
                                    indexing='ij')
        ind0 = self.grid.global_ind(sub0s[0:n_dims])
        if np.mod(kernel_size[0] / 2, 1) == 0:               # even kernel: centres fall between nodes
            di, dj = np.meshgrid(np.arange(-half_kernel[0], half_kernel[0]), np.arange(-half_kernel[1], half_kernel[1]),
                                 indexing='ij')
            grid_shift = [-self.grid.delta[0] / 2, -self.grid.delta[1] / 2, 0][0:n_dims]
        else:
            di, dj = np.meshgrid(np.arange(-half_kernel[0], half_kernel[0] + 1), np.arange(-half_kernel[1], half_kernel[1] + 1),
                                 indexing='ij')
            grid_shift = [0, 0, 0][0:n_dims]
        wt0 = np.ones(kernel_size[0:2], dtype=float)
        if taper:
            for ii in [0, -1]:
                wt0[ii, :] /= 2
                wt0[:, ii] /= 2
        wt0 = wt0.ravel()
        if lag is None:
            delta_subs = [di.ravel(), dj.ravel(), np.zeros_like(di.ravel())]
            wt = wt0
            grid_shift = [0, 0, 0]
        else:
            delta_subs = [np.concatenate([di.ravel(), di.ravel()]), np.concatenate([dj.ravel(), dj.ravel()]),
                          np.concatenate([np.zeros_like(di.ravel(), dtype=int), np.zeros_like(di.ravel(), dtype=int) + lag])]
            wt = np.concatenate([-wt0, wt0]) / (lag * self.grid.delta[2])
            grid_shift[2] = 0.5 * lag * self.grid.delta[2]
        n_grid_dims = len(self.grid.shape)
        self.diff_op(delta_subs[:n_grid_dims], wt.astype(float), which_nodes=ind0, valid_equations_only=valid_equations_only)
        kernel_out = (np.maximum(1, kernel_size - 1) if taper else kernel_size)[:n_grid_dims]
        self.update_dst_grid(grid_shift[:n_grid_dims], kernel_out)
        return self

    def apply_mask(self, mask=None, row_N=None, time_step_overlap=1):
        """Multiply every entry by the mask value at its column's grid node (lin_op.py:693-732).  The reference walks
        the CSR rows in Python; here the canonical CSR is formed on the device and its values are scaled by one gather.
        Like the reference the operator is left as the flat triplets of that CSR (explicit zeros where the mask is 0)."""
        if time_step_overlap > 1:
            raise NotImplementedError('time_step_overlap > 1 belongs to the 3-D mask path (out of scope)')
        if mask is None:
            mask = self.grid.mask
        return self._scale_by_mask(np.asarray(mask), self._csr_device(row_N=row_N, col_N=self.col_N))

    def apply_2d_mask(self, mask=None):
        """Same with a mask over the first two grid dimensions (lin_op.py:669-691)."""
        if mask is None:
            mask = self.grid.mask
        mask = np.asarray(mask)
        if mask.ndim != 2:
            raise ValueError('apply_2d_mask needs a 2-D mask')
        return self._scale_by_mask(mask, self._csr_device())

    def _scale_by_mask(self, mask, csr):
        shape = [int(n) for n in self.grid.shape]
        if list(mask.shape) != shape[:mask.ndim]:
            raise ValueError('mask shape %s does not match grid shape %s' % (mask.shape, tuple(shape)))
        trailing = int(np.prod(shape[mask.ndim:])) if mask.ndim < len(shape) else 1
        local = csr.indices.long() - int(self.grid.col_0)
        m_d = _cabi.to_dev(np.ascontiguousarray(mask, dtype=float).ravel(), torch.float64)
        counts = (csr.indptr[1:] - csr.indptr[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(csr.shape[0], device=counts.device, dtype=torch.int32), counts)
        vals = csr.data * m_d[torch.div(local, trailing, rounding_mode='floor')]
        self._set_device(rows.contiguous(), csr.indices.contiguous(), vals.contiguous())
        return self

    def normalize_by_unit_product(self, wt=1):
        """Scale each row so that sum(|v|) == wt (lin_op.py:311-321)."""
        unit = lin_op(col_N=self.col_N)
        unit.N_eq = self.N_eq
        unit._set_device(self._device('r').reshape(-1), self._device('c').reshape(-1), self._device('v').reshape(-1).abs())
        unit._update_shape()
        csr = unit._csr_device(row_N=int(unit.N_eq))
        norm = csr.dot(torch.ones(csr.shape[1], dtype=torch.float64, device=_cabi.device()))
        scale = torch.where(norm > 0, 1. / norm, torch.zeros_like(norm))
        r = self._device('r').reshape(-1).long()
        new_v = self._device('v').reshape(-1) * scale[r] * wt
        shape = tuple(self._dev['v'].shape) if self._dev['v'] is not None else np.shape(self._host['v'])
        self._set_device(self._device('r'), self._device('c'), new_v.reshape(shape))

    def mean_of_mask(self, mask, dzdt_lag=None):
        """Weighted mean over the nodes a 2-D mask selects, one row per epoch, or the lagged rate of that mean with
        ``dzdt_lag`` (lin_op.py:347-402).  ``mask``: anything with ``interp(x, y)`` returning its values at node
        coordinates (the reference's pointCollection grid), or an array already on the (y, x) nodes of the grid.
        Weights = mask x cell area, normalised; the destination grid is a single point at the weighted centroid."""
        g = self.grid
        if hasattr(mask, 'interp'):
            yy, xx = np.meshgrid(g.ctrs[0], g.ctrs[1], indexing='ij')
            w = np.array(mask.interp(xx, yy), dtype=float)
        else:
            w = np.array(mask, dtype=float)
        if w.shape != tuple(g.shape[0:2]):
            raise ValueError('mean_of_mask: mask values must sit on the (y, x) nodes of the grid')
        w[~np.isfinite(w)] = 0
        i0, j0 = np.nonzero(w)
        sel = np.flatnonzero(w)
        v0 = w.ravel()[sel] if g.cell_area is None else (w * g.cell_area).ravel()[sel]
        v0 = v0 / v0.sum()
        y0 = np.sum((g.bds[0][0] + i0 * g.delta[0]) * v0)
        x0 = np.sum((g.bds[1][0] + j0 * g.delta[1]) * v0)
        if len(g.shape) < 3:
            # one row: the average of the grid itself
            self._set_host('r', np.zeros_like(i0))
            self.c = g.global_ind([i0, j0])
            self.v = v0
            self.N_eq = 1
            self.col_N = int(np.max(self.c)) + 1
            self.__update_size_and_shape__()
            self.dst_grid = fd_grid([[y0, y0], [x0, x0]], g.delta, 0, col_N=0, srs_proj4=g.srs_proj4)
            self.dst_ind0 = np.array([0]).astype(int)
            return self
        nt = int(g.shape[2])
        if dzdt_lag is None:
            n_out = nt
            terms = [(0, v0)]
            t_vals = g.ctrs[2]
        else:
            n_out = nt - dzdt_lag
            terms = [(0, -v0 / dzdt_lag / g.delta[2]), (dzdt_lag, v0 / dzdt_lag / g.delta[2])]
            t_vals = g.ctrs[-1][:-dzdt_lag] + g.delta[-1] * dzdt_lag / 2
        rr, cc, vv = [], [], []
        for ii in range(n_out):
            for shift, vals in terms:
                rr.append(np.zeros_like(i0) + ii)
                cc.append(g.global_ind([i0, j0, np.zeros_like(i0) + ii + shift]))
                vv.append(vals)
        self._set_host('r', np.concatenate(rr))
        self.c, self.v = np.concatenate(cc), np.concatenate(vv)
        self.N_eq = int(self.r.max()) + 1
        self.dst_grid = fd_grid([[y0, y0], [x0, x0], [t_vals[0], t_vals[-1]]], g.delta, 0, col_N=self.N_eq, srs_proj4=g.srs_proj4)
        self.__update_size_and_shape__()
        self.dst_ind0 = np.arange(self.N_eq, dtype=int)
        return self

    def update_dst_grid(self, grid_shift, kernel_size):
        """Destination grid of an averaging / differencing operator (lin_op.py:490-502)."""
        sub0 = np.unravel_index(self.ind0 - self.grid.col_0, self.grid.shape)
        dims = range(len(self.grid.shape))
        self.dst_grid = fd_grid([[self.grid.ctrs[d][sub0[d][j]] + grid_shift[d] for j in [0, -1]] for d in dims],
                                kernel_size * self.grid.delta, name=self.name)
        out_subs = [((sub0[d] - sub0[d][0]) / kernel_size[d]).astype(int) for d in dims]
        self.dst_ind0 = np.ravel_multi_index(out_subs, self.dst_grid.shape)
        return self

    def grid_prod(self, m, grid=None):
        """Operator times model vector, mapped onto the destination grid (lin_op.py:524-538)."""
        if grid is None:
            grid = self.grid if self.dst_grid is None else self.dst_grid
        ind0 = self.ind0 if self.dst_ind0 is None else self.dst_ind0
        P = np.zeros(grid.col_N + 1) + np.nan
        csr = self._csr_device(row_N=ind0.size, col_N=np.size(m))
        y = csr.dot(_cabi.to_dev(np.asarray(m, dtype=float).ravel(), torch.float64))
        P[ind0] = y.cpu().numpy().ravel()
        return P[grid.col_0:grid.col_N].reshape(grid.shape)

    def print_TOC(self):
        for rc in ('cols', 'rows'):
            print(rc)
            first = {k: np.min(self.TOC[rc][k]) for k in self.TOC[rc].keys()}
            for key in sorted(first, key=first.get):
                print("\t%s\t%d : %d" % (key, np.min(self.TOC[rc][key]), np.max(self.TOC[rc][key])))

    # ------------------------------------------------------------------ export
    def _csr_device(self, col_N=None, row_N=None):
        """Canonical CSR of the operator on the GPU (zero entries dropped, duplicates summed)."""
        if col_N is None:
            col_N = self.col_N
        r, c, v = (self._device(f).reshape(-1) for f in _FIELDS)
        cacheable = self._on_device()
        if row_N is None:
            key0 = ('extent',)
            if cacheable and key0 in self._csr_cache:
                mx = self._csr_cache[key0]
            else:
                mx, _ = _cabi.coo_extent(r, v)
                if cacheable:
                    self._csr_cache[key0] = mx
            if mx < 0:
                raise ValueError('zero-size array to reduction operation maximum which has no identity')
            row_N = mx + 1                                   # lin_op.py:751-752
        key = (int(row_N), int(col_N))
        if cacheable and key in self._csr_cache:
            return self._csr_cache[key]
        res = None
        if cacheable and self._blocks and len(self._blocks) <= 8 and int(self._blocks[0][0].shape[0]) == int(row_N):
            res = _cabi.blocks_to_csr(self._blocks, int(row_N), int(col_N))      # row-regular blocks: no sort
        if res is None:
            res = _cabi.coo_to_csr(r, c, v, int(row_N), int(col_N))
        indptr, indices, data = res
        out = DeviceCSR(indptr, indices, data, key)
        if cacheable:
            self._csr_cache[key] = out
        return out

    def toCSR(self, col_N=None, row_N=None):
        """scipy.sparse.csr_matrix of the operator (lin_op.py:745-753): built on the GPU, copied out."""
        self.fix_dtypes()
        return self._csr_device(col_N=col_N, row_N=row_N).to_scipy()


class fd_op(lin_op):
    """Finite-difference template operator.

    The reference README (README.md:15) describes ``fd_op`` as the ``lin_op`` subclass that applies a
    stencil to every node of a grid; in the code that role is ``lin_op.diff_op`` (lin_op.py:80-132).
    This class packages it: ``fd_op(grid, delta_subs, vals, name=...)`` is
    ``lin_op(grid, name=name).diff_op(delta_subs, vals)``.
    """

    def __init__(self, grid, delta_subs=None, vals=None, name=None, row_0=0, which_nodes=None,
                 valid_equations_only=True):
        super().__init__(grid=grid, row_0=row_0, name=name)
        if delta_subs is not None:
            self.diff_op(delta_subs, vals, which_nodes=which_nodes, valid_equations_only=valid_equations_only)
