"""TEST DOUBLE of the C ABI for the CPU test tier (`-m "not gpu"`).

TEST INFRASTRUCTURE ONLY -- lives under tests/, is never imported by the package, and is only switched on by
the ``host_double`` fixture of tests/test_host_double.py.  It is NOT a CPU fallback of the product: without a GPU the
package raises (tests/test_cabi_symbols.py::test_no_cpu_fallback) and this file changes nothing about that.

Purpose: the host-side mirror of the reference interface (lin_op bookkeeping -- TOC, row offsets, ``expected`` /
``prior`` vectors, column ranges of bias / sensor grids, the edit loop's control flow, parse_model's reshaping) is
Python that cannot run in the build container because every kernel call needs a device.  The double replaces the
thin functional wrappers of ``lssurf_b200._cabi`` (and the solver handle) by the oracle's NumPy restatements on CPU
tensors, so the same host code can be driven end to end against the reference fixtures in tests/golden where there
is no GPU.  What it checks is therefore the HOST logic; the kernels themselves are only checked by the `-m gpu`
tier, through the real C ABI.
"""
import contextlib

import numpy as np
import torch

from oracle import lssurf_oracle as O


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


_I32, _I64, _F64, _U8 = torch.int32, torch.int64, torch.float64, torch.uint8


def _arg(t, dtype, what):
    """What the real wrappers demand of a pointer argument: `_cabi.ptr` asserts a contiguous tensor, and the C ABI reads
    the memory with a fixed element type (include/lssurf_b200.h) -- a tensor of another dtype would be reinterpreted
    silently on the GPU, so the double rejects it here."""
    if t is None:
        return None
    assert isinstance(t, torch.Tensor), f'{what}: the C ABI takes device tensors, got {type(t).__name__}'
    assert t.is_contiguous(), f'{what}: expected a contiguous tensor'
    assert t.dtype == dtype, f'{what}: the C ABI reads {dtype}, got {t.dtype}'
    return t


def _geom(g):
    nd = int(g.ndim)
    return dict(name='', ndim=nd, shape=np.array([g.shape[d] for d in range(nd)], dtype=np.int64),
                stride=np.array([g.stride[d] for d in range(nd)], dtype=np.int64),
                lo=np.array([g.lo[d] for d in range(nd)]), hi=np.array([g.hi[d] for d in range(nd)]),
                delta=np.array([g.delta[d] for d in range(nd)]), col_0=int(g.col_0))


def device():
    return torch.device('cpu')


def to_dev(a, dtype=None):
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def to_host(t):
    return t.detach().numpy().copy()


def interp_coo(g, pts, row_0=0):
    go = _geom(g)
    for p in pts[:go['ndim']]:
        _arg(p, _F64, 'interp_coo pts')
    P = [_np(p).astype(float) for p in pts[:go['ndim']]]
    n, nc = P[0].size, 1 << go['ndim']
    ok = O.points_in_grid(go, P)
    rows = np.repeat((row_0 + np.arange(n, dtype=np.int64))[:, None], nc, axis=1)
    cols = np.zeros((n, nc), dtype=np.int64)
    vals = np.zeros((n, nc))
    if ok.any():
        _, c, v = O.interp_triplets(go, [p[ok] for p in P])
        cols[ok], vals[ok] = c, v
    return (torch.from_numpy(rows.astype(np.int32)), torch.from_numpy(cols.astype(np.int32)), torch.from_numpy(vals),
            torch.from_numpy(ok.astype(np.uint8)), int(n - ok.sum()))


def stencil_coo(g, offsets, coeffs, row_0=0, valid_only=True):
    go = _geom(g)
    offs = np.asarray(offsets, dtype=np.int64).reshape(go['ndim'], -1)
    coef = np.asarray(coeffs, dtype=float).ravel()
    if valid_only:
        r, c, v, ind0 = O.stencil_triplets(go, list(offs), coef, row_0=row_0)
    else:
        # every node is a centre; offsets that leave the grid get column 0 and value 0 (lin_op.py:118-121)
        nd, shape, stride = go['ndim'], go['shape'], go['stride']
        ctr = [m.ravel() for m in np.meshgrid(*[np.arange(int(n)) for n in shape], indexing='ij')]
        n_eq, n_off = ctr[0].size, offs.shape[1]
        ind0 = go['col_0'] + sum(ctr[d] * int(stride[d]) for d in range(nd))
        r = np.repeat((row_0 + np.arange(n_eq, dtype=np.int64))[:, None], n_off, axis=1)
        c = np.zeros((n_eq, n_off), dtype=np.int64)
        v = np.zeros((n_eq, n_off))
        for k in range(n_off):
            sub = [ctr[d] + int(offs[d][k]) for d in range(nd)]
            ok = np.ones(n_eq, dtype=bool)
            for d in range(nd):
                ok &= (sub[d] >= 0) & (sub[d] < int(shape[d]))
            c[ok, k] = go['col_0'] + sum(sub[d][ok] * int(stride[d]) for d in range(nd))
            v[:, k] = coef[k] * ok
    return (torch.from_numpy(r.astype(np.int32)), torch.from_numpy(c.astype(np.int32)), torch.from_numpy(v),
            torch.from_numpy(ind0.astype(np.int64)))


def coo_extent(rows, vals):
    _arg(rows, _I32, 'coo_extent rows'), _arg(vals, _F64, 'coo_extent vals')
    r, v = _np(rows).ravel(), _np(vals).ravel()
    nz = v != 0
    return (int(r[nz].max()) if nz.any() else -1), int(nz.sum())


def coo_to_csr(rows, cols, vals, n_rows, n_cols):
    _arg(rows, _I32, 'coo_to_csr rows'), _arg(cols, _I32, 'coo_to_csr cols'), _arg(vals, _F64, 'coo_to_csr vals')
    assert rows.numel() == cols.numel() == vals.numel()
    indptr, indices, data, _ = O.coo_to_csr(_np(rows), _np(cols), _np(vals), n_rows=int(n_rows), n_cols=int(n_cols))
    return torch.from_numpy(indptr), torch.from_numpy(indices), torch.from_numpy(data)


def blocks_to_csr(blocks, n_rows, n_cols):
    """lss_blocks_to_csr: rows take their entries block after block, zeros dropped; None unless every row ascends."""
    for c, v in blocks:
        _arg(c, _I32, 'blocks_to_csr cols'), _arg(v, _F64, 'blocks_to_csr vals')
        assert c.shape == v.shape and c.dim() == 2 and c.shape[0] == n_rows
    cols = np.concatenate([_np(c) for c, _ in blocks], axis=1)
    vals = np.concatenate([_np(v) for _, v in blocks], axis=1)
    keep = vals != 0
    big = np.where(keep, cols, -1).astype(np.int64)
    run = np.maximum.accumulate(big, axis=1)
    prev = np.concatenate([np.full((cols.shape[0], 1), -1, dtype=np.int64), run[:, :-1]], axis=1)
    if np.any(keep & ((cols <= prev) | (cols >= n_cols))):
        return None
    indptr = np.concatenate([[0], np.cumsum(keep.sum(axis=1))]).astype(np.int32)
    return torch.from_numpy(indptr), torch.from_numpy(cols[keep].astype(np.int32)), torch.from_numpy(vals[keep].astype(float))


def csr_transpose(indptr, indices, data, n_rows, n_cols):
    _arg(indptr, _I32, 'csr_transpose indptr'), _arg(indices, _I32, 'csr_transpose indices'), _arg(data, _F64, 'csr_transpose data')
    T = O.csr_matrix(_np(indptr), _np(indices), _np(data), (n_rows, n_cols)).T.tocsr()
    T.sort_indices()
    return (torch.from_numpy(T.indptr.astype(np.int32)), torch.from_numpy(T.indices.astype(np.int32)),
            torch.from_numpy(T.data.astype(float)))


def dense_rank(x, nan_value=-9999.0, max_unique=1 << 20):
    _arg(x, _F64, 'dense_rank x')
    a = _np(x).astype(float)
    a = np.where(np.isnan(a), nan_value, a)
    u, inv = np.unique(a, return_inverse=True)
    return torch.from_numpy(inv.astype(np.int32)), u


def group_median(ids, vals, n_groups):
    _arg(ids, _I32, 'group_median ids'), _arg(vals, _F64, 'group_median vals')
    i, v = _np(ids).astype(np.int64), _np(vals).astype(float)
    out = np.full(int(n_groups), np.nan)
    for k in range(int(n_groups)):
        sel = v[i == k]
        if np.isfinite(sel).any():
            out[k] = np.nanmedian(sel)
    return torch.from_numpy(out)


def _spmv(indptr, indices, data, x, n_rows, mode, s):
    _arg(indptr, _I32, 'spmv indptr'), _arg(indices, _I32, 'spmv indices'), _arg(data, _F64, 'spmv data')
    _arg(x, _F64, 'spmv x'), _arg(s, _F64, 'spmv s')
    A = O.csr_matrix(_np(indptr), _np(indices), _np(data), (int(n_rows), int(_np(x).size)))
    y = A @ _np(x)
    if mode == 1:
        y = y * _np(s)
    elif mode == 2:
        y = _np(s) - y
    return y


def spmv(indptr, indices, data, x, n_rows, n_cols, mode=0, s=None, out=None):
    _arg(out, _F64, 'spmv out')
    y = torch.from_numpy(_spmv(indptr, indices, data, x, n_rows, mode, s))
    if out is not None:
        out.copy_(y)
        return out
    return y


class SpmvPlan(object):
    def __init__(self, indptr, n_rows):
        pass


def spmv_planned(plan, indptr, indices, data, x, n_rows, mode=0, s=None, out=None):
    return spmv(indptr, indices, data, x, n_rows, None, mode, s, out)


def csr_rmatvec(indptr, indices, data, u, n_rows, n_cols):
    _arg(indptr, _I32, 'rmatvec indptr'), _arg(indices, _I32, 'rmatvec indices'), _arg(data, _F64, 'rmatvec data'), _arg(u, _F64, 'rmatvec u')
    A = O.csr_matrix(_np(indptr), _np(indices), _np(data), (int(n_rows), int(n_cols)))
    return torch.from_numpy(A.T @ _np(u))


def _all_ranks(a):
    """Rows of every rank in rank order (what the library's histogram all-reduce makes the quantiles see)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return a
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, np.asarray(a))
    return np.concatenate(parts)


def rde(r, sigma=None, mask=None, s1=0.0, comm=None):
    _arg(r, _F64, 'rde r'), _arg(sigma, _F64, 'rde sigma'), _arg(mask, _U8, 'rde mask')
    x = _all_ranks(_np(r).astype(float))
    if sigma is not None:
        x = x / np.sqrt(s1 * s1 + _all_ranks(_np(sigma)) ** 2)
    if mask is not None:
        x = x[_all_ranks(_np(mask)).astype(bool)]
    return float(O.rde(x)), int(x.size)


def sigma_extra(r, sigma, mask, comm=None):
    _arg(r, _F64, 'sigma_extra r'), _arg(sigma, _F64, 'sigma_extra sigma'), _arg(mask, _U8, 'sigma_extra mask')
    r, sigma, mask = _all_ranks(_np(r)), _all_ranks(_np(sigma)), _all_ranks(_np(mask))
    m = _np(mask).astype(bool)
    if m.sum() < 10:
        return 0.0, 0
    vals = O.sigma_extra(_np(r), _np(sigma), m)
    return float(vals[np.flatnonzero(m)[0]]), 1


def edit_mask(r, sigma, sigma_extra_t, n_sigma=3.0, comm=None, out=None):
    _arg(r, _F64, 'edit_mask r'), _arg(sigma, _F64, 'edit_mask sigma'), _arg(sigma_extra_t, _F64, 'edit_mask sigma_extra')
    rr, sg, se = _np(r), _np(sigma), _np(sigma_extra_t)
    ratio = np.abs(rr / np.sqrt(sg ** 2 + se ** 2))
    m = ratio < n_sigma
    m_all, sg_all = _all_ranks(m), _all_ranks(sg)
    st = dict(count=int(m_all.sum()), min_sigma=float(sg_all[m_all].min()) if m_all.any() else np.inf,
              near_threshold=int(np.sum(np.abs(_all_ranks(ratio) - n_sigma) < 1e-6)))
    return torch.from_numpy(m.astype(np.uint8)), st


def spatial_order(fine, coarse, pts):
    raise NotImplementedError('host double: run with sort_points=False')


class NormalSolver(object):
    """Stand-in for the solver handle: same constructor / methods, direct CPU solve of the same weighted system."""

    def __init__(self, G, Cn, z, prior, sigma, Ec, fixed_cols, dist=None, fused=True, structure=None, G_local=None, stages=None):
        for A in (G, Cn) + (() if G_local is None else (G_local,)):
            _arg(A.indptr, _I32, 'solver indptr'), _arg(A.indices, _I32, 'solver indices'), _arg(A.data, _F64, 'solver data')
        for v, what in ((z, 'z'), (prior, 'prior'), (sigma, 'sigma'), (Ec, 'Ec')):
            _arg(v, _F64, 'solver ' + what)
        self.G = O.csr_matrix(_np(G.indptr), _np(G.indices), _np(G.data), G.shape)
        self.C = O.csr_matrix(_np(Cn.indptr), _np(Cn.indices), _np(Cn.data), Cn.shape)
        Gl = G if G_local is None else G_local              # replicated solve: residuals over this rank's rows only
        self.G_local = O.csr_matrix(_np(Gl.indptr), _np(Gl.indices), _np(Gl.data), Gl.shape)
        self.z, self.prior, self.sigma, self.Ec = _np(z), _np(prior), _np(sigma), _np(Ec)
        self.n_cols = G.shape[1]
        free = np.ones(self.n_cols, dtype=bool)
        free[np.asarray(fixed_cols, dtype=np.int64)] = False
        self.free = np.flatnonzero(free)
        self.structure = structure
        self.last = {}
        self.x = torch.zeros(self.n_cols, dtype=torch.float64)

    def set_rows(self, in_TSE, sigma_extra=None):
        _arg(in_TSE, _U8, 'data_weights mask'), _arg(sigma_extra, _F64, 'data_weights sigma_extra')
        self.rows = _np(in_TSE).astype(bool)
        self.E = self.sigma if sigma_extra is None else np.sqrt(self.sigma ** 2 + _np(sigma_extra) ** 2)

    def set_profile(self, on):
        pass

    def solve(self, rtol=1e-12, max_iter=20000, warm_start=True):
        import scipy.sparse as sp
        k = np.flatnonzero(self.rows)
        A = sp.vstack([sp.diags(1. / self.E[k]) @ self.G[k][:, self.free], sp.diags(1. / self.Ec) @ self.C[:, self.free]]).tocsr()
        b = np.concatenate([self.z[k] / self.E[k], self.prior / self.Ec])
        x = np.zeros(self.n_cols)
        x[self.free] = O.solve_least_squares(A, b)
        self.x = torch.from_numpy(x)
        self.last = dict(iterations=0, relres=0.0, ms=0.0, form='host_double')
        return self.x

    def residual(self, x, z, out=None):
        _arg(x, _F64, 'residual x'), _arg(z, _F64, 'residual z'), _arg(out, _F64, 'residual out')
        r = torch.from_numpy(_np(z) - self.G_local @ _np(x))
        if out is not None:
            out.copy_(r)
            return out
        return r

    def close(self):
        pass


class NormalMatrix(object):
    """Stand-in for lssurf_b200.solver.NormalMatrix (lss_normal_*): N = A^T diag(w) A with SciPy, the rows and columns of
    the fixed unknowns left empty like the device matrix."""

    def __init__(self, A, fixed_cols=None):
        _arg(A.indptr, _I32, 'normal indptr'), _arg(A.indices, _I32, 'normal indices'), _arg(A.data, _F64, 'normal data')
        self.A = O.csr_matrix(_np(A.indptr), _np(A.indices), _np(A.data), A.shape)
        self.n_cols = A.shape[1]
        self.keep = np.ones(self.n_cols)
        self.keep[np.asarray(fixed_cols if fixed_cols is not None else [], dtype=np.int64)] = 0.

    def update(self, w):
        import scipy.sparse as sp
        _arg(w, _F64, 'normal weights')
        K = sp.diags(self.keep)
        self.N = (K @ (self.A.T @ sp.diags(_np(w)) @ self.A) @ K).tocsr()
        self.N.eliminate_zeros()
        self.N.sort_indices()

    def export(self):
        return (torch.from_numpy(self.N.indptr.astype(np.int32)), torch.from_numpy(self.N.indices.astype(np.int32)),
                torch.from_numpy(self.N.data.astype(np.float64)))

    def close(self):
        pass


_CABI_NAMES = ('device', 'to_dev', 'to_host', 'interp_coo', 'stencil_coo', 'coo_extent', 'coo_to_csr', 'blocks_to_csr', 'csr_transpose',
               'dense_rank', 'group_median', 'spmv', 'SpmvPlan', 'spmv_planned', 'csr_rmatvec', 'rde', 'sigma_extra',
               'edit_mask', 'spatial_order')


@contextlib.contextmanager
def installed():
    """Swap the wrappers in for the duration of a test and put the real ones back afterwards."""
    import sys
    from lssurf_b200 import _cabi
    import lssurf_b200.smooth_fit          # noqa: F401  (the package attribute of that name is the function)
    sf = sys.modules['lssurf_b200.smooth_fit']
    me = sys.modules[__name__]
    saved = {n: getattr(_cabi, n) for n in _CABI_NAMES}
    saved_solver = sf.NormalSolver
    import lssurf_b200.propagate_errors as pe
    saved_normal = pe.NormalMatrix
    from lssurf_b200.parallel import DistContext
    saved_reduce = DistContext.allreduce_sum
    saved_array = torch.Tensor.__array__

    def no_implicit_numpy(self, *a, **k):  # a CUDA tensor cannot be handed to NumPy; make CPU tensors behave the same
        raise TypeError("host double: implicit tensor -> NumPy conversion (fails on a CUDA tensor); use _cabi.to_host / .cpu()")

    def allreduce_sum(self, t):            # the library's NCCL all-reduce, over the CPU process group
        import torch.distributed as dist
        if self.world > 1:
            dist.all_reduce(t)
        return t
    try:
        DistContext.allreduce_sum = allreduce_sum
        torch.Tensor.__array__ = no_implicit_numpy
        for n in _CABI_NAMES:
            setattr(_cabi, n, getattr(me, n))
        sf.NormalSolver = NormalSolver
        pe.NormalMatrix = NormalMatrix
        yield
    finally:
        for n, f in saved.items():
            setattr(_cabi, n, f)
        sf.NormalSolver = saved_solver
        pe.NormalMatrix = saved_normal
        DistContext.allreduce_sum = saved_reduce
        torch.Tensor.__array__ = saved_array
