"""CPU oracle for the LSsurf ``smooth_fit`` hot path (numpy / scipy restatement).

TEST INFRASTRUCTURE ONLY.  This module is the *checker* for the CUDA product in
``lssurf_b200``: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package never imports anything under
``oracle/``.

It restates, as plain functions over numpy arrays, what the reference computes on the path
``fd_grid -> lin_op.interp_mtx / diff_op -> toCSR -> smooth_fit assembly -> iterate_fit``
(SURVEY.md section 8(a)); every function cites the reference lines it follows (paths relative
to /root/reference).  Parity pinning: ``oracle/make_golden.py`` runs the UNMODIFIED reference
python (under the stub modules in ``oracle/ref_stubs``) on seeded inputs and checks this
module against it bit-for-bit (CSR structure, values, sigma, masks) and to 1e-9 (solutions);
the resulting fixtures live in ``tests/golden``.

Known unpinned boundary: the reference's least-squares solver is SuiteSparseQR through the
third-party ``sparseqr`` module (requirements.txt:7, unpinned, source not in the tree).  The
solver here (normal equations + SuperLU + iterative refinement, dense lstsq for small systems)
is a mathematical stand-in: parity at that boundary is "unpinned" (no reference test vector
exists); the answer is the unique least-squares solution, checked against dense QR.
"""
import math

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

# --------------------------------------------------------------------------------------------
# grid geometry  (LSsurf/fd_grid.py:61-80)
# --------------------------------------------------------------------------------------------


def grid_geometry(bounds, deltas, col_0=0, name=''):
    """Node counts, centres, closed bounds, C-order strides and column range of a regular grid.

    fd_grid.py:61 (shape by float->int truncation), :62 (centres), :63 (bounds reset to the
    first/last centre), :67 (strides), :68/:76 (column range).
    """
    deltas = np.asarray(deltas, dtype=float)
    shape = np.array([((b[1] - b[0]) / d) + 1 for b, d in zip(bounds, deltas)]).astype(np.int64)
    ctrs = [b[0] + d * np.arange(n) for b, d, n in zip(bounds, deltas, shape)]
    lo = np.array([c[0] for c in ctrs])
    hi = np.array([c[-1] for c in ctrs])
    stride = np.ones(len(shape), dtype=np.int64)
    for d in range(len(shape) - 2, -1, -1):
        stride[d] = stride[d + 1] * shape[d + 1]
    n_nodes = int(np.prod(shape))
    return dict(name=name, shape=shape, ctrs=ctrs, lo=lo, hi=hi, delta=deltas, stride=stride,
                ndim=len(shape), n_nodes=n_nodes, col_0=int(col_0), col_N=int(col_0) + n_nodes)


def points_in_grid(g, pts):
    """Closed-interval containment test, finite first coordinate (fd_grid.py:90-96)."""
    ok = np.isfinite(pts[0])
    for d in range(g['ndim']):
        ok &= (pts[d] >= g['lo'][d]) & (pts[d] <= g['hi'][d])
    return ok


# --------------------------------------------------------------------------------------------
# interpolation rows  (LSsurf/lin_op.py:163-247, fd_grid.py:111-145)
# --------------------------------------------------------------------------------------------

def interp_triplets(g, pts, bounds_error=True):
    """Bi/tri-linear interpolation triplets for points ``pts`` (list of 1-D arrays, (y,x[,t])).

    Returns (rows, cols, vals) each shaped [n_valid, 2**ndim]; rows hold the ORIGINAL point
    index (lin_op.py:201, 230).  Corner order: dimension 0 slowest (np.mgrid, lin_op.py:216-222);
    weight = 1.0 * prod_d (frac_d or 1-frac_d) in ascending d (lin_op.py:227-236);
    frac = (p-lo)/delta - floor((p-lo)/delta) (fd_grid.py:117, 127; lin_op.py:212).
    """
    nd = g['ndim']
    pts = [np.asarray(p, dtype=float).ravel() for p in pts[:nd]]
    keep = np.flatnonzero(points_in_grid(g, pts))
    if bounds_error and keep.size < pts[0].size:
        raise ValueError(f"Found {pts[0].size - keep.size} points out of bounds for grid {g['name']}")
    frac, base = [], np.full(keep.size, g['col_0'], dtype=np.int64)
    for d in range(nd):
        s = (pts[d][keep] - g['lo'][d]) / g['delta'][d]
        c = np.floor(s)
        frac.append(s - c)
        base += c.astype(np.int64) * g['stride'][d]
    nc = 1 << nd
    rows = np.repeat(keep[:, None], nc, axis=1)
    cols = np.empty((keep.size, nc), dtype=np.int64)
    vals = np.ones((keep.size, nc))
    for k in range(nc):
        off = 0
        for d in range(nd):
            bit = (k >> (nd - 1 - d)) & 1
            off += bit * int(g['stride'][d])
            vals[:, k] = vals[:, k] * (frac[d] if bit else (1.0 - frac[d]))
        cols[:, k] = base + off
    return rows, cols, vals


# --------------------------------------------------------------------------------------------
# finite-difference stencil rows  (LSsurf/lin_op.py:80-132, 249-309)
# --------------------------------------------------------------------------------------------

def stencil_triplets(g, offsets, coeffs, row_0=0):
    """One row per node where the whole stencil fits (``valid_equations_only``), C order.

    ``offsets``: per-dimension lists of integer offsets (lin_op.py:92-103, 111-117);
    returns rows, cols, vals [n_eq, n_off] and ind0 = global index of each centre (:128).
    """
    nd = g['ndim']
    offsets = [np.asarray(o, dtype=np.int64) for o in offsets]
    axes = [np.arange(max(0, -int(o.min())), min(int(n), int(n) - int(o.max())))
            for o, n in zip(offsets, g['shape'])]
    mesh = np.meshgrid(*axes, indexing='ij')
    ctr = [m.ravel() for m in mesh]
    n_eq = ctr[0].size
    n_off = offsets[0].size
    ind0 = np.full(n_eq, g['col_0'], dtype=np.int64)
    for d in range(nd):
        ind0 += ctr[d] * g['stride'][d]
    rows = np.repeat((row_0 + np.arange(n_eq, dtype=np.int64))[:, None], n_off, axis=1)
    cols = np.empty((n_eq, n_off), dtype=np.int64)
    vals = np.empty((n_eq, n_off))
    for k in range(n_off):
        cols[:, k] = ind0 + sum(int(offsets[d][k]) * int(g['stride'][d]) for d in range(nd))
        vals[:, k] = coeffs[k]
    return rows, cols, vals, ind0


def stencil_families(g, kind, t_lag=1):
    """(name, offsets, coefficients) of each family of a composite constraint operator.

    grad2 lin_op.py:292-299; grad :255-261; grad2_dzdt :301-309; grad_dzdt :263-269;
    d2z_dt2 :286-290; one :249-253.  ``DOF`` naming follows constraint_functions.py:42-78.
    """
    d0 = g['delta'][0]
    if kind == 'grad2_z0':
        c = np.array([-1., 2., -1.]) / (d0 ** 2)
        cxy = 0.5 * np.array([-1., 1., 1., -1]) / (d0 ** 2)
        return [('d2z0_dx2', ([0, 0, 0], [-1, 0, 1]), c),
                ('d2z0_dy2', ([-1, 0, 1], [0, 0, 0]), c),
                ('d2z0_dxdy', ([-1, -1, 1, 1], [-1, 1, -1, 1]), cxy)]
    if kind == 'grad_z0':
        c = np.array([-1., 1.]) / d0
        return [('dz0_dx', ([0, 0], [-1, 0]), c), ('dz0_dy', ([-1, 0], [0, 0]), c)]
    d2 = g['delta'][2]
    L = t_lag
    if kind == 'grad2_dzdt':
        c = np.array([-1., 2., -1., 1., -2., 1.]) / (L * d0 ** 2. * d2)
        cxy = np.array([-1., 1., 1., -1., 1., -1., -1., 1.]) / (d0 ** 2 * d2)
        return [('d3z_dx2dt', ([0] * 6, [-1, 0, 1, -1, 0, 1], [-L, -L, -L, 0, 0, 0]), c),
                ('d3z_dy2dt', ([-1, 0, 1, -1, 0, 1], [0] * 6, [-L, -L, -L, 0, 0, 0]), c),
                ('d3z_dxdydt', ([-1, 0, -1, 0, -1, 0, -1, 0], [-1, -1, 0, 0, -1, -1, 0, 0],
                                [-L, -L, -L, -L, 0, 0, 0, 0]), cxy)]
    if kind == 'grad_dzdt':
        c = np.array([-1., 1., 1., -1.]) / (L * d0 * d2)
        return [('d2z_dxdt', ([0] * 4, [-1, 0, -1, 0], [-L, -L, 0, 0]), c),
                ('d2z_dydt', ([-1, 0, -1, 0], [0] * 4, [-L, -L, 0, 0]), c)]
    if kind == 'd2z_dt2':
        c = np.array([-1, 2, -1]) / ((L * d2) ** 2)
        return [('d2z_dt2', ([0, 0, 0], [0, 0, 0], [-L, 0, L]), c)]
    if kind == 'one':
        return [('one', [[0]] * g['ndim'], np.array([1.]))]
    raise KeyError(kind)


# --------------------------------------------------------------------------------------------
# COO -> canonical CSR  (LSsurf/lin_op.py:745-753; scipy coo_tocsr + sum_duplicates)
# --------------------------------------------------------------------------------------------

def coo_to_csr(r, c, v, n_rows=None, n_cols=None):
    """Drop v==0, sort by (row, col), sum duplicates, build indptr (int32/int32/float64)."""
    r = np.asarray(r).ravel().astype(np.int64)
    c = np.asarray(c).ravel().astype(np.int64)
    v = np.asarray(v, dtype=float).ravel()
    good = v != 0
    r, c, v = r[good], c[good], v[good]
    if n_rows is None:
        n_rows = int(r.max()) + 1
    order = np.lexsort((c, r))          # stable: duplicates stay in input order
    r, c, v = r[order], c[order], v[order]
    head = np.ones(r.size, dtype=bool)
    head[1:] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
    starts = np.flatnonzero(head)
    # duplicates are summed left to right in input order (scipy's own order is unspecified for runs
    # of three or more because csr_sort_indices uses an unstable sort; for one or two it is the same)
    run = np.diff(np.append(starts, v.size))
    data = v[starts].copy()
    for k in range(1, int(run.max()) if run.size else 0):
        more = run > k
        data[more] = data[more] + v[starts[more] + k]
    indices = c[starts].astype(np.int32)
    counts = np.bincount(r[starts], minlength=n_rows)
    indptr = np.zeros(n_rows + 1, dtype=np.int32)
    np.cumsum(counts, out=indptr[1:])
    return indptr, indices, data, (int(n_rows), int(n_cols) if n_cols is not None else None)


def csr_matrix(indptr, indices, data, shape):
    return sp.csr_matrix((data, indices, indptr), shape=shape)


# --------------------------------------------------------------------------------------------
# robust spread and sigma_extra  (LSsurf/RDE.py:10-18; LSsurf/calc_sigma_extra.py:13-44)
# --------------------------------------------------------------------------------------------

def rde(x):
    """Half the 16-84 percentile width by linear interpolation on sorted finite values."""
    x = np.asarray(x, dtype=float)
    fin = np.isfinite(x)
    n = int(fin.sum())
    if n < 2:
        return np.nan
    q = np.interp(np.array([0.16, 0.84]) * n, np.arange(0.5, n), np.sort(x[fin]))
    return (q[1] - q[0]) / 2.


def rde_select(x):
    """Selection form of ``rde``: four exact order statistics."""
    x = np.asarray(x, dtype=float)
    s = np.sort(x[np.isfinite(x)])
    n = s.size
    if n < 2:
        return np.nan
    q = []
    for p in (0.16, 0.84):
        u = p * n
        if u <= 0.5:
            q.append(s[0])
        elif u >= n - 0.5:
            q.append(s[-1])
        else:
            j = int(math.floor(u - 0.5))
            q.append((s[j + 1] - s[j]) * (u - (j + 0.5)) + s[j])
    return (q[1] - q[0]) / 2.


def fminbound(f, lo, hi, xatol=1e-5, maxfun=500):
    """Bounded scalar minimiser (Forsythe/Malcolm/Moler FMIN, as MATLAB fminbnd).

    Restates the third-party routine the reference calls at calc_sigma_extra.py:41
    (scipy.optimize.minimize_scalar(method='bounded'); scipy 1.18.1 `_optimize.py:2289-2433`):
    same constants (golden section 0.5*(3-sqrt 5), sqrt(2.2e-16), xatol/3) and the same
    acceptance tests, so the sequence of abscissae is identical for identical f values.
    Returns (x, fx, nfev).
    """
    if not (math.isfinite(lo) and math.isfinite(hi)):
        raise ValueError('Optimization bounds must be finite scalars.')
    if lo > hi:
        raise ValueError('The lower bound exceeds the upper bound.')
    seps = math.sqrt(2.2e-16)
    gold = 0.5 * (3.0 - math.sqrt(5.0))
    a, b = lo, hi
    v = w = xf = a + gold * (b - a)          # v: third-best (fulc), w: second-best (nfc)
    step = prev = 0.0                         # rat, e
    fx = f(xf)
    fv = fw = fx
    nfev = 1
    mid = 0.5 * (a + b)
    t1 = seps * abs(xf) + xatol / 3.0
    t2 = 2.0 * t1
    while abs(xf - mid) > (t2 - 0.5 * (b - a)):
        use_golden = True
        if abs(prev) > t1:
            use_golden = False
            r = (xf - w) * (fx - fv)
            q = (xf - v) * (fx - fw)
            p = (xf - v) * q - (xf - w) * r
            q = 2.0 * (q - r)
            if q > 0.0:
                p = -p
            q = abs(q)
            r = prev
            prev = step
            if abs(p) < abs(0.5 * q * r) and p > q * (a - xf) and p < q * (b - xf):
                step = (p + 0.0) / q
                x = xf + step
                if (x - a) < t2 or (b - x) < t2:
                    sgn = float(np.sign(mid - xf)) + float((mid - xf) == 0)
                    step = t1 * sgn
            else:
                use_golden = True
        if use_golden:
            prev = (a - xf) if xf >= mid else (b - xf)
            step = gold * prev
        sgn = float(np.sign(step)) + float(step == 0)
        x = xf + sgn * max(abs(step), t1)
        fu = f(x)
        nfev += 1
        if fu <= fx:
            if x >= xf:
                a = xf
            else:
                b = xf
            v, fv = w, fw
            w, fw = xf, fx
            xf, fx = x, fu
        else:
            if x < xf:
                a = x
            else:
                b = x
            if fu <= fw or w == xf:
                v, fv = w, fw
                w, fw = x, fu
            elif fu <= fv or v == xf or v == w:
                v, fv = x, fu
        mid = 0.5 * (a + b)
        t1 = seps * abs(xf) + xatol / 3.0
        t2 = 2.0 * t1
        if nfev >= maxfun:
            break
    return xf, fx, nfev


def sigma_extra(r, sigma, mask, groups=None):
    """sigma_extra >= 0 with RDE(r/sqrt(sigma^2+s^2)) = 1 per mask group (calc_sigma_extra.py:28-44).

    Groups with fewer than 10 selected points keep 0 (:34); the value is assigned to every member
    of the group, edited or not (:41); solver exceptions are swallowed (:40-43).
    """
    if groups is None:
        groups = {'all': np.ones(r.shape, dtype=bool)}
    out = np.zeros_like(r)
    for _, ii in groups.items():
        sel = ii & mask
        if sel.sum() < 10:
            continue
        rr, ss = r[sel], sigma[sel]
        top = rde(rr)
        try:
            out[ii] = fminbound(lambda s1: (rde(rr / np.sqrt(s1 ** 2 + ss ** 2)) - 1) ** 2, 0, top)[0]
        except Exception as e:   # same policy as the reference
            print(e)
    return out


# --------------------------------------------------------------------------------------------
# least-squares solve: stand-in for sparseqr.solve (LSsurf/smooth_fit.py:142)
# --------------------------------------------------------------------------------------------

def factor_normal(N, trailing=0):
    """Sparse LU of the SPD normal matrix N = A^T A with a fill-reducing order chosen for the fit's structure
    (with bias_params the bias columns must come last, else SuperLU's fill explodes):
    the last `trailing` columns (per-(rgt, cycle) biases: each touches every grid node under its track) stay last, the
    grid part is ordered by minimum degree (taken from a throw-away factorisation of the grid block alone, which is
    cheap); no pivoting (N is symmetric positive definite).  Returns (lu, perm): ``x[perm] = lu.solve(rhs[perm])``."""
    N = sp.csc_matrix(N)
    n = N.shape[0]
    opts = dict(diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    if trailing <= 0 or trailing >= n:
        lu = spl.splu(N, permc_spec='MMD_AT_PLUS_A', **opts)
        return lu, np.arange(n)
    ng = n - trailing
    pg = spl.splu(N[:ng][:, :ng].tocsc(), permc_spec='MMD_AT_PLUS_A', **opts).perm_c
    perm = np.concatenate([np.argsort(pg), np.arange(ng, n)])
    lu = spl.splu(N[perm][:, perm].tocsc(), permc_spec='NATURAL', **opts)
    return lu, perm


def solve_least_squares(A, b, dense_max=3000, refine=3, info=None, trailing=0):
    """Stand-in for sparseqr.solve (LSsurf/smooth_fit.py:142; SuiteSparseQR is not installable here): the unique
    least-squares solution of the full-column-rank system.  Small systems: dense lstsq (gold).  Otherwise SuperLU on the
    normal equations + `refine` steps of iterative refinement with the residual formed through A."""
    A = sp.csr_matrix(A)
    b = np.asarray(b, dtype=float).ravel()
    if A.shape[1] <= dense_max:
        return np.linalg.lstsq(A.toarray(), b, rcond=None)[0]
    At = A.T.tocsr()
    lu, perm = factor_normal(At @ A, trailing)

    def apply(r):
        x = np.empty_like(r)
        x[perm] = lu.solve(r[perm])
        return x
    x = apply(At @ b)
    for _ in range(refine):
        x = x + apply(At @ (b - A @ x))
    if info is not None:
        info.update(method='splu(AtA, minimum degree on the grid block, %d bias columns last) + %d refinements' % (trailing, refine),
                    fill=int(lu.L.nnz + lu.U.nnz))
    return x


# --------------------------------------------------------------------------------------------
# problem assembly  (LSsurf/grid_functions.py:26-134; constraint_functions.py:23-151;
#                    bias_functions.py:25-115; smooth_fit.py:406-634)
# --------------------------------------------------------------------------------------------

DEFAULT_MASK_SCALE = {0: 10, 1: 1}


def setup_grids(ctr, W, spacing):
    """z0 / dz / t grids; dz columns start after z0 (grid_functions.py:46, 68, 78-86)."""
    bds = {k: ctr[k] + np.array([-0.5, 0.5]) * W[k] for k in ('x', 'y', 't')}
    z0 = grid_geometry([bds['y'], bds['x']], spacing['z0'] * np.ones(2), name='z0')
    dz = grid_geometry([bds['y'], bds['x'], bds['t']], [spacing['dz'], spacing['dz'], spacing['dt']],
                       col_0=z0['n_nodes'], name='dz')
    t = grid_geometry([bds['t']], [spacing['dt']], name='t')
    return {'z0': z0, 'dz': dz, 't': t}, bds


def smoothness_constraints(grids, E_RMS, mask_scale=None):
    """Constraint families in stacking order with per-row expected sigma.

    Order and sigma rules: constraint_functions.py:40-78 with the all-True default mask
    (fd_grid.py:80) mapped through mask_scale (smooth_fit.py:362; lin_op.py:661-665).
    Returns a list of dict(op, name, rows, cols, vals, ind0, expected).
    """
    if mask_scale is None:
        mask_scale = DEFAULT_MASK_SCALE
    out = []
    ms = float(mask_scale[1])      # mask is all True == 1 without mask data

    def add(kind, grid, sig, masked=True):
        for name, offs, coef in stencil_families(grid, kind):
            r, c, v, i0 = stencil_triplets(grid, offs, coef)
            exp = sig * (np.zeros(r.shape[0]) + ms) if masked else np.zeros(r.shape[0]) + sig
            out.append(dict(op=kind, name=name, rows=r, cols=c, vals=v, ind0=i0, expected=exp))

    z0, dz = grids['z0'], grids['dz']
    if 'd2z0_dx2' in E_RMS:
        root_a = np.sqrt(np.prod(z0['delta']))
        add('grad2_z0', z0, E_RMS['d2z0_dx2'] / root_a)
    if 'dz0_dx' in E_RMS:
        add('grad_z0', z0, E_RMS['dz0_dx'] / root_a)
    if E_RMS.get('z0') is not None:
        name, offs, coef = stencil_families(z0, 'one')[0]
        r, c, v, i0 = stencil_triplets(z0, offs, coef)
        out.append(dict(op='mag_z0', name='mag_z0', rows=r, cols=c, vals=v, ind0=i0,
                        expected=E_RMS['z0'] / root_a * np.ones(r.shape[0])))
    root_v = np.sqrt(np.prod(dz['delta']))
    if E_RMS.get('d3z_dx2dt') is not None:
        add('grad2_dzdt', dz, E_RMS['d3z_dx2dt'] / root_v)
    if E_RMS.get('d2z_dxdt') is not None:
        add('grad_dzdt', dz, E_RMS['d2z_dxdt'] / root_v)
    if E_RMS.get('d2z_dt2') is not None:
        add('d2z_dt2', dz, E_RMS['d2z_dt2'] / root_v, masked=False)
    return out


def bias_ids(params):
    """Bias ID per point = index of its row among the lexicographically sorted unique rows of the
    bias-parameter columns, NaN -> -9999 (bias_functions.py:55-77)."""
    tab = np.nan_to_num(np.column_stack(params), nan=-9999)
    u, inv = np.unique(tab, axis=0, return_inverse=True)
    return np.asarray(inv).ravel().astype(np.int64), u


def assemble(x, y, t, z, sigma, ctr, W, spacing, E_RMS, reference_epoch=0, mask_scale=None,
             bias_params=None, sigma_corr=None, grid_bias=None):
    """Everything smooth_fit builds before iterate_fit (smooth_fit.py:406-634), as plain arrays.

    Returns a dict with the canonical CSR of G_data and Gc (full column space), Ed, Ec, rhs,
    the kept (non reference-epoch) columns, the valid-point mask and bookkeeping.
    ``grid_bias`` = dict(values, bounds=(y0, y1, x0, x1), spacing, expected_rms, expected_rms_grad): one gridded bias that
    multiplies a per-point covariate (setup_grid_bias.py:4-43), placed where the reference's predecessor driver called it
    (deprecated/smooth_xytb_fit.py:465-467): columns after dz, constraint rows after the smoothness families.
    """
    grids, _ = setup_grids(ctr, W, spacing)
    z0, dz = grids['z0'], grids['dz']
    valid = np.isfinite(z) & points_in_grid(z0, [y, x]) & points_in_grid(dz, [y, x, t])
    xs, ys, ts, zs, ss = x[valid], y[valid], t[valid], z[valid], sigma[valid]
    n = xs.size
    r0, c0, v0 = interp_triplets(z0, [ys, xs])
    r1, c1, v1 = interp_triplets(dz, [ys, xs, ts])
    rr = [r0.ravel(), r1.ravel()]
    cc = [c0.ravel(), c1.ravel()]
    vv = [v0.ravel(), v1.ravel()]
    col_N = dz['col_N']
    cons = smoothness_constraints(grids, E_RMS, mask_scale)
    if grid_bias is not None:
        b = grid_bias['bounds']
        gb = grid_geometry([[b[0], b[1]], [b[2], b[3]]], grid_bias['spacing'] * np.ones(2), col_0=col_N, name='grid_bias')
        rg, cg, vg = interp_triplets(gb, [ys, xs])
        rr.append(rg.ravel())
        cc.append(cg.ravel())
        vv.append(vg.ravel() * grid_bias['values'][valid][rg.ravel()])          # setup_grid_bias.py:29
        if grid_bias.get('expected_rms_grad') is not None:
            for name, offs, coef in stencil_families(gb, 'grad2_z0'):
                r_, c_, v_, i0 = stencil_triplets(gb, offs, coef)
                # the division by sqrt(cell area) applies to a zero array in the reference (setup_grid_bias.py:33-34)
                cons.append(dict(op='grad2_grid_bias', name=name + '_grid_bias', rows=r_, cols=c_, vals=v_, ind0=i0,
                                 expected=grid_bias['expected_rms_grad'] + np.zeros(r_.shape[0])))
        if grid_bias.get('expected_rms') is not None:
            k = np.arange(gb['n_nodes'], dtype=np.int64)
            cons.append(dict(op='mag_grid_bias', name='mag_grid_bias', rows=k[:, None], cols=(col_N + k)[:, None],
                             vals=np.ones((k.size, 1)), ind0=np.zeros(0, dtype=np.int64),
                             expected=grid_bias['expected_rms'] + np.zeros(k.size)))
        col_N += gb['n_nodes']
    bias = None
    if bias_params is not None:
        ids, u = bias_ids([p[valid] for p in bias_params])
        nb = int(ids.max()) + 1
        rr.append(np.arange(n, dtype=np.int64))
        cc.append(col_N + ids)
        vv.append(np.ones(n))
        sc = sigma_corr[valid]
        e_bias = np.array([np.nanmedian(sc[ids == k]) for k in range(nb)])
        if np.any(e_bias == 0):
            raise ValueError('found an zero value in the expected biases')
        k = np.arange(nb, dtype=np.int64)
        cons.append(dict(op='constraint_data_bias', name='constraint_data_bias', rows=k[:, None],
                         cols=(col_N + k)[:, None], vals=np.ones((nb, 1)), ind0=np.zeros(0, dtype=np.int64),
                         expected=e_bias))
        bias = dict(ids=ids, unique=u, col_0=col_N, n=nb, expected=e_bias)
        col_N += nb
    G = coo_to_csr(np.concatenate(rr), np.concatenate(cc), np.concatenate(vv), n_rows=n, n_cols=col_N)
    row_0, gr, gc, gv, ec, toc = 0, [], [], [], [], {}
    for fam in cons:
        ne = fam['rows'].shape[0]
        gr.append(fam['rows'].ravel() + row_0)
        gc.append(fam['cols'].ravel())
        gv.append(fam['vals'].ravel())
        ec.append(fam['expected'])
        toc[fam['name']] = np.arange(row_0, row_0 + ne)
        toc.setdefault(fam['op'], np.zeros(0, dtype=np.int64))
        if fam['op'] != fam['name']:
            toc[fam['op']] = np.concatenate([toc[fam['op']], toc[fam['name']]])
        row_0 += ne
    Ec = np.concatenate(ec)
    if np.any(ss == 0):
        raise ValueError('zero value found in data sigma')
    if np.any(Ec == 0):
        raise ValueError('zero value found in constraint sigma')
    C = coo_to_csr(np.concatenate(gr), np.concatenate(gc), np.concatenate(gv), n_rows=row_0, n_cols=col_N)
    # reference-epoch columns are removed from the solve (constraint_functions.py:128-151)
    iy, ix = np.meshgrid(np.arange(dz['shape'][0]), np.arange(dz['shape'][1]), indexing='ij')
    ref_cols = dz['col_0'] + iy.ravel() * dz['stride'][0] + ix.ravel() * dz['stride'][1] + reference_epoch
    keep = np.setdiff1d(np.arange(col_N), ref_cols)
    return dict(grids=grids, valid=valid, n_data=n, col_N=col_N, G=G, C=C, Ed=ss, Ec=Ec, z=zs,
                prior=np.zeros(row_0), keep_cols=keep, toc_rows=toc, bias=bias,
                dz_cols=np.arange(dz['col_0'], dz['col_N']), x=xs, y=ys, t=ts)


# --------------------------------------------------------------------------------------------
# the edit / reweight loop  (LSsurf/smooth_fit.py:100-210)
# --------------------------------------------------------------------------------------------

def iterate(sys_, max_iterations=10, min_iterations=2, converge_tol_dz=0.05, sigma_extra_relax=False,
            in_TSE=None, bias_nsigma_iteration=2, solver=None, history=None):
    """Outer loop: weighted LS solve on the selected rows, residuals, sigma_extra, 3-sigma edit.

    Row weights 1/sigma (smooth_fit.py:123-135), solve (:142), residual over ALL data rows
    (:146), sigma_extra (:153), edit (:159-163), the three convergence rules (:180-208).
    Returns m0 (full column space, reference epoch = 0), sigma_extra, in_TSE, rs_data.
    """
    if solver is None:
        nb = sys_['bias']['n'] if sys_.get('bias') else 0          # bias columns are the last kept columns
        solver = (lambda A, b: solve_least_squares(A, b, trailing=nb)) if nb else solve_least_squares
    G = csr_matrix(*sys_['G'][:3], shape=(sys_['n_data'], sys_['col_N']))
    C = csr_matrix(*sys_['C'][:3], shape=(sys_['Ec'].size, sys_['col_N']))
    keep = sys_['keep_cols']
    Gk, Ck = G[:, keep], C[:, keep]
    sig, z = sys_['Ed'], sys_['z']
    n = sys_['n_data']
    if in_TSE is None:
        in_TSE = np.ones(n, dtype=bool)
    s_extra = np.zeros(n)
    m0 = np.zeros(sys_['col_N'])
    last = max_iterations <= 1
    rs = None
    for it in range(max_iterations):
        e2 = sig ** 2
        if last and sigma_extra_relax:
            e2 = e2 + s_extra ** 2
        wd = 1. / np.sqrt(e2)
        sel = np.flatnonzero(in_TSE)
        A = sp.vstack([sp.diags(wd[sel]) @ Gk[sel, :], sp.diags(1. / sys_['Ec']) @ Ck]).tocsr()
        b = np.concatenate([wd[sel] * z[sel], sys_['prior'] / sys_['Ec']])
        m_prev = m0
        m0 = np.zeros(sys_['col_N'])
        m0[keep] = solver(A, b)
        r = z - G @ m0
        rs = r / sig
        if history is not None:
            history.append(dict(m0=m0.copy(), mask_in=in_TSE.copy(), r=r.copy()))
        if last:
            break
        s_extra = sigma_extra(r, sig, in_TSE)
        in_TSE = np.abs(r / np.sqrt(sig ** 2 + s_extra ** 2)) < 3.0
        if history is not None:
            history[-1].update(sigma_extra=s_extra.copy(), mask_out=in_TSE.copy())
        if not np.any(in_TSE):
            return m0, s_extra, in_TSE, rs
        if np.max(np.abs((m_prev - m0)[sys_['dz_cols']])) < converge_tol_dz and it > min_iterations:
            last = True
        if it >= max(min_iterations, bias_nsigma_iteration + 1):
            if np.all(s_extra < 0.5 * np.min(sig[in_TSE])):
                last = True
        if it == max_iterations - 2:
            last = True
    return m0, s_extra, in_TSE, rs


def smooth_fit(x, y, t, z, sigma, ctr, W, spacing, E_RMS, reference_epoch=0, max_iterations=10,
               bias_params=None, sigma_corr=None, history=None, grid_bias=None, **kw):
    """assemble + iterate; returns dict(m0, sigma_extra, in_TSE, z_est, valid, sys)."""
    s = assemble(x, y, t, z, sigma, ctr, W, spacing, E_RMS, reference_epoch,
                 bias_params=bias_params, sigma_corr=sigma_corr, grid_bias=grid_bias)
    m0, se, tse, rs = iterate(s, max_iterations=max_iterations, history=history, **kw)
    G = csr_matrix(*s['G'][:3], shape=(s['n_data'], s['col_N']))
    return dict(m0=m0, sigma_extra=se, in_TSE=tse, rs_data=rs, z_est=G @ m0, valid=s['valid'], sys=s)
