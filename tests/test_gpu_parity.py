"""GPU tier: the CUDA product (through the C ABI) against the oracle and the reference fixtures.

Bit-exact for integer / index / structure work and for the fp64 assembly values; 1e-6 relative L2
(BASELINE.json north_star) for solutions, written in each test; edit masks identical.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import lssurf_oracle as O
from tests.util import FIT_CASES, GOLDEN, check_parse_model_outputs, csr_digest, digest, load_case, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def L():
    import lssurf_b200
    lssurf_b200._cabi.lib()          # fails loudly if the .so or the GPU is missing
    return lssurf_b200


def make_data(L, pts):
    return L.point_data.data().from_dict(pts)


# ------------------------------------------------------------------------------------ assembly
@pytest.mark.parametrize('g', ['g1', 'g2', 'g3'])
def test_interp_mtx_edge_points(L, g):
    fx = np.load(f'{GOLDEN}/lowlevel.npz')
    bounds, deltas, col_0 = fx[g + '_bounds'], fx[g + '_deltas'], int(fx[g + '_col0'])
    grid = L.fd_grid(bounds, deltas, name=g, col_0=col_0)
    pts = list(fx[g + '_pts'])
    op = L.lin_op(grid, name='interp').interp_mtx(pts, bounds_error=False)
    assert op.N_eq == int(fx[g + '_nvalid'])
    assert op.r.dtype == np.int64 and op.v.dtype == np.float64 and op.r.shape == (op.N_eq, 2 ** len(bounds))
    assert digest(op.r, op.c, op.v) == str(fx[g + '_rcv'])               # reference triplets, bit for bit
    M = op.toCSR(row_N=len(pts[0]))
    assert M.has_canonical_format and csr_digest(M) == str(fx[g + '_csr'])
    with pytest.raises(ValueError) as e:
        L.lin_op(grid).interp_mtx(pts)
    assert str(e.value) == str(fx[g + '_err'])


def test_stencil_families(L):
    fx = np.load(f'{GOLDEN}/lowlevel.npz')
    g2 = L.fd_grid(fx['g2_bounds'], fx['g2_deltas'], name='z0')
    g3 = L.fd_grid(fx['g3_bounds'], fx['g3_deltas'], name='dz', col_0=g2.N_nodes)
    want = {s.split(':')[0]: s.split(':')[1:] for s in fx['stencils']}
    builders = {'grad2_z0': lambda: L.lin_op(g2, name='grad2_z0').grad2(DOF='z0'),
                'grad_z0': lambda: L.lin_op(g2, name='grad_z0').grad(DOF='z0'),
                'grad2_dzdt': lambda: L.lin_op(g3, name='grad2_dzdt').grad2_dzdt(DOF='z', t_lag=1),
                'grad_dzdt': lambda: L.lin_op(g3, name='grad_dzdt').grad_dzdt(DOF='z', t_lag=1),
                'd2z_dt2': lambda: L.lin_op(g3, name='d2z_dt2').d2z_dt2(DOF='z')}
    for kind, build in builders.items():
        op = build()
        d = digest(op.r.ravel(), op.c.ravel(), op.v.ravel(), np.asarray(op.ind0).astype(np.int64))
        assert [str(op.N_eq), d] == want[kind], kind
    offs = ([-1, -1, -1, 0, 0, 0, 1, 1, 1], [-1, 0, 1, -1, 0, 1, -1, 0, 1])
    coef = np.array([0.5, -1.25, 0.5, 2., -3.5, 2., 0.5, -1.25, 0.5])
    for op in (L.lin_op(g2, name='custom').diff_op(offs, coef), L.fd_op(g2, offs, coef, name='custom')):
        d = digest(op.r, op.c, op.v, np.asarray(op.ind0).astype(np.int64))
        assert [str(op.N_eq), d] == want['custom9']


def test_stencil_invalid_equations_and_which_nodes(L):
    g = L.fd_grid([[0., 6.], [0., 9.]], [1., 1.], name='z')
    go = O.grid_geometry([[0., 6.], [0., 9.]], [1., 1.], name='z')
    op = L.lin_op(g, name='a').diff_op(([-1, 0, 1], [0, 0, 2]), np.array([1., -2., 1.]), valid_equations_only=False)
    assert op.N_eq == g.N_nodes
    r, c, v = op.r, op.c, op.v
    iy, ix = np.unravel_index(np.arange(g.N_nodes), g.shape)
    for k, (dy, dx, w) in enumerate(zip([-1, 0, 1], [0, 0, 2], [1., -2., 1.])):
        ok = (iy + dy >= 0) & (iy + dy < g.shape[0]) & (ix + dx >= 0) & (ix + dx < g.shape[1])
        assert np.array_equal(v[:, k], w * ok) and np.array_equal(c[ok, k], (iy + dy)[ok] * g.shape[1] + (ix + dx)[ok])
        assert np.all(c[~ok, k] == 0)
    sel = np.array([0, 5, 17, 33])
    op2 = L.lin_op(g, name='one').one(which_nodes=sel)
    assert np.array_equal(op2.ind0, sel) and np.array_equal(op2.r.ravel(), np.arange(4)) and np.array_equal(op2.c.ravel(), sel)


def test_coo_to_csr_random_duplicates_and_empty_rows(L):
    rng = np.random.default_rng(11)
    n = 200_000
    r = rng.integers(0, 50_000, n); c = rng.integers(0, 3000, n); v = rng.normal(size=n)
    v[rng.random(n) < 0.05] = 0.
    r[:2000], c[:2000] = r[2000:4000], c[2000:4000]
    r[r == 17] = 18                                         # an empty row in the middle
    op = L.lin_op(col_N=3000)
    op.r, op.c, op.v = r, c, v
    M = op.toCSR(row_N=50_010)                              # trailing empty rows
    ip, ix, dat, _ = O.coo_to_csr(r, c, v, n_rows=50_010, n_cols=3000)
    assert np.array_equal(M.indptr, ip) and np.array_equal(M.indices, ix)
    assert digest(M.data) == digest(dat)                    # duplicates summed left to right in input order, bit exact
    # idempotence: CSR -> COO -> CSR is the identity
    coo = M.tocoo()
    op2 = L.lin_op(col_N=3000)
    op2.r, op2.c, op2.v = coo.row, coo.col, coo.data
    assert csr_digest(op2.toCSR(row_N=50_010)) == csr_digest(M)
    # default row_N = last row holding a non-zero (lin_op.py:751-752)
    assert op.toCSR().shape[0] == int(r[v != 0].max()) + 1
    # all-zero operator
    op3 = L.lin_op(col_N=5)
    op3.r, op3.c, op3.v = np.array([0, 1]), np.array([1, 2]), np.zeros(2)
    assert op3.toCSR(row_N=3).nnz == 0


def test_transpose_and_spmv_modes(L):
    rng = np.random.default_rng(12)
    rows = 30_000
    M = sp.random(rows, 4000, density=2e-3, random_state=5, format='csr')
    M = sp.vstack([M, sp.csr_matrix(np.ones((1, 4000)))]).tocsr()          # one long row (warp path + carries)
    M.sort_indices()
    _c = L._cabi
    A = L.DeviceCSR(_c.to_dev(M.indptr.astype(np.int32)), _c.to_dev(M.indices.astype(np.int32)), _c.to_dev(M.data), M.shape)
    x = rng.normal(size=4000); s = rng.uniform(0.5, 2, M.shape[0])
    xd, sd = _c.to_dev(x), _c.to_dev(s)
    tol = 1e-13 * np.abs(M).dot(np.abs(x)).max()
    assert np.max(np.abs(A.dot(xd).cpu().numpy() - M.dot(x))) <= tol
    assert np.max(np.abs(A.dot(xd, mode=_c.SPMV_SCALE, s=sd).cpu().numpy() - s * M.dot(x))) <= tol * 2
    assert np.max(np.abs(A.dot(xd, mode=_c.SPMV_RESIDUAL, s=sd).cpu().numpy() - (s - M.dot(x)))) <= tol
    At = A.transpose()
    Mt = M.T.tocsr(); Mt.sort_indices()
    assert np.array_equal(At.indptr.cpu().numpy(), Mt.indptr) and np.array_equal(At.indices.cpu().numpy(), Mt.indices)
    assert digest(At.data.cpu().numpy()) == digest(Mt.data)
    y = rng.normal(size=M.shape[0])
    # adjointness <A x, y> == <x, A^T y>
    lhs = float(np.dot(A.dot(xd).cpu().numpy(), y)); rhs = float(np.dot(x, At.dot(_c.to_dev(y)).cpu().numpy()))
    assert abs(lhs - rhs) <= 1e-12 * (abs(lhs) + 1)
    # determinism: bitwise repeatable
    assert digest(A.dot(xd).cpu().numpy()) == digest(A.dot(xd).cpu().numpy())


def test_argsort(L):
    rng = np.random.default_rng(13)
    keys = rng.integers(0, 1 << 37, 300_001).astype(np.uint64)
    keys[:1000] = keys[1000:2000]
    order = L._cabi.argsort_u64(L._cabi.to_dev(keys.view(np.int64)).view(torch.int64), 37).cpu().numpy()
    assert np.array_equal(order, np.argsort(keys, kind='stable'))


def test_dense_rank_group_median_and_bias_ids(L):
    """Device bias bookkeeping (lss_dense_rank_f64 / lss_group_median) vs np.unique / np.nanmedian and vs the host
    assign_bias_ID mirror of bias_functions.py:25-82."""
    rng = np.random.default_rng(21)
    n = 200_003
    x = rng.integers(-5, 40, n).astype(float)
    x[rng.random(n) < 0.01] = np.nan
    x[:3] = [0.0, -0.0, 1e-300]
    _c = L._cabi
    rank, uniq = _c.dense_rank(_c.to_dev(x), nan_value=-9999.0)
    u, inv = np.unique(np.nan_to_num(x, nan=-9999.0), return_inverse=True)
    assert np.array_equal(uniq, u) and np.array_equal(rank.cpu().numpy(), inv)
    ids = rng.integers(0, 37, n).astype(np.int32)
    vals = rng.normal(size=n)
    vals[rng.random(n) < 0.02] = np.nan
    ids[ids == 5] = 6                                            # an empty group
    med = _c.group_median(_c.to_dev(ids), _c.to_dev(vals), 37).cpu().numpy()
    with np.errstate(all='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            want = np.array([np.nanmedian(vals[ids == g]) if np.any(ids == g) else np.nan for g in range(37)])
    assert np.array_equal(np.isnan(med), np.isnan(want)) and np.array_equal(med[~np.isnan(med)], want[~np.isnan(want)])
    # the device bias model equals the host mirror of the reference routine
    from lssurf_b200.smooth_fit import _assign_bias_ID_device
    from lssurf_b200.bias_functions import assign_bias_ID
    m = 50_000
    rgt = rng.integers(0, 12, m).astype(float); cyc = rng.integers(1, 9, m).astype(float); cyc[rng.random(m) < 0.01] = np.nan
    sc = rng.uniform(0.01, 0.1, m)
    D = L.point_data.data().from_dict({'rgt': rgt, 'cycle': cyc, 'sigma_corr': sc, 'z': np.zeros(m)})
    D, bm = assign_bias_ID(D, ['rgt', 'cycle'])
    ids_d, bm_d = _assign_bias_ID_device({'rgt': _c.to_dev(rgt), 'cycle': _c.to_dev(cyc)}, ['rgt', 'cycle'], _c.to_dev(sc))
    assert np.array_equal(ids_d.cpu().numpy(), D.bias_ID.astype(int))
    assert bm_d['bias_param_dict']['ID'] == bm['bias_param_dict']['ID']
    assert all(bm_d['E_bias'][k] == bm['E_bias'][k] for k in bm['E_bias'])
    for k in bm['bias_ID_dict']:
        a, b = bm['bias_ID_dict'][k], bm_d['bias_ID_dict'][k]
        assert all((np.isnan(a[f]) and np.isnan(b[f])) or a[f] == b[f] for f in ('rgt', 'cycle'))


def test_rmatvec_and_spatial_order(L):
    rng = np.random.default_rng(22)
    M = sp.random(20_000, 3000, density=3e-3, random_state=7, format='csr'); M.sort_indices()
    _c = L._cabi
    A = L.DeviceCSR(_c.to_dev(M.indptr.astype(np.int32)), _c.to_dev(M.indices.astype(np.int32)), _c.to_dev(M.data), M.shape)
    u = rng.normal(size=M.shape[0]); u[::7] = 0.
    got = A.rdot(_c.to_dev(u)).cpu().numpy()
    want = M.T.dot(u)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.abs(M).T.dot(np.abs(u)).max()
    # rows that repeat a few column templates in runs AND interleaved (A B A B ...): the warp folds runs of equal columns
    # before the atomics, and equal columns that are not adjacent must not be folded across the row in between
    tmpl = [np.sort(rng.choice(400, 6, replace=False)) for _ in range(3)] + [np.sort(rng.choice(400, 3, replace=False))]
    which = np.concatenate([np.repeat([0, 1, 2, 3], 37), np.tile([0, 1], 64), rng.integers(0, 4, 4000)])
    cols = np.concatenate([tmpl[w] for w in which])
    ptr = np.concatenate([[0], np.cumsum([tmpl[w].size for w in which])])
    T = sp.csr_matrix((rng.normal(size=cols.size), cols.astype(np.int32), ptr.astype(np.int32)), shape=(which.size, 400))
    B = L.DeviceCSR(_c.to_dev(T.indptr.astype(np.int32)), _c.to_dev(T.indices.astype(np.int32)), _c.to_dev(T.data), T.shape)
    v = rng.normal(size=T.shape[0]); v[::5] = 0.
    got, want = B.rdot(_c.to_dev(v)).cpu().numpy(), T.T.dot(v)
    assert np.max(np.abs(got - want)) <= 1e-12 * np.abs(T).T.dot(np.abs(v)).max()
    fine = L.fd_grid([[0., 100.], [0., 100.]], [1., 1.], name='z0')
    coarse = L.fd_grid([[0., 100.], [0., 100.], [0., 2.]], [4., 4., 0.5], name='dz')
    y, x, t = rng.uniform(0, 100, 5000), rng.uniform(0, 100, 5000), rng.uniform(0, 2, 5000)
    order = _c.spatial_order(fine.as_struct(), coarse.as_struct(), [_c.to_dev(y), _c.to_dev(x), _c.to_dev(t)]).cpu().numpy()
    assert np.array_equal(np.sort(order), np.arange(5000))           # a permutation
    key = ((np.floor(y / 4) * 26 + np.floor(x / 4)) * 101 + np.floor(y)) * 101 + np.floor(x)
    key = key * 5 + np.floor(t / 0.5)
    assert np.all(np.diff(key[order]) >= 0)                           # sorted by (coarse cell, fine cell, epoch)


# ------------------------------------------------------------------------------------ robust statistics
def test_rde_and_sigma_extra_known_answers(L):
    fx = np.load(f'{GOLDEN}/lowlevel.npz')
    x = fx['rde_x']
    assert L.RDE(x) == float(fx['rde'])                      # the reference's own RDE value, bit for bit
    xs = x.copy(); xs[~np.isfinite(xs)] = 0.
    se = L.calc_sigma_extra(xs, fx['rde_sigma'], fx['rde_mask'])
    assert np.all(se == se[0]) and abs(se[0] - float(fx['sigma_extra'])) < 1e-12
    assert np.isnan(L.RDE(np.array([1.0]))) and np.isnan(L.RDE(np.array([np.nan, np.inf, 2.0])))
    rng = np.random.default_rng(14)
    for n in (2, 3, 9, 10, 11, 1000, 100_003):
        v = rng.standard_t(2, n) * 10 ** rng.uniform(-3, 3)
        assert L.RDE(v) == O.rde(v), n
    assert L.RDE(np.zeros(100)) == 0.0 and L.RDE(-np.ones(7)) == 0.0
    # fewer than 10 selected points: sigma_extra stays 0 (calc_sigma_extra.py:34)
    m = np.zeros(50, dtype=bool); m[:9] = True
    assert np.all(L.calc_sigma_extra(rng.normal(size=50), np.ones(50), m) == 0)
    # groups: each group gets its own value, assigned to all its members
    r = rng.normal(0, 3, 4000); sg = np.full(4000, 0.5); msk = rng.random(4000) < 0.95
    groups = {'a': np.arange(4000) < 1500, 'b': np.arange(4000) >= 1500}
    got = L.calc_sigma_extra(r, sg, msk, groups)
    want = O.sigma_extra(r, sg, msk, groups)
    assert np.max(np.abs(got - want)) < 1e-12


def test_sigma_extra_constant_sigma_path_is_bit_identical(L, monkeypatch):
    """With one sigma for all selected points the scaled residuals keep the order of the residuals for every trial value,
    so lss_sigma_extra evaluates RDE from the four order statistics it selected once (host arithmetic with the device's
    roundings).  Same bits and same number of Brent evaluations as the general path (a select over all points per
    evaluation); a single deviating sigma -- inside or outside the mask -- is handled, NaN residuals are skipped."""
    rng = np.random.default_rng(31)
    for n in (37, 5000, 300_001):
        r = rng.standard_t(3, n) * 0.7
        r[rng.random(n) < 0.001] = np.nan
        r_in = np.where(np.isfinite(r), r, 0.)
        msk = rng.random(n) < 0.9
        for sg in (np.full(n, 0.05), np.full(n, 3.0)):
            monkeypatch.delenv('LSS_SIGMA_EXTRA_GENERAL', raising=False)
            fast = L.calc_sigma_extra(r_in, sg, msk)
            monkeypatch.setenv('LSS_SIGMA_EXTRA_GENERAL', '1')
            slow = L.calc_sigma_extra(r_in, sg, msk)
            monkeypatch.delenv('LSS_SIGMA_EXTRA_GENERAL')
            assert np.array_equal(fast, slow), (n, sg[0])
            assert abs(fast[0] - O.sigma_extra(r_in, sg, msk)[0]) < 1e-12
        sg = np.full(n, 0.05)
        out = np.flatnonzero(~msk)[0]
        sg[out] = 7.0                                 # outside the mask: still constant where it counts
        assert np.array_equal(L.calc_sigma_extra(r_in, sg, msk), L.calc_sigma_extra(r_in, np.full(n, 0.05), msk))
        sg[np.flatnonzero(msk)[3]] = 0.0500001        # inside: the general path must take over
        assert abs(L.calc_sigma_extra(r_in, sg, msk)[0] - O.sigma_extra(r_in, sg, msk)[0]) < 1e-12


def test_sort_free_fast_paths_match_the_sorting_paths(L, monkeypatch):
    """Row-regular COO blocks -> CSR without a sort, small-integer dense ranks by presence table, medians of constant
    groups by min == max: each equals the general sorting path bit for bit (and falls back to it when it must)."""
    _c = L._cabi
    rng = np.random.default_rng(31)
    n = 50_000
    # interpolation-like blocks: ascending columns per row, some exact zeros, two grids + one bias column per row
    c1 = np.sort(rng.choice(1000, (n, 4)), axis=1).astype(np.int32); c1 += np.arange(4, dtype=np.int32) * 1000
    c2 = (4000 + np.sort(rng.choice(500, (n, 8)), axis=1) + np.arange(8) * 500).astype(np.int32)
    c3 = (8000 + rng.integers(0, 50, (n, 1))).astype(np.int32)
    v1, v2, v3 = rng.normal(size=(n, 4)), rng.normal(size=(n, 8)), np.ones((n, 1))
    v1[rng.random((n, 4)) < 0.2] = 0.0
    v2[::7] = 0.0
    blocks = [(_c.to_dev(c, torch.int32), _c.to_dev(v)) for c, v in ((c1, v1), (c2, v2), (c3, v3))]
    got = _c.blocks_to_csr(blocks, n, 8050)
    rows = np.repeat(np.arange(n, dtype=np.int32), 13)
    cols = np.concatenate([c1, c2, c3], axis=1).ravel(); vals = np.concatenate([v1, v2, v3], axis=1).ravel()
    want = _c.coo_to_csr(_c.to_dev(rows, torch.int32), _c.to_dev(cols, torch.int32), _c.to_dev(vals), n, 8050)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    c1[17, 0], c1[17, 1] = c1[17, 1], c1[17, 0]                       # one row out of order -> the caller must sort
    v1[17, :2] = 1.0
    assert _c.blocks_to_csr([(_c.to_dev(c1, torch.int32), _c.to_dev(v1))] + blocks[1:], n, 8050) is None
    assert _c.blocks_to_csr(blocks, n, 8049) is None                  # column outside the matrix
    # dense ranks of integer-valued doubles (with NaN -> -9999 and -0.0)
    x = rng.integers(-40, 1400, 300_000).astype(float); x[::11] = np.nan; x[5] = -0.0
    xd = _c.to_dev(x)
    r_fast, u_fast = _c.dense_rank(xd)
    monkeypatch.setenv('LSS_RANK_SORT', '1')
    r_sort, u_sort = _c.dense_rank(xd)
    assert torch.equal(r_fast, r_sort) and np.array_equal(u_fast, u_sort)
    # group medians: constant groups (fast) and mixed groups (fallback), an empty group in both
    ids = rng.integers(0, 40, 200_000).astype(np.int32); ids[ids == 13] = 12
    const = 0.01 * (1 + ids.astype(float))
    m_sort_c = _c.group_median(_c.to_dev(ids, torch.int32), _c.to_dev(const), 40).cpu().numpy()
    mixed = const.copy(); mixed[ids == 3] = rng.normal(size=int((ids == 3).sum()))
    m_sort_m = _c.group_median(_c.to_dev(ids, torch.int32), _c.to_dev(mixed), 40).cpu().numpy()
    monkeypatch.delenv('LSS_RANK_SORT')
    m_fast_c = _c.group_median(_c.to_dev(ids, torch.int32), _c.to_dev(const), 40).cpu().numpy()
    m_fast_m = _c.group_median(_c.to_dev(ids, torch.int32), _c.to_dev(mixed), 40).cpu().numpy()
    assert np.array_equal(m_fast_c, m_sort_c, equal_nan=True) and np.array_equal(m_fast_m, m_sort_m, equal_nan=True)
    assert np.isnan(m_fast_c[13]) and m_fast_c[5] == 0.06


def test_rde_candidate_selection_is_exact(L):
    """Above 2^16 points one RDE evaluation is a binned histogram pass + a compaction pass + an exact select over the
    candidates: the order statistics (and the interpolation between them) stay bit-identical to NumPy's sort for
    heavy tails, huge / tiny magnitudes, ties, masks, per-point sigmas and degenerate inputs."""
    rng = np.random.default_rng(21)
    _c = L._cabi
    n = 400_000
    cases = {'normal': rng.normal(0, 1, n), 'heavy': rng.standard_t(1.5, n) * 1e3, 'tiny': rng.normal(0, 1e-9, n),
             'huge': rng.normal(0, 1e7, n), 'ties': np.round(rng.normal(0, 2, n)), 'constant': np.full(n, 0.25),
             'one_sided': np.abs(rng.normal(0, 1, n)), 'mixed_zero': np.where(rng.random(n) < 0.7, 0.0, rng.normal(0, 1, n))}
    for name, v in cases.items():
        assert L.RDE(v) == O.rde(v), name
    r = rng.normal(0, 1, n); sg = rng.uniform(0.05, 2.0, n); m = rng.random(n) < 0.9
    for s1 in (0.0, 0.37, 5.0):
        got, n_used = _c.rde(_c.to_dev(r), _c.to_dev(sg), _c.to_dev(m.astype(np.uint8), torch.uint8), s1)
        want = O.rde((r / np.sqrt(s1 ** 2 + sg ** 2))[m])
        assert got == want and n_used == int(m.sum()), s1
    got = L.calc_sigma_extra(r * 3, sg, m)
    want = O.sigma_extra(r * 3, sg, m)
    assert np.max(np.abs(got - want)) < 1e-12


def test_edit_mask_kernel(L):
    rng = np.random.default_rng(15)
    n = 100_000
    r = rng.normal(0, 1, n); sg = rng.uniform(0.2, 0.5, n); se = np.full(n, 0.3)
    r[5] = np.nan
    r[6] = 3.0 * np.sqrt(sg[6] ** 2 + 0.09)                  # exactly on the threshold -> excluded (strict <)
    _c = L._cabi
    m, st = _c.edit_mask(_c.to_dev(r), _c.to_dev(sg), _c.to_dev(se), 3.0)
    want = np.abs(r / np.sqrt(sg ** 2 + se ** 2)) < 3.0
    assert np.array_equal(m.cpu().numpy().astype(bool), want)
    assert st['count'] == int(want.sum()) and st['min_sigma'] == sg[want].min() and st['near_threshold'] >= 1


# ------------------------------------------------------------------------------------ whole fits
@pytest.mark.parametrize('name', FIT_CASES)
def test_smooth_fit_against_reference_fixture(L, name):
    fx, pts, kw = load_case(name)
    D = make_data(L, pts)
    fo = L.smooth_fit(data=D.copy(), VERBOSE=False, return_fit_objects=True, **kw)
    G, C = fo['G_data'].toCSR(), fo['Gc'].toCSR()
    assert G.shape == tuple(fx['G_shape']) and C.shape == tuple(fx['C_shape'])
    assert csr_digest(G) == str(fx['G_digest'])              # CSR structure + values: bit exact vs the reference
    assert csr_digest(C) == str(fx['C_digest'])
    assert digest(np.asarray(fo['Ec'], dtype=float)) == str(fx['Ec_digest'])
    want_toc = {s.split(':')[0]: int(s.split(':')[1]) for s in fx['toc_rows']}
    got_toc = {k: len(np.unique(v)) for k, v in fo['Gc'].TOC['rows'].items()}
    assert got_toc == want_toc
    hist = []
    S = L.smooth_fit(data=D.copy(), VERBOSE=False, history=hist, **kw)
    assert len(S['timing']['pcg']) == int(fx['n_outer'])
    for k, packed in enumerate(fx['masks']):                 # identical edit mask after every outer iteration
        assert np.array_equal(hist[k]['mask_out'], packed), f'mask after iteration {k}'
        assert abs(hist[k]['sigma_extra']['all'] - fx['sigma_extra'][k]) < 2e-5
    assert np.array_equal(np.packbits(S['data'].three_sigma_edit), fx['in_TSE'])
    assert rel_l2(S['m']['all'], fx['m0']) < 1e-6            # north-star tolerance
    assert rel_l2(S['m']['z0'].z0.ravel(), fx['m0'][:S['m']['z0'].z0.size]) < 1e-6
    assert np.max(np.abs(S['data'].z_est - fx['z_est'])) < 1e-5
    ref_cols = S['grids']['dz'].global_ind([*np.meshgrid(np.arange(S['grids']['dz'].shape[0]), np.arange(S['grids']['dz'].shape[1]),
                                                         indexing='ij'), kw['reference_epoch'] + np.zeros(S['grids']['dz'].shape[:2], dtype=int)])
    assert np.all(S['m']['all'][ref_cols.ravel()] == 0.0)    # reference epoch held at exactly zero
    assert S['valid_data'].sum() == int(np.unpackbits(fx['in_TSE'])[:int(fx['n_data'])].sum())
    check_parse_model_outputs(fx, S)                         # counts, misfits, R / RMS, per-bias statistics


def test_smooth_fit_medium_against_oracle(L):
    """1e5 uniform points, 3-D fit with outliers: CUDA path vs the oracle run here on the same inputs."""
    from lssurf_b200 import synthetic
    pts, kw = synthetic.uniform_3d(n=100_000, width=2.e4, seed=3)
    hist_o, hist = [], []
    res = O.smooth_fit(pts['x'], pts['y'], pts['time'], pts['z'], pts['sigma'], kw['ctr'], kw['W'], kw['spacing'],
                       kw['E_RMS'], reference_epoch=kw['reference_epoch'], max_iterations=kw['max_iterations'], history=hist_o)
    S = L.smooth_fit(data=make_data(L, pts), VERBOSE=False, history=hist, **kw)
    assert len(hist) == sum('mask_out' in h for h in hist_o)
    for k, h in enumerate(hist):
        assert np.array_equal(h['mask_out'], np.packbits(hist_o[k]['mask_out'])), f'mask after iteration {k}'
    assert np.array_equal(S['data'].three_sigma_edit, res['in_TSE'])
    assert rel_l2(S['m']['all'], res['m0']) < 1e-6
    G = csr_digest(L.smooth_fit(data=make_data(L, pts), VERBOSE=False, return_fit_objects=True, **kw)['G_data'].toCSR())
    assert G == digest(*res['sys']['G'][:3])


def test_fused_and_generic_solver_agree(L):
    """Block form (fused ELL kernel with RED atomics) vs generic form (two deterministic CSR products): same system,
    same solution to solver tolerance, same masks."""
    import os
    from lssurf_b200 import synthetic
    pts, kw = synthetic.icesat2_tracks(n=150_000, width=2.e4, seed=6)
    S1 = L.smooth_fit(data=make_data(L, pts), VERBOSE=False, **kw)
    os.environ['LSS_NO_FUSED'] = '1'
    try:
        S2 = L.smooth_fit(data=make_data(L, pts), VERBOSE=False, **kw)
    finally:
        del os.environ['LSS_NO_FUSED']
    assert S1['timing']['pcg'][0]['fused'] and not S2['timing']['pcg'][0]['fused']
    assert rel_l2(S1['m']['all'], S2['m']['all']) < 1e-8        # both stop at rtol 1e-12 of their own preconditioned norm
    assert np.array_equal(S1['data'].three_sigma_edit, S2['data'].three_sigma_edit)
    assert np.allclose(S1['m']['bias']['val'], S2['m']['bias']['val'], rtol=0, atol=1e-8)


def test_explicit_normal_matrix_against_scipy(L):
    """lss_normal_*: pattern and values of N = A^T W A against scipy's (A^T W A) on the oracle's assembly of a small
    ICESat-2-shaped case ([G; C] rows, reference-epoch columns left out), then an incremental re-weighting."""
    from lssurf_b200 import synthetic
    from lssurf_b200.lin_op import DeviceCSR
    from lssurf_b200.solver import NormalMatrix
    pts, kw = synthetic.icesat2_tracks(n=40_000, width=1.e4, seed=2)
    s = O.assemble(pts['x'], pts['y'], pts['time'], pts['z'], pts['sigma'], kw['ctr'], kw['W'], kw['spacing'], kw['E_RMS'],
                   kw['reference_epoch'], bias_params=[pts['rgt'], pts['cycle']], sigma_corr=pts['sigma_corr'])
    G = O.csr_matrix(*s['G'][:3], shape=(s['n_data'], s['col_N']))
    Cm = O.csr_matrix(*s['C'][:3], shape=(s['Ec'].size, s['col_N']))
    A = sp.vstack([G, Cm]).tocsr()
    A.sort_indices()
    fixed = np.setdiff1d(np.arange(s['col_N']), s['keep_cols'])
    keep = np.zeros(s['col_N']); keep[s['keep_cols']] = 1.0
    dev = lambda a, dt: L._cabi.to_dev(np.asarray(a), dt)
    Ad = DeviceCSR(dev(A.indptr, torch.int32), dev(A.indices, torch.int32), dev(A.data, torch.float64), A.shape)
    Nm = NormalMatrix(Ad, fixed_cols=fixed)
    rng = np.random.default_rng(0)
    w = np.r_[1. / s['Ed'] ** 2, 1. / s['Ec'] ** 2]
    K = sp.diags(keep)

    def check(w):
        Nm.update(dev(w, torch.float64))
        ip, ix, dat = [t.cpu().numpy() for t in Nm.export()]
        got = sp.csr_matrix((dat, ix, ip), shape=(A.shape[1],) * 2)
        assert got.has_sorted_indices
        want = (K @ (A.T @ sp.diags(w) @ A) @ K).tocsr()
        want.eliminate_zeros()
        # the pattern is weight independent: it must cover the exact product's pattern, values agree to rounding
        d = (got - want).tocsr()
        scale = np.abs(want.data).max()
        assert (np.abs(d.data).max() if d.nnz else 0.0) < 1e-12 * scale
        P = got.copy(); P.data[:] = 1.0
        Pw = (K @ (abs(A).T @ abs(A)) @ K).tocsr(); Pw.eliminate_zeros(); Pw.data[:] = 1.0
        assert (P - Pw).nnz == 0 and got.nnz == Nm.nnz
        x = rng.normal(size=A.shape[1])
        y = Nm.dot(dev(x, torch.float64)).cpu().numpy()
        assert np.allclose(y, want @ x, rtol=0, atol=1e-11 * np.abs(want @ x).max())

    check(w)
    w2 = w.copy()
    off = rng.choice(s['n_data'], 900, replace=False)
    w2[off] = 0.0                                                   # 3-sigma edit switches rows off ...
    check(w2)
    w2[off[:300]] = w[off[:300]] * 0.25                             # ... and a relaxed pass re-weights some
    check(w2)
    Nm.close()


def test_operator_forms_and_preconditioners_agree(L):
    """Explicit normal matrix with block-Jacobi + eliminated bias columns (default) vs the same with scalar Jacobi vs
    fused ELL pass vs generic CSR pair: same masks, same solution; the structured preconditioner needs fewer iterations."""
    import os
    from lssurf_b200 import synthetic
    pts, kw = synthetic.icesat2_tracks(n=150_000, width=2.e4, seed=7)
    out = {}
    cases = (('structured', None, 'explicit_normal'), ('blocks_only', 'LSS_NO_ELIM', 'explicit_normal'),
             ('elim_only', 'LSS_NO_BLOCKS', 'explicit_normal'), ('jacobi', 'LSS_NO_STRUCTURE', 'explicit_normal'),
             ('fused_ell', 'LSS_NO_EXPLICIT_N', 'fused_ell'), ('csr_pair', 'LSS_NO_FUSED', 'csr_pair'))
    for name, env, form in cases:
        if env:
            os.environ[env] = '1'
        try:
            out[name] = L.smooth_fit(data=make_data(L, pts), VERBOSE=False, **kw)
        finally:
            if env:
                del os.environ[env]
        assert out[name]['timing']['pcg'][0]['form'] == form
    ref = out['csr_pair']
    for name in out:
        if name == 'csr_pair':
            continue
        assert rel_l2(out[name]['m']['all'], ref['m']['all']) < 1e-8, name     # each stops at rtol 1e-12 of its own preconditioned norm
        assert np.array_equal(out[name]['data'].three_sigma_edit, ref['data'].three_sigma_edit), name
        assert np.allclose(out[name]['m']['bias']['val'], ref['m']['bias']['val'], rtol=0, atol=1e-8), name
    p0 = out['structured']['timing']['pcg'][0]
    assert p0['nnz_N'] > 0 and p0['precond_blocks'] > 0 and p0['eliminated'] == len(ref['m']['bias']['val'])
    assert out['blocks_only']['timing']['pcg'][0]['eliminated'] == 0 and out['elim_only']['timing']['pcg'][0]['precond_blocks'] == 0
    assert out['jacobi']['timing']['pcg'][0]['precond_blocks'] == 0
    its = {k: sum(p['iterations'] for p in v['timing']['pcg']) for k, v in out.items()}
    print('CG iterations per fit:', its)
    assert its['structured'] < 0.7 * its['jacobi']


def test_normal_product_and_step_kernel_variants_agree(L):
    """Every product kernel for the explicit normal matrix (nnz-tiled, row-per-lane-group, padded
    chunks of 8, TMA-streamed row-aligned tiles with and without per-tile column dictionaries) and every CG vector-step kernel (warp per block, thread per block row, CTA tile) give the same fit: same masks, solutions equal to solver tolerance."""
    from lssurf_b200 import synthetic
    pts, kw = synthetic.icesat2_tracks(n=150_000, width=2.e4, seed=7)
    out = {}
    try:
        for prod, step in ((0, 0), (0, 1), (1, 2), (5, 1), (8, 2), (10, 2), (11, 2), (11, 0)):
            L._cabi.set_option('normal_product', prod)
            L._cabi.set_option('step_kernel', step)
            out[(prod, step)] = L.smooth_fit(data=make_data(L, pts), VERBOSE=False, **kw)
            assert out[(prod, step)]['timing']['pcg'][0]['form'] == 'explicit_normal'
    finally:
        L._cabi.set_option('normal_product', 0)
        L._cabi.set_option('step_kernel', 2)
    ref = out[(0, 0)]
    its0 = sum(p['iterations'] for p in ref['timing']['pcg'])
    for key, o in out.items():
        assert rel_l2(o['m']['all'], ref['m']['all']) < 1e-9, key
        assert np.array_equal(o['data'].three_sigma_edit, ref['data'].three_sigma_edit), key
        its = sum(p['iterations'] for p in o['timing']['pcg'])
        assert abs(its - its0) <= 0.05 * its0 + 5, (key, its, its0)


def test_two_level_preconditioner_agrees_with_block_jacobi(L, monkeypatch):
    """The coarse level (lss_solver_set_coarse: Galerkin Schur operator on a bilinear coarse space, dense inverse,
    fused into the CG step kernel) only changes the iteration count: same masks after every iteration, same solution
    to solver tolerance, for a coarse operator frozen at the first weights (default) and for one rebuilt with every mask, with and without bias elimination."""
    from lssurf_b200 import synthetic
    for gen, n, width in ((synthetic.icesat2_tracks, 300_000, 3.e4), (synthetic.uniform_3d, 200_000, 3.e4)):
        pts, kw = gen(n=n, width=width, seed=11)
        out = {}
        for key, env in (('bj', {'LSS_COARSE': '0'}), ('ml', {'LSS_COARSE': '1', 'LSS_COARSE_MIN_BLOCKS': '1'}),
                         ('ml_refresh', {'LSS_COARSE': '1', 'LSS_COARSE_MIN_BLOCKS': '1', 'LSS_COARSE_REFRESH': '1'}),
                        # the coarse inverse by the plain Gauss-Jordan sweep (one grid barrier per pivot) instead of the blocked one
                        ('ml_plain_gj', {'LSS_COARSE': '1', 'LSS_COARSE_MIN_BLOCKS': '1', 'LSS_GJ_PLAIN': '1'})):
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            hist = []
            out[key] = L.smooth_fit(data=make_data(L, pts), VERBOSE=False, history=hist, **kw)
            out[key]['hist'] = hist
            for k in env:
                monkeypatch.delenv(k)
        assert out['bj']['timing']['pcg'][0]['coarse'] == 0 and out['ml']['timing']['pcg'][0]['coarse'] > 0
        its = {k: sum(p['iterations'] for p in o['timing']['pcg']) for k, o in out.items()}
        print(gen.__name__, 'CG iterations', its, 'coarse unknowns', out['ml']['timing']['pcg'][0]['coarse'])
        assert abs(its['ml'] - its['ml_plain_gj']) <= 3, its          # the two inverses agree to rounding
        for key in ('ml', 'ml_refresh', 'ml_plain_gj'):
            assert rel_l2(out[key]['m']['all'], out['bj']['m']['all']) < 1e-9, key
            assert len(out[key]['hist']) == len(out['bj']['hist'])
            for a, b in zip(out[key]['hist'], out['bj']['hist']):
                assert np.array_equal(a['mask_out'], b['mask_out']), key
            assert its[key] < 1.25 * its['bj'], (key, its)      # the coarse level pays off on large domains only (bench workload: -60 %)


def test_structured_normal_product_agrees_with_the_csr_product(L, monkeypatch):
    """The grid-structured layout of N (lss_solver_set_structured: z0-dz pairs stored once for both triangles, banded dz-dz
    blocks, z0-z0 diagonals, CSR remainder) only changes how q = N p is computed: same masks after every iteration, same
    solution to solver tolerance as the CSR product; with the eliminated bias block (remainder = bias coupling) and without."""
    from lssurf_b200 import synthetic
    for gen, n, width in ((synthetic.icesat2_tracks, 300_000, 3.e4), (synthetic.uniform_3d, 200_000, 3.e4)):
        pts, kw = gen(n=n, width=width, seed=23)
        out = {}
        for key in ('0', '1'):
            monkeypatch.setenv('LSS_STRUCTURED_N', key)
            hist = []
            out[key] = L.smooth_fit(data=make_data(L, pts), VERBOSE=False, history=hist, **kw)
            out[key]['hist'] = hist
        monkeypatch.delenv('LSS_STRUCTURED_N')
        assert out['0']['timing']['pcg'][0]['structured_bytes'] == 0 and out['1']['timing']['pcg'][0]['structured_bytes'] > 0
        assert out['1']['timing']['pcg'][0]['structured_bytes'] < 12 * out['0']['timing']['pcg'][0]['nnz_product']
        assert rel_l2(out['1']['m']['all'], out['0']['m']['all']) < 1e-9
        assert len(out['1']['hist']) == len(out['0']['hist'])
        for a, b in zip(out['1']['hist'], out['0']['hist']):
            assert np.array_equal(a['mask_out'], b['mask_out']), a['iteration']
        its = [sum(p['iterations'] for p in out[k]['timing']['pcg']) for k in ('0', '1')]
        print(gen.__name__, 'CG iterations csr / structured', its)
        assert abs(its[0] - its[1]) <= 0.05 * its[0] + 5


def test_row_sharded_solver_on_one_rank_matches_the_plain_solver(L, monkeypatch):
    """The row-sharded PCG (own blocks of N, p exchanged through the P2P mailbox inside the step kernel) run with a single
    rank -- the mailbox is then local, every exchange goes to itself -- gives the plain solver's fit: same masks after
    every iteration, same solution; with and without the coarse level and the eliminated bias block."""
    from lssurf_b200 import synthetic, parallel
    dist = parallel.single_rank_context()
    try:
        for gen, n, width in ((synthetic.icesat2_tracks, 200_000, 2.5e4), (synthetic.uniform_3d, 150_000, 2.5e4)):
            pts, kw = gen(n=n, width=width, seed=17)
            for coarse in ('0', '1'):
                monkeypatch.setenv('LSS_COARSE', coarse)
                monkeypatch.setenv('LSS_COARSE_MIN_BLOCKS', '1')
                h0, h1 = [], []
                ref = L.smooth_fit(data=make_data(L, pts), VERBOSE=False, history=h0, **kw)
                monkeypatch.setenv('LSS_SHARD_SINGLE', '1')
                got = L.smooth_fit(data=make_data(L, pts), VERBOSE=False, history=h1, dist=dist, **kw)
                monkeypatch.delenv('LSS_SHARD_SINGLE')
                assert rel_l2(got['m']['all'], ref['m']['all']) < 1e-9, (gen.__name__, coarse)
                assert len(h0) == len(h1)
                for a, b in zip(h0, h1):
                    assert np.array_equal(a['mask_out'], b['mask_out']), (gen.__name__, coarse, a['iteration'])
                assert np.array_equal(got['data'].three_sigma_edit, ref['data'].three_sigma_edit)
                assert np.allclose(got['data'].z_est, ref['data'].z_est, rtol=0, atol=1e-7)
                its = [sum(p['iterations'] for p in o['timing']['pcg']) for o in (ref, got)]
                print(gen.__name__, 'coarse', coarse, 'CG iterations plain / sharded', its)
                assert abs(its[0] - its[1]) <= 0.05 * its[0] + 5
    finally:
        dist.destroy()


def test_large_scale_properties(L):
    """Full-size properties (cfg 3 shapes, 1e6 points) that need no oracle: row sums of the interpolation
    matrix are exactly 1 up to rounding, CSR is canonical, G^T products are adjoint, the solve reduces the residual."""
    from lssurf_b200 import synthetic
    pts, kw = synthetic.uniform_3d(n=1_000_000, seed=0)
    fo = L.smooth_fit(data=make_data(L, pts), VERBOSE=False, return_fit_objects=True, **kw)
    Gd = fo['G_data']._csr_device(row_N=1_000_000)
    assert Gd.shape == (1_000_000, 254_622)
    ip = Gd.indptr.cpu().numpy(); ix = Gd.indices.cpu().numpy()
    assert ip[0] == 0 and np.all(np.diff(ip) >= 0) and ip[-1] == Gd.nnz
    assert np.all((np.diff(ix) > 0) | np.isin(np.arange(1, ix.size), ip[1:-1]))      # sorted, no duplicates inside rows
    ones = torch.ones(Gd.shape[1], dtype=torch.float64, device='cuda')
    rs = Gd.dot(ones).cpu().numpy()
    assert np.max(np.abs(rs - 2.0)) < 1e-12                                           # z0 weights + dz weights each sum to 1
    rng = np.random.default_rng(0)
    x = L._cabi.to_dev(rng.normal(size=Gd.shape[1])); y = L._cabi.to_dev(rng.normal(size=Gd.shape[0]))
    lhs = float((Gd.dot(x) * y).sum().item()); rhs = float((x * Gd.transpose().dot(y)).sum().item())
    assert abs(lhs - rhs) < 1e-10 * abs(lhs)
