"""CPU tier: the oracle restatement against the golden fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py), plus third-party restatements against the real libraries."""
import numpy as np
import pytest
import scipy.optimize
import scipy.sparse as sp

from oracle import lssurf_oracle as O
from tests.util import FIT_CASES, digest, load_case, rel_l2, GOLDEN


@pytest.mark.parametrize('name', FIT_CASES)
def test_oracle_matches_reference_fixture(name):
    fx, pts, kw = load_case(name)
    hist = []
    bias = [pts[k] for k in kw['bias_params']] if kw.get('bias_params') else None
    res = O.smooth_fit(pts['x'], pts['y'], pts['time'], pts['z'], pts['sigma'], kw['ctr'], kw['W'], kw['spacing'],
                       kw['E_RMS'], reference_epoch=kw['reference_epoch'], max_iterations=kw['max_iterations'],
                       bias_params=bias, sigma_corr=pts.get('sigma_corr'), history=hist)
    s = res['sys']
    # structure: bit exact (sha256 of the reference's CSR arrays)
    assert digest(*s['G'][:3]) == str(fx['G_digest'])
    assert digest(*s['C'][:3]) == str(fx['C_digest'])
    assert digest(s['Ec']) == str(fx['Ec_digest'])
    assert s['n_data'] == int(fx['n_data']) and len(hist) == int(fx['n_outer'])
    # edit masks after every outer iteration: identical
    for k, packed in enumerate(fx['masks']):
        assert np.array_equal(np.packbits(hist[k]['mask_out']), packed), f'mask after iteration {k}'
        assert abs(np.median(hist[k]['sigma_extra']) - fx['sigma_extra'][k]) < 1e-5
    assert np.array_equal(np.packbits(res['in_TSE']), fx['in_TSE'])
    assert rel_l2(res['m0'], fx['m0']) < 1e-9
    assert np.max(np.abs(res['z_est'] - fx['z_est'])) < 1e-7


def test_oracle_lowlevel_fixture():
    fx = np.load(f'{GOLDEN}/lowlevel.npz')
    for g in ('g1', 'g2', 'g3'):
        geo = O.grid_geometry(fx[g + '_bounds'], fx[g + '_deltas'], col_0=int(fx[g + '_col0']), name=g)
        pts = list(fx[g + '_pts'])
        r, c, v = O.interp_triplets(geo, pts, bounds_error=False)
        assert digest(r.astype(np.int64), c.astype(np.int64), v) == str(fx[g + '_rcv'])
        assert r.shape[0] == int(fx[g + '_nvalid'])
        ip, ix, dat, _ = O.coo_to_csr(r, c, v, n_rows=len(pts[0]), n_cols=geo['col_N'])
        assert digest(ip, ix, dat) == str(fx[g + '_csr'])
        with pytest.raises(ValueError) as e:
            O.interp_triplets(geo, pts)
        assert str(e.value) == str(fx[g + '_err'])
    assert O.rde(fx['rde_x']) == float(fx['rde']) == O.rde_select(fx['rde_x'])
    x = fx['rde_x'].copy()
    x[~np.isfinite(x)] = 0.
    assert O.sigma_extra(x, fx['rde_sigma'], fx['rde_mask'])[0] == float(fx['sigma_extra'])


def test_oracle_stencil_fixture():
    fx = np.load(f'{GOLDEN}/lowlevel.npz')
    g2 = O.grid_geometry(fx['g2_bounds'], fx['g2_deltas'], name='z0')
    g3 = O.grid_geometry(fx['g3_bounds'], fx['g3_deltas'], col_0=g2['n_nodes'], name='dz')
    want = {s.split(':')[0]: s.split(':')[1:] for s in fx['stencils']}
    for kind, g in (('grad2_z0', g2), ('grad_z0', g2), ('grad2_dzdt', g3), ('grad_dzdt', g3), ('d2z_dt2', g3)):
        row_0, rr, cc, vv, i0 = 0, [], [], [], []
        for name, offs, coef in O.stencil_families(g, kind):
            r, c, v, ind0 = O.stencil_triplets(g, offs, coef)
            rr.append(r.ravel() + row_0); cc.append(c.ravel()); vv.append(v.ravel()); i0.append(ind0)
            row_0 += r.shape[0]
        d = digest(np.concatenate(rr), np.concatenate(cc), np.concatenate(vv), np.concatenate(i0))
        assert [str(row_0), d] == want[kind], kind


def test_fminbound_matches_scipy():
    """The Brent restatement follows scipy's iterates exactly (same x, f, nfev)."""
    rng = np.random.default_rng(3)
    for k in range(40):
        a, b, c = rng.uniform(0.1, 3), rng.uniform(0.2, 5), rng.uniform(-1, 1)
        f = lambda x: (np.sin(a * x) + c * x - 0.3) ** 2 + 0.01 * abs(x - b)   # noqa: E731
        hi = rng.uniform(0.5, 10)
        ref = scipy.optimize.minimize_scalar(f, method='bounded', bounds=[0, hi])
        x, fx, nfev = O.fminbound(f, 0, hi)
        assert x == ref['x'] and fx == ref['fun'] and nfev == ref['nfev']
    assert O.fminbound(lambda x: x * x, 0., 0.)[0] == 0.


def test_coo_to_csr_matches_scipy():
    rng = np.random.default_rng(5)
    n = 5000
    r = rng.integers(0, 300, n); c = rng.integers(0, 200, n); v = rng.normal(size=n)
    v[rng.random(n) < 0.1] = 0.
    r[:50], c[:50] = r[50:100], c[50:100]          # duplicates
    M = sp.csr_matrix((v[v != 0], (r[v != 0], c[v != 0])), shape=(301, 200))
    ip, ix, dat, _ = O.coo_to_csr(r, c, v, n_rows=301, n_cols=200)
    assert np.array_equal(ip, M.indptr) and np.array_equal(ix, M.indices)
    assert np.allclose(dat, M.data, rtol=1e-15, atol=0)


def test_rde_edge_cases():
    assert np.isnan(O.rde(np.array([1.0])))
    assert np.isnan(O.rde(np.array([np.nan, np.inf])))
    for n in (2, 3, 4, 7, 50):
        x = np.random.default_rng(n).normal(size=n)
        assert O.rde(x) == O.rde_select(x)
