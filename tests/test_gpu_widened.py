"""GPU tier, SURVEY section 8(f) rows: per-sensor gridded bias, per-sensor slope bias and gridded sigma_extra, through the
real C ABI, against fixtures taken from the UNMODIFIED reference (oracle/make_golden.py::case_widened).  Same bars as
the hot path: CSR of G_data / Gc bit-exact, identical edit mask after every outer iteration, model within 1e-6."""
import numpy as np
import pytest

from tests.test_host_double import check_widened_outputs
from tests.util import check_parse_model_outputs, WIDENED_CASES, csr_digest, digest, load_case, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def L():
    import lssurf_b200
    lssurf_b200._cabi.lib()          # fails loudly if the .so or the GPU is missing
    return lssurf_b200


def test_bias_filter_groups_like_the_filtered_parameters(L):
    """bias_filter (bias_functions.py:53-57): a filter that blanks `cycle` gives one bias per rgt -- the same operator
    as a fit whose data carry a constant cycle."""
    fx, pts, kw = load_case('tracks_bias_small')

    def per_track(D):
        out = D.copy()
        out.assign({'cycle': np.zeros_like(np.asarray(D.cycle))})
        return out
    a = L.smooth_fit(data=L.point_data.data().from_dict(pts), VERBOSE=False, return_fit_objects=True, bias_filter=per_track, **kw)
    b = L.smooth_fit(data=L.point_data.data().from_dict(dict(pts, cycle=np.zeros_like(pts['cycle']))), VERBOSE=False,
                     return_fit_objects=True, **kw)
    assert csr_digest(a['G_data'].toCSR()) == csr_digest(b['G_data'].toCSR())
    assert csr_digest(a['Gc'].toCSR()) == csr_digest(b['Gc'].toCSR()) and np.array_equal(a['Ec'], b['Ec'])
    assert a['G_data'].shape[1] < int(fx['G_shape'][1])


def test_setup_grid_bias_operators(L):
    """setup_grid_bias (setup_grid_bias.py:4-43) at the operator level against the reference digests."""
    from tests.test_host_double import check_grid_bias, grid_bias_operators
    check_grid_bias(*grid_bias_operators(L))


def test_iterate_fit_reference_signature(L):
    """iterate_fit with the reference's own argument list (smooth_fit.py:100-101), NumPy in / NumPy out."""
    from tests.test_host_double import iterate_fit_like_a_reference_caller
    iterate_fit_like_a_reference_caller(L)


@pytest.mark.parametrize('name', WIDENED_CASES)
def test_widened_fit_against_reference_fixture(L, name):
    fx, pts, kw = load_case(name)
    D = L.point_data.data().from_dict(pts)
    fo = L.smooth_fit(data=D.copy(), VERBOSE=False, return_fit_objects=True, **kw)
    G, C = fo['G_data'].toCSR(), fo['Gc'].toCSR()
    assert G.shape == tuple(fx['G_shape']) and C.shape == tuple(fx['C_shape'])
    assert csr_digest(G) == str(fx['G_digest']) and csr_digest(C) == str(fx['C_digest'])
    assert digest(np.asarray(fo['Ec'], dtype=float)) == str(fx['Ec_digest'])
    want_toc = {s.split(':')[0]: int(s.split(':')[1]) for s in fx['toc_rows']}
    assert {k: len(np.unique(v)) for k, v in fo['Gc'].TOC['rows'].items()} == want_toc
    hist = []
    S = L.smooth_fit(data=D.copy(), VERBOSE=False, history=hist, **kw)
    assert len(S['timing']['pcg']) == int(fx['n_outer'])
    for k, packed in enumerate(fx['masks']):
        assert np.array_equal(hist[k].get('mask_next', hist[k]['mask_out']), packed), f'mask after iteration {k}'
    assert np.array_equal(np.packbits(S['data'].three_sigma_edit), fx['in_TSE'])
    assert rel_l2(S['m']['all'], fx['m0']) < 1e-6            # north-star tolerance
    assert np.max(np.abs(S['data'].z_est - fx['z_est'])) < 1e-5
    check_widened_outputs(fx, S)
    check_parse_model_outputs(fx, S)


def test_compute_E_against_reference_fixture(L):
    """Error propagation: per-node, per-operator and per-bias errors from the covariance of the device
    normal matrix against the reference's R^-1 row norms."""
    from tests.test_host_double import check_compute_E
    S = check_compute_E(L)
    assert 'compute_E' not in S['timing'] or S['timing']['compute_E'] >= 0


def test_errors_and_early_returns_like_the_reference(L):
    """ValueErrors, silent point drops and the no-data early return of the reference driver, on the device path."""
    from tests.test_host_double import check_errors_and_early_returns
    check_errors_and_early_returns(L)


def test_calc_sigma_extra_on_grid_numpy_api(L):
    """Reference signature, NumPy in / out: constant noise -> one value everywhere; a noisier half -> larger values there."""
    from lssurf_b200.calc_sigma_extra import calc_sigma_extra_on_grid
    rng = np.random.default_rng(3)
    n = 40_000
    x, y = rng.uniform(-2e4, 2e4, n), rng.uniform(-2e4, 2e4, n)
    sigma = np.full(n, 0.1)
    r = rng.normal(0, 0.1, n) + rng.normal(0, 0.5, n) * (x > 0)
    se = calc_sigma_extra_on_grid(x, y, r, sigma, np.ones(n, dtype=bool), spacing=1.e4)
    assert se.shape == (n,) and np.all(np.isfinite(se))
    assert np.median(se[x > 1.2e4]) > 0.3 and np.median(se[x < -1.2e4]) < 0.05
    capped = calc_sigma_extra_on_grid(x, y, r, sigma, np.ones(n, dtype=bool), spacing=1.e4, sigma_extra_max=0.2)
    assert capped.max() <= 0.2 + 1e-12


def test_true_div_is_ieee_division(L):
    """`tensor / python_scalar` on CUDA multiplies by the rounded reciprocal; `_cabi.true_div` must give NumPy's bits
    (data_slope_bias.py:47 divides by 1000, calc_sigma_extra.py:73-74 by the bin spacing)."""
    rng = np.random.default_rng(11)
    x = rng.normal(0, 3e4, 200_000)
    for s in (1000., 1.e4, 3., 2.5e3):
        got = L._cabi.to_host(L._cabi.true_div(L._cabi.to_dev(x), s))
        assert np.array_equal(got, x / s)


def test_solve_is_a_drop_in_for_sparseqr_solve(L):
    """lssurf_b200.solve(A, b) at the one-line seam of the reference (sparseqr.solve, smooth_fit.py:142; aniso notebook cells
    6, 13, 20; BASELINE.json configs[0..1]): the systems are bit-identical to the reference's, the solutions agree with its
    direct solve to 1e-6 relative L2 (north-star tolerance)."""
    from tests import notebook_cases
    from tests.util import GOLDEN
    fx = np.load(f'{GOLDEN}/notebook_solves.npz')
    systems = notebook_cases.build(L, lambda d: L.point_data.data().from_dict(d))
    for name, (A, b, shape) in systems.items():
        assert csr_digest(A) == str(fx[name + '_A']) and digest(np.asarray(b, dtype=float)) == str(fx[name + '_b']), name
        info = {}
        # the directional operator of cell 13 smooths along one direction only: Jacobi-preconditioned CG needs ~3e5 iterations
        x = L.solve(A, b, info=info, maxiter=1_000_000, rtol=1e-10 if name == 'cfg2_aniso' else 1e-12)
        step = int(fx[name + '_step'])
        got, want = x.reshape(shape)[::step, ::step], fx[name + '_x']
        rel = float(np.linalg.norm(got - want) / np.linalg.norm(want))
        print(name, info, 'rel-L2 %.2e' % rel)
        assert info['converged'] and rel < 1e-6, (name, rel, info)
        assert abs(np.linalg.norm(x) - float(fx[name + '_norm'])) < 1e-6 * float(fx[name + '_norm'])
    # other accepted inputs: COO, a lin_op; a wrong right-hand side length is refused
    A, b, shape = systems['cfg1b']
    assert rel_l2(L.solve(A.tocoo(), b), L.solve(A, b)) < 1e-9
    with pytest.raises(ValueError):
        L.solve(A, b[:-1])
