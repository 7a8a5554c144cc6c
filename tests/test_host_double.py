"""CPU tier: the package's HOST logic (lin_op bookkeeping, smooth_fit assembly, edit loop control flow, parse_model)
driven end to end against the reference fixtures, with the C-ABI wrappers replaced by the test double in
tests/host_double.py (oracle arithmetic on CPU tensors).  The kernels are NOT under test here -- that is the
`-m gpu` tier's job, through the real library."""
import os

import numpy as np
import pytest

import lssurf_b200 as L
from tests import host_double
from tests.util import FIT_CASES, GOLDEN, WIDENED_CASES, check_parse_model_outputs, csr_digest, digest, load_case, rel_l2


@pytest.fixture()
def double():
    with host_double.installed():
        yield


def make_data(pts):
    return L.point_data.data().from_dict({k: np.array(v) for k, v in pts.items()})


@pytest.mark.parametrize('name', FIT_CASES + WIDENED_CASES)
def test_host_logic_against_reference_fixture(double, name):
    fx, pts, kw = load_case(name)
    fo = L.smooth_fit(data=make_data(pts), VERBOSE=False, return_fit_objects=True, **kw)
    G, C = fo['G_data'].toCSR(), fo['Gc'].toCSR()
    assert G.shape == tuple(fx['G_shape']) and C.shape == tuple(fx['C_shape'])
    assert csr_digest(G) == str(fx['G_digest']) and csr_digest(C) == str(fx['C_digest'])
    assert digest(np.asarray(fo['Ec'], dtype=float)) == str(fx['Ec_digest'])
    want_toc = {s.split(':')[0]: int(s.split(':')[1]) for s in fx['toc_rows']}
    assert {k: len(np.unique(v)) for k, v in fo['Gc'].TOC['rows'].items()} == want_toc
    hist = []
    S = L.smooth_fit(data=make_data(pts), VERBOSE=False, history=hist, sort_points=False, **kw)
    assert len(S['timing']['pcg']) == int(fx['n_outer'])
    for k, packed in enumerate(fx['masks']):
        assert np.array_equal(hist[k].get('mask_next', hist[k]['mask_out']), packed), f'mask after iteration {k}'
    assert np.array_equal(np.packbits(S['data'].three_sigma_edit), fx['in_TSE'])
    assert rel_l2(S['m']['all'], fx['m0']) < 1e-8
    assert np.max(np.abs(S['data'].z_est - fx['z_est'])) < 1e-6
    assert S['m']['z0'].z0.shape == tuple(S['grids']['z0'].shape)
    check_widened_outputs(fx, S)
    check_parse_model_outputs(fx, S)


def check_widened_outputs(fx, S):
    """Outputs only the section-8(f) fixtures hold: column ranges, sensor bias grids, slope biases, per-point sigma_extra."""
    if 'toc_cols' in fx.files:
        want = {s.split(':')[0]: (int(s.split(':')[1]), int(s.split(':')[2])) for s in fx['toc_cols']}
        got = {k: (int(np.min(v)), int(np.max(v))) for k, v in S['TOC']['cols'].items()}      # Gc's table; the fixture holds G_data's
        assert all(got[k] == v for k, v in want.items() if k in got) and 'z0' in got
    for key in (k for k in fx.files if k.startswith('grid_')):
        got = S['m']['sensor_bias_grids'][key[5:]].z
        assert got.shape == fx[key].shape and np.max(np.abs(got - fx[key])) < 1e-6 * max(1., np.max(np.abs(fx[key])))
    if 'slope_bias' in fx.files:
        for sensor, sx, sy in fx['slope_bias']:
            got = S['m']['slope_bias'][sensor]
            assert abs(got['slope_x'] - sx) < 1e-6 * max(1., abs(sx)) and abs(got['slope_y'] - sy) < 1e-6 * max(1., abs(sy))
    for key in (k for k in fx.files if k.startswith('avgop_')):       # dz/dt and block-average grids
        name = key[6:]
        got, want = np.asarray(getattr(S['m'][name], name)), fx[key]
        assert got.shape == want.shape and np.array_equal(np.isnan(got), np.isnan(want))
        assert np.nanmax(np.abs(got - want)) < 1e-6 * max(1., np.nanmax(np.abs(want))), name
        if 'avgarea_' + name in fx.files:
            assert np.allclose(S['m'][name].cell_area, fx['avgarea_' + name], rtol=1e-12, atol=0), name
    if 'sigma_extra_final' in fx.files:
        assert np.max(np.abs(S['data'].sigma_extra - fx['sigma_extra_final'])) < 2e-5
    if 'lagrangian_dz' in fx.files:
        got, want = np.asarray(S['m']['lagrangian_dz'].dz), fx['lagrangian_dz']
        assert got.shape == want.shape and np.max(np.abs(got - want)) < 1e-6 * max(1., np.max(np.abs(want)))
    if 'gridbias_values' in fx.files:
        name = str(fx['gridbias_param']) + '_bias'
        got = np.asarray(getattr(S['m']['grid_biases'][name], name)).ravel()
        assert got.shape == fx['gridbias_values'].shape and np.max(np.abs(got - fx['gridbias_values'])) < 1e-6 * max(1., np.max(np.abs(fx['gridbias_values'])))


def check_compute_E(Lp, **extra):
    """compute_E against the reference run (oracle/make_golden.py::case_compute_E; smooth_fit.py:212-274, 689-712,
    lin_op.py:540-553).  The reference reads errors off the rows of R^-1 with entries below 1e-5 dropped, the package off
    the exact covariance N^-1: at most n x 1e-10 (about 1e-7) of variance is missing from a reference value, and forming
    N squares the condition number of the QR; the bar is 1e-4 relative."""
    fx, pts, kw = load_case('compute_E')
    S = Lp.smooth_fit(data=make_data(pts), VERBOSE=False, **dict(kw, **extra))
    assert np.array_equal(np.packbits(S['data'].three_sigma_edit), fx['in_TSE'])
    assert rel_l2(S['m']['all'], fx['m0']) < 1e-6
    keys = [k[2:] for k in fx.files if k.startswith('E_') and k != 'E_sigma_bias_val']
    assert set(keys) <= set(S['E']) and len(keys) == 7, (keys, sorted(S['E']))
    for key in keys:
        want, got = fx['E_' + key], np.asarray(getattr(S['E'][key], key), dtype=float)
        assert got.shape == want.shape and np.array_equal(np.isnan(got), np.isnan(want)), key
        ok = ~np.isnan(want)
        assert np.max(np.abs(got[ok] - want[ok]) / np.maximum(want[ok], 1e-3)) < 1e-4, key
    want, got = fx['E_sigma_bias_val'], np.asarray(S['E']['sigma_bias']['val'], dtype=float)
    assert got.shape == want.shape and np.max(np.abs(got - want) / want) < 1e-4
    return S


def test_compute_E_against_reference_fixture(double):
    check_compute_E(L, sort_points=False)


def test_mask_update_function_sees_the_parsed_model(double):
    """mask_update_function (smooth_fit.py:664-666): called once after the edit loop with (grids, m, args), m already parsed
    (no averaging grids yet); what it does to the grids is seen by the averaging operators built afterwards."""
    fx, pts, kw = load_case('uniform_avg_scales')
    seen = {}

    def hook(grids, m, args):
        seen['keys'] = sorted(m)
        seen['dz_shape'] = m['dz'].dz.shape
        seen['n_calls'] = seen.get('n_calls', 0) + 1
        grids['dz'].cell_area = grids['dz'].cell_area * 2.          # doubles the areas the block averages report
    S = L.smooth_fit(data=make_data(pts), VERBOSE=False, sort_points=False, mask_update_function=hook, **kw)
    assert seen['n_calls'] == 1 and 'z0' in seen['keys'] and 'dz' in seen['keys'] and not any(k.startswith('avg_') for k in seen['keys'])
    assert seen['dz_shape'] == tuple(S['grids']['dz'].shape)
    assert rel_l2(S['m']['all'], fx['m0']) < 1e-6
    name = [k for k in fx.files if k.startswith('avgarea_avg_dz_')][0][len('avgarea_'):]
    assert np.allclose(S['m'][name].cell_area, 2. * fx['avgarea_' + name], rtol=1e-12)


def test_public_mask_functions_drop_points_in_place(double):
    """validate_by_dz_mask / setup_mask as public functions (LSsurf/__init__.py exports both; grid_functions.py:327-374): the
    reference's in-place contract -- data subset, valid_data updated -- and its thresholds (> 0.5 of the interpolated dz mask,
    != 0 of the interpolated z0 mask)."""
    fx, pts, kw = load_case('uniform_masked')
    args = dict(L.smooth_fit.__globals__['DEFAULTS'], **kw)
    grids, bds = L.setup_grids(args)
    for fn, extra in ((L.validate_by_dz_mask, ()), (L.setup_mask, (bds, args))):
        D = make_data(pts)
        valid = np.ones(D.size, dtype=bool)
        fn(D, grids, valid, *extra)
        assert 0 < valid.sum() < valid.size and D.size == int(valid.sum())
        assert np.array_equal(np.asarray(D.x), pts['x'][valid])
        g = grids['dz'] if fn is L.validate_by_dz_mask else grids['z0']
        # a point whose four surrounding nodes are all inside the mask is kept, one whose four nodes are all outside is dropped
        iy = np.clip(np.floor((pts['y'] - g.bds[0][0]) / g.delta[0]).astype(int), 0, g.shape[0] - 2)
        ix = np.clip(np.floor((pts['x'] - g.bds[1][0]) / g.delta[1]).astype(int), 0, g.shape[1] - 2)
        corners = np.stack([g.mask[iy, ix], g.mask[iy + 1, ix], g.mask[iy, ix + 1], g.mask[iy + 1, ix + 1]])
        assert np.all(valid[corners.all(axis=0)]) and not np.any(valid[~corners.any(axis=0)])


def check_errors_and_early_returns(Lp, **extra):
    """The failure handling of the reference driver: ValueErrors for a missing required keyword
    (smooth_fit.py:403-405), zero data sigma (:604-605), a zero expected constraint value (constraint_functions.py:107-109) and a
    zero expected bias (bias_functions.py:112-113); points outside the grids or with a non-finite z dropped silently
    (:406, 423-425); the early return with data None and empty TOC / R / RMS when nothing is left (:427-430)."""
    fx, pts, kw = load_case('tracks_bias_small')
    with pytest.raises(ValueError):
        Lp.smooth_fit(data=make_data(pts), VERBOSE=False, **extra, **{k: v for k, v in kw.items() if k != 'spacing'})
    with pytest.raises(ValueError, match='zero value found in data sigma'):
        Lp.smooth_fit(data=make_data(dict(pts, sigma=np.where(np.abs(pts['time']) < 1.0, 0., pts['sigma']))), VERBOSE=False, **extra, **kw)
    with pytest.raises(ValueError, match='found zero value in the expected values for d2z_dt2'):
        Lp.smooth_fit(data=make_data(pts), VERBOSE=False, **extra, **dict(kw, E_RMS=dict(kw['E_RMS'], d2z_dt2=0.)))
    with pytest.raises(ValueError, match='zero value in the expected biases'):
        Lp.smooth_fit(data=make_data(dict(pts, sigma_corr=np.zeros_like(pts['sigma_corr']))), VERBOSE=False, **extra, **kw)
    # points outside the domain / NaN heights: dropped, reported through valid_data, the rest fitted as if they were never there
    far = dict(pts)
    far['x'] = pts['x'].copy(); far['z'] = pts['z'].copy()
    far['x'][:5] += 1.e6
    far['z'][5:9] = np.nan
    S = Lp.smooth_fit(data=make_data(far), VERBOSE=False, **extra, max_iterations=1, **{k: v for k, v in kw.items() if k != 'max_iterations'})
    assert S['valid_data'].shape == (pts['x'].size,)
    assert not S['valid_data'][:9].any() and S['data'].size <= pts['x'].size - 9
    # nothing valid: the early return
    gone = dict(pts, x=pts['x'] + 1.e6)
    E0 = Lp.smooth_fit(data=make_data(gone), VERBOSE=False, **extra, **kw)
    assert E0['data'] is None and E0['TOC'] == {} and E0['R'] == {} and E0['RMS'] == {} and E0['m'] == {} and not E0['valid_data'].any()
    assert set(E0) >= {'m', 'E', 'data', 'grids', 'valid_data', 'TOC', 'R', 'RMS', 'timing', 'E_RMS'}


def test_errors_and_early_returns_like_the_reference(double):
    check_errors_and_early_returns(L, sort_points=False)


def test_constraints_that_leave_out_the_dz_columns_fail_like_the_reference(double):
    """No dz constraint at all: the constraint operator ends at the last z0 column, and the reference fails when it stacks
    G_data on Gc (smooth_fit.py:621, scipy: 'incompatible dimensions for axis 1').  Found by tools/diff_fuzz.py (seed 777)."""
    fx, pts, kw = load_case('uniform3d_small')
    E = {k: v for k, v in kw['E_RMS'].items() if k in ('d2z0_dx2', 'dz0_dx')}
    with pytest.raises(ValueError, match='incompatible dimensions for axis 1'):
        L.smooth_fit(data=make_data(pts), VERBOSE=False, sort_points=False, **dict(kw, E_RMS=E))


def test_no_iterations_returns_the_validity_mask(double):
    """max_iterations = 0 (smooth_fit.py:648-662 skipped): empty m / E / R / RMS, the operators' TOC, and valid_data as the
    validity mask of the points.  Found by tools/diff_fuzz.py ('more' mode)."""
    fx, pts, kw = load_case('uniform3d_small')
    far = dict(pts, x=pts['x'].copy())
    far['x'][:7] += 1.e6
    for p_, n_valid in ((pts, pts['x'].size), (far, pts['x'].size - 7)):
        S = L.smooth_fit(data=make_data(p_), VERBOSE=False, sort_points=False, **dict(kw, max_iterations=0))
        assert S['m'] == {} and S['E'] == {} and S['R'] == {} and S['RMS'] == {} and 'cols' in S['TOC']
        assert S['valid_data'].dtype == bool and S['valid_data'].shape == (pts['x'].size,) and S['valid_data'].sum() == n_valid


def test_double_is_removed_afterwards():
    from lssurf_b200 import _cabi
    assert _cabi.interp_coo.__module__ == 'lssurf_b200._cabi' and _cabi.device.__module__ == 'lssurf_b200._cabi'


def test_bias_filter_groups_like_the_filtered_parameters(double):
    """bias_filter (bias_functions.py:53-57): a filter that blanks `cycle` gives one bias per rgt -- the same operator
    as a fit whose data carry a constant cycle."""
    fx, pts, kw = load_case('tracks_bias_small')

    def per_track(D):
        out = D.copy()
        out.assign({'cycle': np.zeros_like(np.asarray(D.cycle))})
        return out
    a = L.smooth_fit(data=make_data(pts), VERBOSE=False, return_fit_objects=True, bias_filter=per_track, **kw)
    b = L.smooth_fit(data=make_data(dict(pts, cycle=np.zeros_like(pts['cycle']))), VERBOSE=False, return_fit_objects=True, **kw)
    assert csr_digest(a['G_data'].toCSR()) == csr_digest(b['G_data'].toCSR())
    assert csr_digest(a['Gc'].toCSR()) == csr_digest(b['Gc'].toCSR()) and np.array_equal(a['Ec'], b['Ec'])
    assert a['G_data'].shape[1] < int(fx['G_shape'][1])          # fewer bias columns than one per (rgt, cycle)


@pytest.mark.parametrize('name', ['tracks_bias_small', 'tracks_sensor_slope', 'tracks_bias_edit', 'uniform_groups_editable', 'uniform_sigma_grid'])
def test_two_rank_row_sharding_reproduces_the_fixture(name, tmp_path):
    """World size 2 over gloo, host double on each rank (tests/two_rank_worker.py): every rank assembles only its block of
    data rows (bias, slope and sensor-grid rows included), the replicated solve sees the gathered rows, per-point outputs
    come back in the caller's order -- and the result is the single-process reference fixture on both ranks."""
    import os
    import subprocess
    import sys
    port = str(31500 + (os.getpid() % 2000))
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'two_rank_worker.py')
    outs = [str(tmp_path / f'rank{r}.npz') for r in range(2)]
    procs = [subprocess.Popen([sys.executable, worker, str(r), '2', port, name, outs[r]], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    logs = []
    for p in procs:
        try:
            logs.append(p.communicate(timeout=300)[0].decode(errors='replace'))
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    assert all(p.returncode == 0 for p in procs), '\n'.join(logs)
    fx, pts, kw = load_case(name)
    res = [np.load(o) for o in outs]
    for r in res:
        assert rel_l2(r['m0'], fx['m0']) < 1e-8 and np.array_equal(r['in_TSE'], fx['in_TSE'])
        assert r['masks'].shape[0] == fx['masks'].shape[0]      # same number of edit iterations (per-iteration masks are per rank)
        assert np.max(np.abs(r['z_est'] - fx['z_est'])) < 1e-6 and np.nansum(r['count']) > 0
        if r['n_edited'].size:                                   # per-bias counts of edited-in points add up to the mask (all-reduced)
            assert r['n_edited'].sum() == int(np.unpackbits(fx['in_TSE'])[:int(fx['n_data'])].sum())
    assert np.array_equal(res[0]['m0'], res[1]['m0'])          # both ranks hold the same model, bit for bit


def test_two_rank_outputs_on_rank_0_only(tmp_path):
    """gather_outputs='root': rank 0 returns what a single process returns, rank 1 the same model and node statistics but no
    per-point outputs (it takes part in the gathers and skips the copies to the host)."""
    import subprocess
    import sys
    name = 'tracks_bias_small'
    port = str(33500 + (os.getpid() % 2000))
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'two_rank_worker.py')
    outs = [str(tmp_path / f'rank{r}.npz') for r in range(2)]
    procs = [subprocess.Popen([sys.executable, worker, str(r), '2', port, name, outs[r], 'root'], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    logs = []
    for p in procs:
        try:
            logs.append(p.communicate(timeout=300)[0].decode(errors='replace'))
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    assert all(p.returncode == 0 for p in procs), '\n'.join(logs)
    fx, pts, kw = load_case(name)
    r0, r1 = (np.load(o) for o in outs)
    assert rel_l2(r0['m0'], fx['m0']) < 1e-8 and np.array_equal(r0['in_TSE'], fx['in_TSE'])
    assert np.max(np.abs(r0['z_est'] - fx['z_est'])) < 1e-6
    assert np.array_equal(r0['m0'], r1['m0']) and np.array_equal(r0['count'], r1['count'], equal_nan=True)
    assert 'z_est' not in r1.files and r1['masks'].shape[0] == fx['masks'].shape[0]


def grid_bias_operators(Lp):
    """The operator-level set-up of the grid_bias fixture (oracle/make_golden.py::case_grid_bias) with the package."""
    fx = np.load(os.path.join(GOLDEN, 'grid_bias.npz'), allow_pickle=False)
    pts = {k[3:]: fx[k] for k in fx.files if k.startswith('in_')}
    D = Lp.point_data.data().from_dict(pts)
    z0 = Lp.fd_grid([[-6e3, 6e3], [-6e3, 6e3]], [500., 500.], name='z0')
    dz = Lp.fd_grid([[-6e3, 6e3], [-6e3, 6e3], [-1.5, 1.5]], [1000., 1000., 0.25], name='dz', col_0=z0.N_nodes)
    grids = {'z0': z0, 'dz': dz}
    G = Lp.lin_op(z0, name='interp_z').interp_mtx(D.coords()[0:2])
    G.add(Lp.lin_op(dz, name='interp_dz').interp_mtx(D.coords()))
    ops = []
    bg = Lp.fd_grid([[-6e3, 6e3], [-6e3, 6e3]], [2000., 2000.], name='cov_bias')
    Lp.setup_grid_bias(D, G, ops, grids, param='cov', grid=bg, expected_rms=0.1, expected_value=0., expected_rms_grad=1e-3)
    Gc = Lp.lin_op(None, name='constraints').vstack(ops)
    return fx, G, Gc, ops, bg, grids


def check_grid_bias(fx, G, Gc, ops, bg, grids):
    assert csr_digest(G.toCSR()) == str(fx['G_digest']) and G.toCSR().shape == tuple(fx['G_shape'])
    assert csr_digest(Gc.toCSR()) == str(fx['C_digest']) and Gc.toCSR().shape == tuple(fx['C_shape'])
    assert np.array_equal(np.concatenate([op.expected for op in ops]), fx['expected'])
    assert [op.name for op in ops] == [str(s) for s in fx['names']] and [bg.col_0, bg.col_N] == fx['col_range'].tolist()
    assert {k: len(np.unique(v)) for k, v in Gc.TOC['rows'].items()} == {s.split(':')[0]: int(s.split(':')[1]) for s in fx['toc_rows']}
    assert grids['cov_bias'] is bg


def test_setup_grid_bias_operators(double):
    """setup_grid_bias (setup_grid_bias.py:4-43) at the operator level against the reference digests."""
    check_grid_bias(*grid_bias_operators(L))


def iterate_fit_like_a_reference_caller(Lp, name='uniform3d_small'):
    """The narrow seam (smooth_fit.py:100-210) driven the way the reference's smooth_fit drives it: operators from
    return_fit_objects, NumPy rhs / TCinv / Ip_c built by the caller, NumPy arrays back."""
    import scipy.sparse as sp
    from lssurf_b200.constraint_functions import build_reference_epoch_matrix
    from lssurf_b200.smooth_fit import DEFAULTS
    fx, pts, kw = load_case(name)
    fo = Lp.smooth_fit(data=Lp.point_data.data().from_dict(pts), VERBOSE=False, return_fit_objects=True, **kw)
    data, G_data, Gc, grids = fo['data'], fo['G_data'], fo['Gc'], fo['grids']
    N_eq = G_data.N_eq + Gc.N_eq
    TCinv = sp.dia_matrix((1. / np.concatenate((fo['Ed'], fo['Ec'])), 0), shape=(N_eq, N_eq))       # smooth_fit.py:613
    rhs = np.zeros(N_eq)
    rhs[0:data.size] = data.z.ravel()                                                                # :616-617
    Ip_c = build_reference_epoch_matrix(G_data, Gc, grids, kw['reference_epoch'])
    args = dict(DEFAULTS, **kw)
    args.update(VERBOSE=False)
    timing = {}
    m0, sigma_extra, in_TSE, rs_data = Lp.iterate_fit(data, None, rhs, TCinv, G_data, Gc, np.ones(data.size, dtype=bool), Ip_c,
                                                      timing, args, grids)
    assert isinstance(m0, np.ndarray) and isinstance(sigma_extra, np.ndarray) and in_TSE.dtype == bool
    assert rs_data.shape == (data.size,) and sigma_extra.shape == (data.size,)
    assert np.array_equal(np.packbits(in_TSE), fx['in_TSE']) and rel_l2(m0, fx['m0']) < 1e-6
    assert abs(np.median(sigma_extra) - fx['sigma_extra'][-1]) < 2e-5
    r = (data.z - G_data.toCSR().dot(m0)) / data.sigma                                               # :146-147
    assert np.max(np.abs(rs_data - r)) < 1e-9
    assert 'sparseqr_solve' in timing


def test_iterate_fit_reference_signature(double):
    iterate_fit_like_a_reference_caller(L)


def test_notebook_systems_are_bit_identical(double):
    """BASELINE.json configs[0..1] at the low-level API (fd_grid / lin_op.interp_mtx / grad2 / one / diff_op with mutated .v /
    add / vstack / toCSR, notebooks/smooth_fit_demo_aniso.ipynb): the scaled systems handed to solve(A, b) have the digests of
    the ones the unmodified reference builds (tests/notebook_cases.py, oracle/make_golden.py::case_notebook_solves)."""
    from tests import notebook_cases
    fx = np.load(f'{GOLDEN}/notebook_solves.npz')
    systems = notebook_cases.build(L, make_data)
    assert set(systems) == {'cfg1b', 'cfg2_iso', 'cfg2_aniso', 'cfg2_circle'}
    for name, (A, b, shape) in systems.items():
        assert csr_digest(A) == str(fx[name + '_A']), name
        assert digest(np.asarray(b, dtype=float)) == str(fx[name + '_b']), name
