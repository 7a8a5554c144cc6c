"""Differential check of the package's HOST logic against the unmodified reference on small random configurations
(build container only: needs /root/reference; the kernels are replaced by tests/host_double.py, so this exercises the
Python mirror of the reference interface -- keyword handling, operator bookkeeping, edit-loop control flow).
    python tools/diff_fuzz.py [n_cases=12] [seed=0] [widened] [compute_E] [more]
With ``widened`` every case also draws from the keywords that take in-memory grids: NumPy masks on both grids (``mask_data``,
``mask_scale``), ``constraint_scaling_maps``, ``avg_masks``, ``prior_args``, ``DEM_tol``, ``lagrangian_coords``; the reference
is driven through the hooks oracle/make_golden.py uses for its fixtures (masks put on the reference's grids, node arrays
wrapped in an object that can "interpolate" itself, prior points handed to the reference's make_prior_op).  With ``compute_E``
half of the small cases also propagate errors (dense stand-ins for sparseqr.rz / inv_tr_upper on the reference side)."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H
from tests import host_double
from tests.util import csr_digest
import lssurf_b200 as L


def one_case(rng, k):
    n = int(rng.integers(300, 2500))
    W = float(rng.choice([6e3, 8e3, 1.2e4]))
    t_span = float(rng.choice([1., 2., 3.]))
    x, y, t = rng.uniform(-W / 2 * 1.05, W / 2 * 1.05, n), rng.uniform(-W / 2, W / 2, n), rng.uniform(-t_span / 2 * 1.1, t_span / 2 * 1.1, n)
    z = 30 * np.sin(x / 3000.) * np.cos(y / 2500.) + t * (1 + x / W) + rng.normal(0, 0.1, n)
    bad = rng.random(n) < 0.05
    z[bad] += rng.normal(0, 10, int(bad.sum()))
    if rng.random() < 0.3:
        z[rng.integers(0, n, 3)] = np.nan
    pts = dict(x=x, y=y, time=t, z=z, sigma=rng.choice([0.1, 0.2], n))
    E = {'d2z0_dx2': 0.02, 'dz0_dx': 50, 'd3z_dx2dt': 3e-4, 'd2z_dxdt': 0.75, 'd2z_dt2': 5000}
    for key in list(E):
        if rng.random() < 0.2:
            E[key] = None if rng.random() < 0.5 else E[key] * 3
    dt = float(rng.choice([0.25, 0.5]))
    n_ep = int(round(t_span / dt))
    kw = dict(W={'x': W, 'y': W, 't': t_span}, ctr={'x': 0., 'y': 0., 't': 0.},
              spacing={'z0': float(rng.choice([250., 500.])), 'dz': 1000., 'dt': dt}, E_RMS=E,
              reference_epoch=int(rng.integers(0, n_ep + 1)), max_iterations=int(rng.choice([1, 2, 4, 10])))
    if rng.random() < 0.5:
        pts.update(rgt=rng.integers(0, 6, n).astype(float), cycle=rng.integers(0, 3, n).astype(float), sigma_corr=np.full(n, 0.05))
        if rng.random() < 0.3:
            pts['cycle'][rng.integers(0, n, 5)] = np.nan
        kw['bias_params'] = ['rgt', 'cycle']
    if rng.random() < 0.3:
        kw['dzdt_lags'] = [1, 2]
    if rng.random() < 0.3:
        kw['sigma_extra_relax'] = True
    if rng.random() < 0.3:
        kw['min_iterations'] = int(rng.integers(0, 4))
        kw['converge_tol_dz'] = float(rng.choice([0.01, 0.05, 1.0]))
    # SURVEY section 8(f) rows and edit-loop options
    if rng.random() < 0.6:
        pts['sensor'] = (rng.random(n) < 0.4).astype(float)
        if rng.random() < 0.5:
            kw['sensor_grid_bias_params'] = [{'sensor': 1., 'spacing': float(rng.choice([2000., 3000.])), 'expected_rms': 0.2,
                                              'expected_rms_grad2': 1e-7}]
        if 'bias_params' in kw and rng.random() < 0.5:
            kw['data_slope_sensors'] = np.array([1.])
        if rng.random() < 0.4:
            kw['sigma_extra_keys'] = {'a': {'sensor': [0.]}, 'b': {'sensor': [1.]}}
    if rng.random() < 0.3:
        pts['editable'] = (rng.random(n) < 0.9).astype(float)
    if rng.random() < 0.3:
        pts['three_sigma_edit'] = (rng.random(n) < 0.97).astype(float)
    if rng.random() < 0.3 and 'sigma_extra_keys' not in kw:
        kw['sigma_extra_bin_spacing'] = float(rng.choice([2000., 4000.]))
        if rng.random() < 0.5:
            kw['sigma_extra_max'] = 1.0
    if 'dzdt_lags' in kw and rng.random() < 0.6:
        kw['avg_scales'] = [2000., 4000.] if rng.random() < 0.5 else [2000.]
        if rng.random() < 0.5:
            kw['z0_average_scale'] = 1000.
    if 'bias_params' in kw and rng.random() < 0.3:
        kw['bias_nsigma_edit'] = float(rng.choice([3., 6.]))
        kw['bias_nsigma_iteration'] = int(rng.integers(0, 3))
    if MORE:
        # less common switches: a magnitude constraint on z0, pre-edited biases, the fractional-change stop rule, no
        # iterations at all, long dz/dt lags, block averages without dz/dt lags
        if rng.random() < 0.3 and E.get('d2z0_dx2') is not None:
            E['z0'] = float(rng.choice([10., 100.]))
        if 'bias_params' in kw and rng.random() < 0.3:
            pts.setdefault('three_sigma_edit', np.ones(n))
            kw['bias_edit_vals'] = {'rgt': [float(rng.integers(0, 6))], 'cycle': [float(rng.integers(0, 3))]}
        if rng.random() < 0.3:
            kw['converge_tol_frac_TSE'] = float(rng.choice([1e-3, 1e-2, 0.1]))
        if rng.random() < 0.1:
            kw['max_iterations'] = 0
        if rng.random() < 0.2:
            kw['dzdt_lags'] = [1, int(rng.integers(2, 8))]
        if rng.random() < 0.15 and 'dzdt_lags' not in kw:
            kw['avg_scales'] = [2000.]
        if rng.random() < 0.2:
            kw['sigma_extra_max'] = float(rng.choice([0.05, 1.0]))
    if COMPUTE_E and W <= 8e3 and kw['spacing']['z0'] == 500. and 'sensor_grid_bias_params' not in kw and rng.random() < 0.5:
        kw['compute_E'] = True              # small enough for the dense QR / triangular-inverse stand-ins of the reference run
        if kw['max_iterations'] == 0 and rng.random() < 0.7:
            pts['sigma_extra'] = rng.choice([0., 0.05], n)      # errors without a fit need the data to carry sigma_extra
    return pts, kw


COMPUTE_E = 'compute_E' in sys.argv[1:]
MORE = 'more' in sys.argv[1:]


def widen(rng, pts, kw):
    """Add keywords that take in-memory grids / points to a case of one_case (package form of the arguments)."""
    W = kw['W']['x']
    n = pts['x'].size
    nz0, ndz = int(W / kw['spacing']['z0']) + 1, int(W / kw['spacing']['dz']) + 1
    for key in ('sigma_extra_bin_spacing', 'sigma_extra_max', 'sensor_grid_bias_params', 'data_slope_sensors'):
        kw.pop(key, None)                    # keep the combinations to what the fixtures establish as supported together
    if rng.random() < 0.5:
        mz, md = np.ones((nz0, nz0), dtype=bool), np.ones((ndz, ndz), dtype=bool)
        for m in (mz, md):
            for _ in range(int(rng.integers(0, 3))):
                a, b = sorted(rng.integers(0, m.shape[0], 2))
                c, d = sorted(rng.integers(0, m.shape[1], 2))
                m[a:min(b, a + m.shape[0] // 3) + 1, c:min(d, c + m.shape[1] // 3) + 1] = False
        kw['mask_data'] = {'z0': mz, 'dz': md}
        if rng.random() < 0.5:
            kw['mask_scale'] = {0: float(rng.choice([3., 10., 30.])), 1: 1}
    if rng.random() < 0.4:
        maps = {}
        if kw['E_RMS'].get('d2z0_dx2') is not None and rng.random() < 0.7:
            m = np.ones((nz0, nz0)); m[:nz0 // 2] = float(rng.choice([0.25, 4.])); m[int(rng.integers(0, nz0)), int(rng.integers(0, nz0))] = np.nan
            maps['d2z0_dx2'] = m
        if kw['E_RMS'].get('d3z_dx2dt') is not None and rng.random() < 0.7:
            m = np.ones((ndz, ndz)); m[:, ndz // 2:] = float(rng.choice([0.5, 4.]))
            maps['d3z_dx2dt'] = m
        if kw['E_RMS'].get('d2z_dt2') is not None and rng.random() < 0.5:
            m = np.ones((ndz, ndz)); m[:3, :3] = 0.5
            maps['d2z_dt2'] = m
        if rng.random() < 0.4:
            m = np.zeros((ndz, ndz)); m[:2, :] = 0.05; m[min(5, ndz - 1), min(5, ndz - 1)] = np.nan
            maps['dz'] = m
        if maps:
            kw['constraint_scaling_maps'] = maps
    if rng.random() < 0.4:
        kw.setdefault('dzdt_lags', [1, 2])
        iy, ix = np.meshgrid(np.arange(ndz), np.arange(ndz), indexing='ij')
        basin = np.clip(1.2 - np.hypot(iy - ndz / 3., ix - ndz / 2.) / (ndz / 3.), 0., 1.)
        if rng.random() < 0.5:
            basin[:, :2] = np.nan
        kw['avg_masks'] = {'basin': basin, 'shelf': (ix >= ndz // 2).astype(float)}
    if rng.random() < 0.4:
        npr = int(rng.integers(20, 120))
        px, py = rng.uniform(-0.45 * W, 0.45 * W, npr), rng.uniform(-0.45 * W, 0.45 * W, npr)
        half_t = kw['W']['t'] / 2
        pt = rng.choice([-half_t / 2, 0., half_t / 2, half_t * 0.9], npr)
        prior = dict(x=px, y=py, t=pt, dz=pt * (1. + px / W) + rng.normal(0, 0.02, npr), sigma_dz=np.full(npr, 0.05))
        kw['prior_args'] = dict(dzs=[prior], ref_epoch=0., sigma_scale=float(rng.choice([1., 2.])))
    if rng.random() < 0.3:
        dem = 30 * np.sin(pts['x'] / 3000.) * np.cos(pts['y'] / 2500.) + rng.normal(0, 0.5, n)
        dem[rng.random(n) < 0.02] += 40.
        pts['DEM'] = dem
        kw['DEM_tol'] = 10.
    if rng.random() < 0.3 and 'bias_params' in kw and 'bias_nsigma_edit' not in kw:
        # a gridded bias that multiplies a per-point covariate (setup_grid_bias.py:4-43; cfg 4 of the benchmark)
        pts['cov'] = np.sin(pts['x'] / 2500.) + 0.1 * rng.standard_normal(n)
        pts['z'] = pts['z'] + pts['cov'] * 0.15 * np.cos(pts['y'] / 4000.)
        kw['grid_bias_spec'] = dict(param='cov', bounds=(-W / 2, W / 2, -W / 2, W / 2), spacing=float(rng.choice([2000., 3000.])),
                                    expected_rms=0.1, expected_rms_grad=1e-3)
    if rng.random() < 0.3 and 'bias_nsigma_edit' not in kw:
        has = rng.random(n) < 0.66
        pts['x0'] = np.where(has, pts['x'] - 150. * pts['time'], np.nan)
        pts['y0'] = np.where(has, pts['y'] + 60. * pts['time'], np.nan)
        kw['lagrangian_coords'] = ['y0', 'x0']
        kw['spacing'] = dict(kw['spacing'], lagrangian_dz=1000.)
        kw['E_RMS'] = dict(kw['E_RMS'], lagrangian_dz=50., lagrangian_dzdx=0.5)
    return pts, kw


def _close(a, b):
    """scalars equal to 1e-6 relative; NaN (no data left) matches NaN"""
    return (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-6 * abs(a) + 1e-12


def _same_E(key, g, E_p):
    """one entry of the error dict: the reference drops entries of R^-1 below 1e-5, the bar is 1e-4 relative"""
    if key in ('sigma_bias', 'sigma_slope_bias'):
        a, b = np.asarray(g.get('val', []), dtype=float), np.asarray(E_p.get(key, {}).get('val', []), dtype=float)
    else:
        a, b = np.asarray(getattr(g, key), dtype=float), np.asarray(getattr(E_p[key], key), dtype=float) if key in E_p else np.zeros(0)
    return bool(a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and
                (not np.any(~np.isnan(a)) or np.nanmax(np.abs(a - b) / np.maximum(np.abs(a), 1e-3)) < 1e-4))


class _NodeGrid(object):
    """What the reference asks of a pointCollection grid that already sits on the nodes it is sampled at."""

    def __init__(self, z):
        self.z = z

    def interp(self, x, y, gridded=False, **_):
        return self.z.copy()


def reference_fit(R, pts, kw, **more):
    """The unmodified reference with the package form of the widened keywords translated by the hooks of
    oracle/make_golden.py::case_widened."""
    from oracle import make_golden as MG
    SF = H.smooth_fit_module()
    kw = dict(kw)
    masks = kw.pop('mask_data', None)
    prior = kw.pop('prior_args', None)
    for key in ('avg_masks', 'constraint_scaling_maps'):
        if kw.get(key) is not None:
            kw[key] = {k: _NodeGrid(np.asarray(v, dtype=float)) for k, v in kw[key].items()}
    gb = kw.pop('grid_bias_spec', None)
    orig_constraints = SF.setup_smoothness_constraints
    if gb is not None:
        # the live reference driver has no call site for setup_grid_bias (smooth_fit.py:487-489); its predecessor called it right
        # after the smoothness constraints (deprecated/smooth_xytb_fit.py:465-467): same hook as oracle/make_golden.py
        def constraints_then_grid_bias(grids, constraint_op_list, *a, **k):
            orig_constraints(grids, constraint_op_list, *a, **k)
            caller = sys._getframe(1).f_locals
            b = gb['bounds']
            g = R.fd_grid([[b[0], b[1]], [b[2], b[3]]], [gb['spacing']] * 2, name=gb['param'] + '_bias')
            R.setup_grid_bias(caller['data'], caller['G_data'], constraint_op_list, grids, param=gb['param'], grid=g,
                              expected_rms=gb['expected_rms'], expected_value=0., expected_rms_grad=gb['expected_rms_grad'])
        SF.setup_smoothness_constraints = constraints_then_grid_bias
    orig_prior = SF.match_prior_dz
    if prior is not None:
        MP = sys.modules['LSsurf.match_priors']

        def prior_ops(grids, **_):
            return [MP.make_prior_op(grids, H.point_data(**d), 'prior_%d' % k, ref_time=prior['ref_epoch'], sigma_scale=prior['sigma_scale'])
                    for k, d in enumerate(prior['dzs'])]
        kw['prior_args'] = {}
        SF.match_prior_dz = prior_ops
    try:
        with MG.with_masks(SF, masks):
            return R.smooth_fit(data=H.point_data(**pts), VERBOSE=False, **more, **kw)
    finally:
        SF.match_prior_dz = orig_prior
        SF.setup_smoothness_constraints = orig_constraints


def package_kw(kw):
    """The keywords as the package takes them: the grid-bias spec becomes ``grid_bias_model_args`` with an fd_grid."""
    kw = dict(kw)
    gb = kw.pop('grid_bias_spec', None)
    if gb is not None:
        b = gb['bounds']
        kw['grid_bias_model_args'] = [dict(param=gb['param'], expected_rms=gb['expected_rms'], expected_value=0., expected_rms_grad=gb['expected_rms_grad'],
                                           grid=L.fd_grid([[b[0], b[1]], [b[2], b[3]]], [gb['spacing']] * 2, name=gb['param'] + '_bias'))]
    return kw


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    widened = 'widened' in sys.argv[3:]
    R = H.load()
    bad = 0
    for k in range(n_cases):
        pts, kw = one_case(rng, k)
        if widened:
            pts, kw = widen(rng, pts, kw)
        tag = 'case %d: n=%d %s' % (k, pts['x'].size, {a: (b if a not in ('mask_data', 'constraint_scaling_maps', 'avg_masks', 'prior_args') else sorted(b))
                                               for a, b in kw.items() if a not in ('W', 'ctr')})
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            try:
                fo_r = reference_fit(R, pts, kw, return_fit_objects=True)
                S_r = reference_fit(R, pts, kw)
                err_r = None
            except Exception as e:          # noqa: BLE001
                err_r = type(e).__name__
                msg_r = str(e)[:80]
            with host_double.installed():
                try:
                    fo_p = L.smooth_fit(data=L.point_data.data().from_dict(pts), VERBOSE=False, return_fit_objects=True, **package_kw(kw))
                    S_p = L.smooth_fit(data=L.point_data.data().from_dict(pts), VERBOSE=False, sort_points=False, **package_kw(kw))
                    Gp, Cp = fo_p['G_data'].toCSR(), fo_p['Gc'].toCSR()
                    err_p = None
                except Exception as e:      # noqa: BLE001
                    err_p = type(e).__name__
                    msg_p = str(e)[:80]
        if err_r or err_p:
            ok = err_r == err_p
            print(('ok   ' if ok else 'DIFF ') + tag + '  errors: reference %s (%s), package %s (%s)' % (
                err_r, msg_r if err_r else '', err_p, msg_p if err_p else ''))
            bad += not ok
            continue
        if 'all' not in S_r['m'] or 'all' not in S_p['m']:       # max_iterations = 0: no solve, nothing parsed
            same = set(S_r['m']) == set(S_p['m']) and set(S_r) <= set(S_p) | {'dzdt_lags'} and set(S_r['E']) == set(S_p['E']) \
                and set(S_r['R']) == set(S_p['R']) and csr_digest(fo_r['G_data'].toCSR()) == csr_digest(Gp) \
                and csr_digest(fo_r['Gc'].toCSR()) == csr_digest(Cp) and S_p['valid_data'] is not None \
                and np.array_equal(S_r['valid_data'], S_p['valid_data']) and all(_same_E(key, g, S_p['E']) for key, g in S_r['E'].items())
            print(('ok   ' if same else 'DIFF ') + tag + '  (no iterations: result keys %s / %s, m %s / %s)' % (
                sorted(S_r), sorted(S_p), sorted(S_r['m']), sorted(S_p['m'])))
            bad += not same
            continue
        checks = dict(G=csr_digest(fo_r['G_data'].toCSR()) == csr_digest(Gp), Gc=csr_digest(fo_r['Gc'].toCSR()) == csr_digest(Cp),
                      Ec=np.array_equal(np.asarray(fo_r['Ec'], dtype=float), np.asarray(fo_p['Ec'], dtype=float)),
                      valid=np.array_equal(S_r['valid_data'], S_p['valid_data']),
                      mask=np.array_equal(np.asarray(S_r['data'].three_sigma_edit, dtype=bool), np.asarray(S_p['data'].three_sigma_edit, dtype=bool)),
                      m0=np.linalg.norm(S_r['m']['all'] - S_p['m']['all']) <= 1e-8 * np.linalg.norm(S_r['m']['all']),
                      keys=set(S_r['m'].keys()) == set(S_p['m'].keys()) - {'grid_biases'},      # the reference has no parser for grid biases (m0 covers their columns)
                      R=set(S_r['R']) == set(S_p['R']) and all(_close(S_r['R'][q], S_p['R'][q]) for q in S_r['R']),
                      RMS=set(S_r['RMS']) == set(S_p['RMS']) and all(_close(S_r['RMS'][q], S_p['RMS'][q]) for q in S_r['RMS']))
        for key in ('z0', 'dz'):
            for f in ('count', 'misfit_rms', 'misfit_scaled_rms'):
                a, b = getattr(S_r['m'][key], f), getattr(S_p['m'][key], f)
                checks['%s.%s' % (key, f)] = np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a, b, rtol=1e-6, atol=1e-9, equal_nan=True)
        for key in S_r['m']:
            if key.startswith(('dzdt_', 'avg_')):
                checks[key] = np.allclose(getattr(S_r['m'][key], key), getattr(S_p['m'][key], key), rtol=1e-6, atol=1e-9, equal_nan=True)
        if 'bias' in S_r['m']:
            for f in ('val', 'N_raw', 'N_edited', 'rms_data_raw', 'rms_data_edited', 'expected'):
                checks['bias.' + f] = np.allclose(np.asarray(S_r['m']['bias'][f], dtype=float), np.asarray(S_p['m']['bias'][f], dtype=float),
                                                  rtol=1e-6, atol=1e-8, equal_nan=True)
        for key, gr in S_r['m'].get('sensor_bias_grids', {}).items():
            checks['grid.' + key] = key in S_p['m']['sensor_bias_grids'] and np.allclose(gr.z, S_p['m']['sensor_bias_grids'][key].z, rtol=1e-6, atol=1e-8)
        if S_r['m'].get('slope_bias'):
            checks['slope'] = all(abs(v['slope_x'] - S_p['m']['slope_bias'][q]['slope_x']) < 1e-8 and
                                  abs(v['slope_y'] - S_p['m']['slope_bias'][q]['slope_y']) < 1e-8 for q, v in S_r['m']['slope_bias'].items())
        for key, g in S_r['E'].items():
            checks['E.' + key] = _same_E(key, g, S_p['E'])
        checks['sigma_extra'] = np.allclose(np.asarray(S_r['data'].sigma_extra, dtype=float), S_p['data'].sigma_extra, rtol=0, atol=2e-5)
        failed = [c for c, v in checks.items() if not v]
        print(('ok   ' if not failed else 'DIFF ') + tag + ('' if not failed else '  FAILED: %s' % failed))
        bad += bool(failed)
    print('%d of %d cases differ' % (bad, n_cases))
    return bad


if __name__ == '__main__':
    sys.exit(1 if main() else 0)
