"""Helpers shared by the CPU and GPU test tiers."""
import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def digest(*arrays):
    """sha256 over dtype, shape and bytes of each array (same recipe as oracle/make_golden.py)."""
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def csr_digest(M):
    M = M.tocsr()
    return digest(M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float64))


def load_case(name):
    fx = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    pts = {k[3:]: fx[k] for k in fx.files if k.startswith('in_') and k != 'in_TSE'}
    kw = eval(str(fx['kw']), {'np': np, 'array': np.array, 'float64': np.float64})   # repr() of plain dicts
    kw.pop('mask_data', None)
    if 'mask_z0' in fx.files:            # NumPy masks of the z0 / dz grids (oracle/make_golden.py::with_masks)
        kw['mask_data'] = {'z0': fx['mask_z0'], 'dz': fx['mask_dz']}
    if any(f.startswith('avgmask_') for f in fx.files):      # named averaging masks, values on the dz nodes
        kw['avg_masks'] = {f[len('avgmask_'):]: fx[f] for f in fx.files if f.startswith('avgmask_')}
    if any(f.startswith('scalemap_') for f in fx.files):     # constraint scaling maps, values on the z0 / dz nodes
        kw['constraint_scaling_maps'] = {f[len('scalemap_'):]: fx[f] for f in fx.files if f.startswith('scalemap_')}
    if 'gridbias_spec' in fx.files:      # gridded bias x covariate (setup_grid_bias through the keyword of the deprecated driver)
        import lssurf_b200
        y0, y1, x0, x1, spacing, rms, rms_grad = (float(v) for v in fx['gridbias_spec'])
        param = str(fx['gridbias_param'])
        kw.pop('grid_bias_spec', None)
        kw['grid_bias_model_args'] = [dict(param=param, expected_rms=rms, expected_value=0., expected_rms_grad=rms_grad,
                                           grid=lssurf_b200.fd_grid([[y0, y1], [x0, x1]], [spacing, spacing], name=param + '_bias'))]
    if 'prior_meta' in fx.files:         # in-memory prior points of match_priors.match_prior_dz
        ref_epoch, sigma_scale, n_sets = fx['prior_meta']
        dzs = []
        for k in range(int(n_sets)):
            dzs.append({f[len(f'prior{k}_'):]: fx[f] for f in fx.files if f.startswith(f'prior{k}_')})
        kw['prior_args'] = dict(dzs=dzs, ref_epoch=float(ref_epoch), sigma_scale=float(sigma_scale))
    return fx, pts, kw


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


FIT_CASES = ['strip_nb', 'strip_outliers', 'uniform3d_small', 'tracks_bias_small']
# fixtures taken from the reference run alone (oracle/make_golden.py::case_widened)
WIDENED_CASES = ['uniform_avg_scales', 'uniform_groups_editable', 'tracks_bias_edit', 'uniform_sigma_grid', 'tracks_sensor_slope',
                 'uniform_masked', 'uniform_prior_dem', 'uniform_avg_masks', 'uniform_scaling_maps', 'tracks_grid_bias',
                 'uniform_lagrangian']


def check_parse_model_outputs(fx, S):
    """parse_model outputs against the reference run (smooth_fit.py:276-352, 676-678; bias_functions.py:140-155): per-node
    data counts and misfits on both grids, masks and cell areas, chi^2 / RMS per constraint family and for the data,
    per-bias residual statistics; the valid-point mask where the fixture holds one."""
    assert 'pm_z0_count' in fx.files
    for g in ('z0', 'dz'):
        for key in ('count', 'misfit_rms', 'misfit_scaled_rms'):
            want, got = fx[f'pm_{g}_{key}'], np.asarray(getattr(S['m'][g], key), dtype=float)
            assert got.shape == want.shape and np.array_equal(np.isnan(got), np.isnan(want)), (g, key)
            if np.any(~np.isnan(want)):
                assert np.nanmax(np.abs(got - want)) <= 1e-6 * max(1., np.nanmax(np.abs(want))), (g, key)
        assert np.allclose(S['m'][g].cell_area, fx[f'pm_{g}_cell_area'], rtol=1e-12, atol=0), g
        assert np.array_equal(np.asarray(S['m'][g].mask, dtype=bool), fx[f'pm_{g}_mask']), g
    for name, got in (('pm_R', S['R']), ('pm_RMS', S['RMS'])):
        want = {s.split(':')[0]: float(s.split(':')[1]) for s in fx[name]}
        assert set(got) == set(want), (name, sorted(got), sorted(want))
        for k, v in want.items():
            assert abs(float(got[k]) - v) <= 1e-6 * max(1., abs(v)) or (np.isnan(v) and np.isnan(got[k])), (name, k, got[k], v)
    for key in (k for k in fx.files if k.startswith('pm_bias_')):
        want, got = fx[key], np.asarray(S['m']['bias'][key[8:]], dtype=float)
        assert got.shape == want.shape and np.array_equal(np.isnan(got), np.isnan(want)), key
        if np.any(~np.isnan(want)):
            assert np.nanmax(np.abs(got - want)) <= 1e-6 * max(1., np.nanmax(np.abs(want))), key
    if 'valid_data' in fx.files:
        assert np.array_equal(np.packbits(np.asarray(S['valid_data'], dtype=bool)), fx['valid_data'])
