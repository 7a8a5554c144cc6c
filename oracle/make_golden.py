#!/usr/bin/env python
"""Pin the oracle against the UNMODIFIED reference and (re)generate tests/golden/*.npz.

TEST INFRASTRUCTURE ONLY; runs only in the build container (needs /root/reference).

    python oracle/make_golden.py            # check oracle == reference on every case, write fixtures
    python oracle/make_golden.py --check    # check only

For each seeded case it runs the reference's own ``smooth_fit`` (through ``oracle/ref_harness``)
and the restatement in ``oracle/lssurf_oracle`` on identical inputs and requires
  * G_data.toCSR() and Gc.toCSR(): indptr / indices / data identical bit for bit, Ec identical;
  * the edit mask after every outer iteration identical; sigma_extra within 1e-5 absolute
    (Brent's own xatol; the two CPU solvers differ at 1e-12);
  * the final model within 1e-9 relative L2.
The fixtures hold the inputs, sha256 digests of the reference CSR arrays and the per-iteration
reference outputs, so the `-m "not gpu"` tests can re-check the oracle, and the GPU tests the
CUDA product, where /root/reference does not exist.
"""
import argparse
import hashlib
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import lssurf_oracle as O      # noqa: E402
from oracle import ref_harness as H        # noqa: E402
from lssurf_b200 import synthetic as S     # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
WIDENED = ['tracks_sensor_slope', 'uniform_sigma_grid', 'uniform_avg_scales', 'tracks_bias_edit', 'uniform_groups_editable', 'uniform_masked', 'uniform_prior_dem',
           'uniform_avg_masks', 'uniform_scaling_maps', 'tracks_grid_bias',
           'uniform_lagrangian']


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def csr_digest(M):
    M = M.tocsr()
    assert M.has_canonical_format
    return digest(M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float64))


def run_reference(pts, kw):
    """Reference smooth_fit with taps on calc_sigma_extra (per-iteration residual/mask/sigma_extra)."""
    L = H.load()
    SF = H.smooth_fit_module()
    D = H.point_data(**pts)
    fo = L.smooth_fit(data=D.copy(), VERBOSE=False, return_fit_objects=True, **kw)
    trace = []
    orig = SF.calc_sigma_extra

    def tap(r, sigma, mask, sigma_extra_masks=None):
        out = orig(r, sigma, mask, sigma_extra_masks)
        trace.append(dict(r=r.copy(), mask_in=mask.copy(), sigma_extra=out.copy()))
        return out
    SF.calc_sigma_extra = tap
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            out = L.smooth_fit(data=D.copy(), VERBOSE=False, **kw)
    finally:
        SF.calc_sigma_extra = orig
    return fo, out, trace


def parse_model_outputs(out):
    """What parse_model adds to the result (smooth_fit.py:276-352, 676-678; bias_functions.py:140-155): per-node data
    counts and misfits on both grids, chi^2 / RMS per constraint family and for the data, per-bias residual statistics."""
    fx = {}
    for g in ('z0', 'dz'):
        for key in ('count', 'misfit_rms', 'misfit_scaled_rms'):
            fx[f'pm_{g}_{key}'] = np.asarray(getattr(out['m'][g], key), dtype=float)
        fx[f'pm_{g}_cell_area'] = np.asarray(out['m'][g].cell_area, dtype=float)
        fx[f'pm_{g}_mask'] = np.asarray(out['m'][g].mask, dtype=bool)
    fx['pm_R'] = np.array([f'{k}:{float(v)!r}' for k, v in out['R'].items()])
    fx['pm_RMS'] = np.array([f'{k}:{float(v)!r}' for k, v in out['RMS'].items()])
    b = out['m'].get('bias')
    if b:
        for key in ('val', 'expected', 'rms_data_raw', 'rms_data_edited', 'N_raw', 'N_edited'):
            if key in b:
                fx['pm_bias_' + key] = np.asarray(b[key], dtype=float)
    return fx


def with_masks(SF, masks):
    """Context manager: the reference's setup_grids with NumPy masks put on the z0 / dz grids the way its own mask
    branch leaves them (grid_functions.py:72-133: masks on the grids, dz cell areas summed from the masked z0 cells,
    z0 cell areas masked last) -- the stub pointCollection cannot interpolate mask grids, everything downstream
    (validate_by_dz_mask, mask_for_ind0 with mask_scale, cell areas) is the unmodified reference."""
    import contextlib
    GF = sys.modules['LSsurf.grid_functions']

    @contextlib.contextmanager
    def ctx():
        orig = SF.setup_grids

        def patched(args):
            grids, bds = orig(args)
            grids['z0'].mask = masks['z0'].astype(bool)
            grids['dz'].mask = masks['dz'].astype(bool)
            grids['z0'].cell_area = GF.calc_cell_area(grids['z0'])
            grids['dz'].cell_area = GF.sum_cell_area(grids['z0'], grids['dz'])
            grids['z0'].cell_area *= grids['z0'].mask
            return grids, bds
        SF.setup_grids = patched
        try:
            yield
        finally:
            SF.setup_grids = orig
    return ctx() if masks is not None else contextlib.nullcontext()


def case_fit(name, pts, kw, write):
    print(f'== {name}: {pts["x"].size} points')
    fo, out, trace = run_reference(pts, kw)
    G_ref, C_ref = fo['G_data'].toCSR(), fo['Gc'].toCSR()
    hist = []
    bias = [pts[k] for k in kw['bias_params']] if kw.get('bias_params') else None
    res = O.smooth_fit(pts['x'], pts['y'], pts['time'], pts['z'], pts['sigma'], kw['ctr'], kw['W'], kw['spacing'],
                       kw['E_RMS'], reference_epoch=kw['reference_epoch'], max_iterations=kw['max_iterations'],
                       bias_params=bias, sigma_corr=pts.get('sigma_corr'), history=hist)
    s = res['sys']
    n, ncol = s['n_data'], s['col_N']
    # --- structure: bit exact
    Gd = digest(*s['G'][:3])
    Cd = digest(*s['C'][:3])
    assert G_ref.shape == (n, ncol), (G_ref.shape, n, ncol)
    assert Gd == csr_digest(G_ref), 'G_data CSR differs from reference'
    assert C_ref.shape[1] == ncol and Cd == csr_digest(C_ref), 'Gc CSR differs from reference'
    assert digest(s['Ec']) == digest(np.asarray(fo['Ec'], dtype=float)), 'Ec differs'
    assert np.array_equal(res['valid'], out['valid_data'] | res['valid']) and res['valid'].sum() == n
    # --- iterations
    m_ref = out['m']['all']
    n_it = len(trace)
    assert n_it == sum('sigma_extra' in h for h in hist), (n_it, len(hist))
    masks, sig_x, n_close = [], [], []
    for k, tr in enumerate(trace):
        assert np.array_equal(tr['mask_in'], hist[k]['mask_in']), f'mask into iteration {k} differs'
        se_ref = tr['sigma_extra']
        assert np.max(np.abs(se_ref - hist[k]['sigma_extra'])) < 1e-5, f'sigma_extra iteration {k}'
        masks.append(np.packbits(hist[k]['mask_out']))
        sig_x.append(float(np.median(se_ref)))
        aug = np.sqrt(pts['sigma'][res['valid']] ** 2 + se_ref ** 2)
        n_close.append(int(np.sum(np.abs(np.abs(tr['r'] / aug) - 3.0) < 1e-6)))
    tse_ref = np.asarray(out['data'].three_sigma_edit, dtype=bool)
    assert np.array_equal(tse_ref, res['in_TSE']), 'final edit mask differs'
    rel = np.linalg.norm(res['m0'] - m_ref) / np.linalg.norm(m_ref)
    print(f'   G {G_ref.shape} nnz={G_ref.nnz}  Gc {C_ref.shape} nnz={C_ref.nnz}  outer={len(hist)}'
          f'  rel-L2(m0)={rel:.2e}  in_TSE={int(tse_ref.sum())}  near-threshold={n_close}')
    assert rel < 1e-9, rel
    if write:
        fx = dict(G_digest=Gd, C_digest=Cd, Ec_digest=digest(s['Ec']), G_shape=np.array(G_ref.shape),
                  C_shape=np.array(C_ref.shape), G_nnz=G_ref.nnz, C_nnz=C_ref.nnz, m0=m_ref,
                  z_est=np.asarray(out['data'].z_est), in_TSE=np.packbits(tse_ref), n_data=n,
                  sigma_extra=np.array(sig_x), masks=np.array(masks), n_outer=len(hist),
                  near_threshold=np.array(n_close), kw=repr(kw),
                  toc_rows=np.array([f'{k}:{len(np.unique(v))}' for k, v in fo['Gc'].TOC['rows'].items()]))
        fx.update(parse_model_outputs(out))
        fx.update({'in_' + k: v for k, v in pts.items()})
        np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **fx)


def case_lowlevel(write):
    """interp_mtx / diff_op / toCSR on edge inputs (upper-bound points, on-node points, NaN,
    out-of-bounds with bounds_error=False) in 1, 2 and 3 dimensions."""
    print('== lowlevel')
    L = H.load()
    rng = np.random.default_rng(7)
    fx = {}
    specs = {'g1': ([[0., 10.]], [0.25]), 'g2': ([[-3., 5.], [100., 131.]], [0.5, 1.7]),
             'g3': ([[-1e3, 1e3], [-2e3, 1e3], [-1., 1.]], [100., 100., 0.25])}
    for gname, (bounds, deltas) in specs.items():
        col_0 = 17 if gname == 'g3' else 0
        gr = L.fd_grid(bounds, deltas, name=gname, col_0=col_0)
        go = O.grid_geometry(bounds, deltas, col_0=col_0, name=gname)
        assert np.array_equal(gr.shape, go['shape']) and np.array_equal(gr.stride, go['stride'])
        m = 400
        pts = [rng.uniform(b[0] - 0.05 * (b[1] - b[0]), b[1] + 0.05 * (b[1] - b[0]), m) for b in bounds]
        for d in range(len(bounds)):
            pts[d][:20] = gr.ctrs[d][rng.integers(0, gr.shape[d], 20)]       # exactly on nodes
            pts[d][20:30] = gr.bds[d][1]                                      # on the upper bound
            pts[d][30:40] = gr.bds[d][0]
        pts[0][40:44] = np.nan
        op = L.lin_op(gr, name='interp').interp_mtx(pts, bounds_error=False)
        r, c, v = O.interp_triplets(go, pts, bounds_error=False)
        assert np.array_equal(op.r, r) and np.array_equal(op.c, c) and digest(op.v) == digest(v), gname
        M = op.toCSR(row_N=m)
        ip, ix, dat, _ = O.coo_to_csr(r, c, v, n_rows=m, n_cols=gr.col_N)
        assert csr_digest(M) == digest(ip, ix, dat), gname
        try:
            L.lin_op(gr).interp_mtx(pts)
            raise AssertionError('reference did not raise')
        except ValueError as e:
            msg = str(e)
        fx[gname + '_bounds'] = np.array(bounds)
        fx[gname + '_deltas'] = np.array(deltas)
        fx[gname + '_col0'] = col_0
        fx[gname + '_pts'] = np.array(pts)
        fx[gname + '_rcv'] = digest(op.r.astype(np.int64), op.c.astype(np.int64), op.v)
        fx[gname + '_csr'] = csr_digest(M)
        fx[gname + '_nvalid'] = op.N_eq
        fx[gname + '_err'] = msg
        print(f'   {gname}: shape={gr.shape} valid={op.N_eq}/{m} csr nnz={M.nnz} err="{msg}"')
    # stencils: every built-in family + a custom 9-point template (aniso notebook style)
    g2r = L.fd_grid(*specs['g2'], name='z0')
    g3r = L.fd_grid(*specs['g3'], name='dz', col_0=g2r.N_nodes)
    g2o = O.grid_geometry(*specs['g2'], name='z0')
    g3o = O.grid_geometry(*specs['g3'], name='dz', col_0=g2o['n_nodes'])
    fam_dig = []
    for kind, gr_, go_, build in (('grad2_z0', g2r, g2o, lambda o: o.grad2(DOF='z0')),
                                  ('grad_z0', g2r, g2o, lambda o: o.grad(DOF='z0')),
                                  ('grad2_dzdt', g3r, g3o, lambda o: o.grad2_dzdt(DOF='z', t_lag=1)),
                                  ('grad_dzdt', g3r, g3o, lambda o: o.grad_dzdt(DOF='z', t_lag=1)),
                                  ('d2z_dt2', g3r, g3o, lambda o: o.d2z_dt2(DOF='z'))):
        op = build(L.lin_op(gr_, name=kind))
        row_0, rr, cc, vv, i0 = 0, [], [], [], []
        for name, offs, coef in O.stencil_families(go_, kind):
            r, c, v, ind0 = O.stencil_triplets(go_, offs, coef)
            assert np.array_equal(op.TOC['rows'][name], np.arange(row_0, row_0 + r.shape[0])), name
            rr.append(r.ravel() + row_0); cc.append(c.ravel()); vv.append(v.ravel()); i0.append(ind0)
            row_0 += r.shape[0]
        assert np.array_equal(op.r.ravel(), np.concatenate(rr)), kind
        assert np.array_equal(op.c.ravel(), np.concatenate(cc)), kind
        assert digest(op.v.ravel()) == digest(np.concatenate(vv)), kind
        assert np.array_equal(op.ind0, np.concatenate(i0)), kind
        d = digest(op.r.ravel().astype(np.int64), op.c.ravel().astype(np.int64), op.v.ravel(),
                   np.asarray(op.ind0).astype(np.int64))
        fam_dig.append(f'{kind}:{op.N_eq}:{d}')
        print(f'   {kind}: N_eq={op.N_eq}')
    offs = ([-1, -1, -1, 0, 0, 0, 1, 1, 1], [-1, 0, 1, -1, 0, 1, -1, 0, 1])
    coef = np.array([0.5, -1.25, 0.5, 2., -3.5, 2., 0.5, -1.25, 0.5])
    op = L.lin_op(g2r, name='custom').diff_op(offs, coef)
    r, c, v, ind0 = O.stencil_triplets(g2o, offs, coef)
    assert np.array_equal(op.r, r) and np.array_equal(op.c, c) and digest(op.v) == digest(v)
    fam_dig.append('custom9:%d:%s' % (op.N_eq, digest(op.r.astype(np.int64), op.c.astype(np.int64), op.v,
                                                      np.asarray(op.ind0).astype(np.int64))))
    fx['stencils'] = np.array(fam_dig)
    # RDE + sigma_extra known answers
    x = rng.standard_t(3, 5001)
    x[::97] = np.nan
    sg = rng.uniform(0.05, 0.2, x.size)
    msk = rng.random(x.size) < 0.9
    fx['rde_x'], fx['rde_sigma'], fx['rde_mask'] = x, sg, msk
    fx['rde'] = L.RDE(x)
    assert fx['rde'] == O.rde(x) == O.rde_select(x)
    xs = x.copy(); xs[~np.isfinite(xs)] = 0.
    se_ref = L.calc_sigma_extra(xs, sg, msk)
    se_or = O.sigma_extra(xs, sg, msk)
    assert np.array_equal(se_ref, se_or), (se_ref[0], se_or[0])
    fx['sigma_extra'] = se_ref[0]
    print(f'   RDE={fx["rde"]:.6f} sigma_extra={se_ref[0]:.6f} (bit-identical to reference)')
    if write:
        np.savez_compressed(os.path.join(GOLDEN, 'lowlevel.npz'), **fx)


def case_widened(name, pts, kw, write):
    """Cases for the SURVEY section 8(f) rows (sensor grid bias, slope bias, gridded sigma_extra): the fixture is taken
    from the reference run alone (the oracle does not restate these host-level compositions; the product is compared
    with the reference outputs directly).  Same fixture keys as case_fit."""
    print(f'== {name}: {pts["x"].size} points (reference outputs only)')
    L = H.load()
    import sparseqr
    SF = H.smooth_fit_module()
    D = H.point_data(**pts)
    kw = dict(kw)
    grid_masks = kw.pop('mask_data', None)         # NumPy masks: injected into the reference's grids (with_masks)
    prior = kw.pop('prior_args', None)             # in-memory prior points: handed to the reference's make_prior_op
    avg_masks = kw.pop('avg_masks', None)          # named masks as arrays on the dz nodes; the reference asks them to interpolate
    scaling_maps = kw.pop('constraint_scaling_maps', None)     # per-node multipliers as arrays on the z0 / dz nodes
    grid_bias = kw.pop('grid_bias_spec', None)                 # gridded bias x covariate: [param, grid bounds, spacing, rms, rms_grad]
    orig_constraints = SF.setup_smoothness_constraints
    if grid_bias is not None:
        # the live reference driver has no call site for setup_grid_bias (smooth_fit.py:487-489); its predecessor called it
        # right after the smoothness constraints (deprecated/smooth_xytb_fit.py:465-467), which is what this hook does with the
        # driver's own data / G_data objects.  The grid object is rebuilt for every run (setup_grid_bias writes col_0 into it).
        def constraints_then_grid_bias(grids, constraint_op_list, *a, **k):
            orig_constraints(grids, constraint_op_list, *a, **k)
            caller = sys._getframe(1).f_locals
            b = grid_bias['bounds']
            g = L.fd_grid([[b[0], b[1]], [b[2], b[3]]], [grid_bias['spacing']] * 2, name=grid_bias['param'] + '_bias')
            L.setup_grid_bias(caller['data'], caller['G_data'], constraint_op_list, grids, param=grid_bias['param'], grid=g,
                              expected_rms=grid_bias['expected_rms'], expected_value=0., expected_rms_grad=grid_bias['expected_rms_grad'])
        SF.setup_smoothness_constraints = constraints_then_grid_bias
    if avg_masks is not None or scaling_maps is not None:
        class NodeMask(object):                    # what lin_op.mean_of_mask needs of a pointCollection grid (lin_op.py:352-354)
            def __init__(self, z):
                self.z = z

            def interp(self, x, y, gridded=False, **_):
                assert (x.size, y.size)[::-1] == self.z.shape if gridded else x.shape == self.z.shape
                return self.z.copy()
    if avg_masks is not None:
        kw['avg_masks'] = {k: NodeMask(np.asarray(v, dtype=float)) for k, v in avg_masks.items()}
    if scaling_maps is not None:
        kw['constraint_scaling_maps'] = {k: NodeMask(np.asarray(v, dtype=float)) for k, v in scaling_maps.items()}
    if prior is not None:
        MP = sys.modules['LSsurf.match_priors']

        def prior_ops(grids, **_):
            return [MP.make_prior_op(grids, H.point_data(**d), 'prior_%d' % k, ref_time=prior['ref_epoch'], sigma_scale=prior['sigma_scale'])
                    for k, d in enumerate(prior['dzs'])]
        kw['prior_args'] = {}
        SF.match_prior_dz = prior_ops
    with warnings.catch_warnings(), with_masks(SF, grid_masks):
        warnings.simplefilter('ignore')
        fo = L.smooth_fit(data=D.copy(), VERBOSE=False, return_fit_objects=True, **kw)
    G_ref, C_ref = fo['G_data'].toCSR(), fo['Gc'].toCSR()
    trace, solves = [], []
    orig_se, orig_grid, orig_solve = SF.calc_sigma_extra, SF.calc_sigma_extra_on_grid, sparseqr.solve

    def tap(r, sigma, mask, sigma_extra_masks=None):
        out = orig_se(r, sigma, mask, sigma_extra_masks)
        trace.append(dict(mask_in=mask.copy(), sigma_extra=out.copy()))
        return out

    def tap_grid(x, y, r, sigma, mask, **k):
        out = orig_grid(x, y, r, sigma, mask, **k)
        trace.append(dict(mask_in=mask.copy(), sigma_extra=out.copy()))
        return out

    def count_solve(*a, **k):
        solves.append(1)
        return orig_solve(*a, **k)
    SF.calc_sigma_extra, SF.calc_sigma_extra_on_grid, sparseqr.solve = tap, tap_grid, count_solve
    try:
        with warnings.catch_warnings(), with_masks(SF, grid_masks):
            warnings.simplefilter('ignore')
            out = L.smooth_fit(data=D.copy(), VERBOSE=False, **kw)
    finally:
        SF.calc_sigma_extra, SF.calc_sigma_extra_on_grid, sparseqr.solve = orig_se, orig_grid, orig_solve
        SF.setup_smoothness_constraints = orig_constraints
    tse_ref = np.asarray(out['data'].three_sigma_edit, dtype=bool)
    if grid_bias is not None:
        # the oracle restates this composition (it is the CPU arm of the bench): pinned here bit for bit
        res = O.smooth_fit(pts['x'], pts['y'], pts['time'], pts['z'], pts['sigma'], kw['ctr'], kw['W'], kw['spacing'], kw['E_RMS'],
                           reference_epoch=kw['reference_epoch'], max_iterations=kw['max_iterations'],
                           bias_params=[pts[k] for k in kw['bias_params']], sigma_corr=pts.get('sigma_corr'),
                           grid_bias=dict(values=pts[grid_bias['param']], bounds=grid_bias['bounds'], spacing=grid_bias['spacing'],
                                          expected_rms=grid_bias['expected_rms'], expected_rms_grad=grid_bias['expected_rms_grad']))
        assert digest(*res['sys']['G'][:3]) == csr_digest(G_ref), 'oracle G with grid bias differs from the reference'
        assert digest(*res['sys']['C'][:3]) == csr_digest(C_ref), 'oracle Gc with grid bias differs from the reference'
        assert digest(res['sys']['Ec']) == digest(np.asarray(fo['Ec'], dtype=float))
        assert np.array_equal(res['in_TSE'], tse_ref), 'oracle edit mask with grid bias differs'
        rel = np.linalg.norm(res['m0'] - out['m']['all']) / np.linalg.norm(out['m']['all'])
        print(f'   oracle with grid bias == reference: CSR / Ec / mask identical, rel-L2(m0) = {rel:.1e}')
        assert rel < 1e-9
    n = int(out['valid_data'].size) if out['data'].size == out['valid_data'].size else int(out['data'].size)
    masks = [np.packbits(trace[k + 1]['mask_in']) for k in range(len(trace) - 1)] + [np.packbits(tse_ref)]
    m_ref = out['m']['all']
    print(f'   G {G_ref.shape} nnz={G_ref.nnz}  Gc {C_ref.shape} nnz={C_ref.nnz}  solves={len(solves)}'
          f'  in_TSE={int(tse_ref.sum())}  cols: {sorted(fo["G_data"].TOC["cols"])}')
    if write:
        fx = dict(G_digest=csr_digest(G_ref), C_digest=csr_digest(C_ref), Ec_digest=digest(np.asarray(fo['Ec'], dtype=float)),
                  G_shape=np.array(G_ref.shape), C_shape=np.array(C_ref.shape), G_nnz=G_ref.nnz, C_nnz=C_ref.nnz, m0=m_ref,
                  z_est=np.asarray(out['data'].z_est), in_TSE=np.packbits(tse_ref), n_data=n,
                  sigma_extra=np.array([float(np.median(t['sigma_extra'])) for t in trace]),
                  sigma_extra_final=np.asarray(out['data'].sigma_extra, dtype=float), masks=np.array(masks),
                  n_outer=len(solves), kw=repr(kw),
                  toc_rows=np.array([f'{k}:{len(np.unique(v))}' for k, v in fo['Gc'].TOC['rows'].items()]),
                  toc_cols=np.array([f'{k}:{int(np.min(v))}:{int(np.max(v))}' for k, v in fo['G_data'].TOC['cols'].items()]))
        if 'lagrangian_dz' in out['m']:
            fx['lagrangian_dz'] = np.asarray(out['m']['lagrangian_dz'].dz, dtype=float)
        for key, g in out['m'].get('sensor_bias_grids', {}).items():
            fx['grid_' + key] = np.asarray(g.z)
        for key, g in out['m'].items():                      # dz/dt and block-average grids (setup_averaging_ops)
            if key.startswith(('avg_', 'dzdt_')) or '_avg_dz' in key:
                fx['avgop_' + key] = np.asarray(getattr(g, key), dtype=float)
                if getattr(g, 'cell_area', None) is not None:
                    fx['avgarea_' + key] = np.asarray(g.cell_area, dtype=float)
        fx.update(parse_model_outputs(out))
        fx['valid_data'] = np.packbits(np.asarray(out['valid_data'], dtype=bool))
        if grid_masks is not None:
            fx['mask_z0'], fx['mask_dz'] = grid_masks['z0'].astype(bool), grid_masks['dz'].astype(bool)
        if avg_masks is not None:
            kw.pop('avg_masks')
            fx['kw'] = repr(kw)
            for k, v in avg_masks.items():
                fx['avgmask_' + k] = np.asarray(v, dtype=float)
        if grid_bias is not None:
            fx['gridbias_param'] = grid_bias['param']
            fx['gridbias_spec'] = np.array(list(grid_bias['bounds']) + [grid_bias['spacing'], grid_bias['expected_rms'], grid_bias['expected_rms_grad']], dtype=float)
            c0, c1 = (int(v) for v in fo['G_data'].TOC['cols'][grid_bias['param'] + '_bias'][[0, -1]])
            fx['gridbias_values'] = np.asarray(m_ref[c0:c1 + 1], dtype=float)
        if scaling_maps is not None:
            kw.pop('constraint_scaling_maps')
            fx['kw'] = repr(kw)
            for k, v in scaling_maps.items():
                fx['scalemap_' + k] = np.asarray(v, dtype=float)
        if prior is not None:
            kw.pop('prior_args')
            fx['kw'] = repr(kw)
            fx['prior_meta'] = np.array([prior['ref_epoch'], prior['sigma_scale'], len(prior['dzs'])], dtype=float)
            for k, d in enumerate(prior['dzs']):
                for f, v in d.items():
                    fx[f'prior{k}_{f}'] = np.asarray(v, dtype=float)
        sb = out['m'].get('slope_bias', {})
        if sb:
            fx['slope_bias'] = np.array([[k, v['slope_x'], v['slope_y']] for k, v in sb.items()], dtype=float)
        fx.update({'in_' + k: v for k, v in pts.items()})
        np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **fx)


def widened_inputs():
    """Seeded inputs of the two section-8(f) cases."""
    pts, kw = S.icesat2_tracks(n=8000, width=1.2e4, seed=5, t_span=3.)
    rng = np.random.default_rng(11)
    sensor = (pts['BP'] == 2).astype(float)                 # one beam pair is a second 'sensor'
    pts['sensor'] = sensor
    pts['z'] = pts['z'] + sensor * (0.2 * np.sin(pts['x'] / 3000.) + 1e-5 * pts['x'] - 2e-5 * pts['y'])
    kw_a = dict(kw, data_slope_sensors=np.array([1.]), E_slope_bias=0.01,
                sensor_grid_bias_params=[{'sensor': 1., 'spacing': 3000., 'expected_rms': 0.2, 'expected_rms_grad2': 1e-7}])
    ptu, kwu = S.uniform_3d(n=6000, width=1.2e4, seed=4)
    ptu['z'] = ptu['z'] + rng.normal(0, 1., ptu['x'].size) * 0.3 * (ptu['x'] > 0)     # noisier eastern half
    kw_b = dict(kwu, sigma_extra_bin_spacing=4000., sigma_extra_max=5.)
    ptc, kwc = S.uniform_3d(n=6000, width=1.2e4, seed=6)
    kw_c = dict(kwc, dzdt_lags=[1, 4], avg_scales=[2000., 4000.], z0_average_scale=2000., max_iterations=3)
    # edit-loop options: bias editing, pre-edited biases, relaxed last pass
    ptd, kwd = S.icesat2_tracks(n=8000, width=1.2e4, seed=8, t_span=3.)
    inside = np.abs(ptd['time']) < 1.4                     # the fit only sees the epochs inside t_span
    ids = np.unique(np.c_[ptd['rgt'], ptd['cycle']][inside], axis=0)
    big, pre = ids[len(ids) // 3], ids[len(ids) // 2]
    ptd['z'] = ptd['z'] + 3.0 * ((ptd['rgt'] == big[0]) & (ptd['cycle'] == big[1]))        # one implausible bias
    ptd['three_sigma_edit'] = np.ones_like(ptd['z'])       # bias_edit_vals writes into it (smooth_fit.py:508)
    kw_d = dict(kwd, bias_nsigma_edit=6., bias_nsigma_iteration=1, sigma_extra_relax=True,
                bias_edit_vals={'rgt': [float(pre[0])], 'cycle': [float(pre[1])]})
    # sigma_extra per group, points that may not be edited, and an initial edit mask
    pte, kwe = S.uniform_3d(n=6000, width=1.2e4, seed=9)
    grp = (rng.random(pte['x'].size) < 0.4).astype(float)
    pte['sensor'] = grp
    pte['z'] = pte['z'] + grp * rng.normal(0, 0.4, grp.size)
    pte['editable'] = (rng.random(grp.size) < 0.9).astype(float)
    pte['three_sigma_edit'] = (rng.random(grp.size) < 0.98).astype(float)
    kw_e = dict(kwe, sigma_extra_keys={'quiet': {'sensor': [0.]}, 'noisy': {'sensor': [1.]}})
    # masks with holes on both grids: points under the dz hole are dropped (validate_by_dz_mask),
    # constraint rows centred outside the masks are loosened by mask_scale (mask_for_ind0), cell areas follow the masks
    ptf, kwf = S.uniform_3d(n=6000, width=1.2e4, seed=12)
    gz = (int(1.2e4 / kwf['spacing']['z0']) + 1,) * 2
    gd = (int(1.2e4 / kwf['spacing']['dz']) + 1,) * 2
    mz, md = np.ones(gz, dtype=bool), np.ones(gd, dtype=bool)
    mz[3:9, 14:21] = False                                  # a lake in the z0 grid
    mz[:, :2] = False                                       # and a strip along its western edge
    md[8:, 9:] = False                                      # the north-east corner of the dz grid
    kw_f = dict(kwf, mask_data={'z0': mz, 'dz': md}, mask_scale={0: 10, 1: 1}, dzdt_lags=[1], max_iterations=4)
    # prior constraints from an existing dz model (match_priors.make_prior_op: interpolation at t minus interpolation at the
    # reference time, non-zero prior values in the right-hand side) and the DEM check of the edit loop (DEM_tol)
    ptg, kwg = S.uniform_3d(n=6000, width=1.2e4, seed=13)
    ptg['DEM'] = 100. * np.cos(2 * np.pi * ptg['x'] / 2.e4) * np.sin(2 * np.pi * ptg['y'] / 3.e4) + rng.normal(0, 0.5, ptg['x'].size)
    far = rng.random(ptg['x'].size) < 0.02
    ptg['DEM'][far] += 40.                                  # DEM blunders: those points fail the check
    npr = 150
    px, py = rng.uniform(-5.5e3, 5.5e3, npr), rng.uniform(-5.5e3, 5.5e3, npr)
    pt = rng.choice([-1.5, -0.75, 0., 0.75, 1.5], npr)       # the reference-time rows (t = 0) are dropped by make_prior_op
    prior_pts = dict(x=px, y=py, t=pt, dz=pt * (1. + np.cos(2 * np.pi * px / 4.e4)) + rng.normal(0, 0.02, npr), sigma_dz=np.full(npr, 0.05))
    kw_g = dict(kwg, DEM_tol=10., prior_args=dict(dzs=[prior_pts], ref_epoch=0., sigma_scale=2.), max_iterations=5)
    # means of dz and dz/dt over named masks (grid_functions.setup_avg_mask_ops, lin_op.mean_of_mask): a fractional mask
    # with NaNs outside, weighted by the cell areas
    pth, kwh = S.uniform_3d(n=6000, width=1.2e4, seed=14)
    ndz = int(1.2e4 / kwh['spacing']['dz']) + 1
    iy, ix = np.meshgrid(np.arange(ndz), np.arange(ndz), indexing='ij')
    basin = np.clip(1.2 - np.hypot(iy - 4., ix - 7.) / 4., 0., 1.)
    basin[:, :2] = np.nan
    shelf = (ix >= 9).astype(float)
    kw_h = dict(kwh, dzdt_lags=[1, 4], avg_masks={'basin': basin, 'shelf': shelf}, max_iterations=3)
    # per-node multipliers of the expected constraint values (smooth_fit.py:467-478, constraint_functions.py:40-90): stiffer
    # z0 curvature in the south, looser dz curvature in the east, d2z/dt2 scaled in a corner, and a 'dz' magnitude
    # constraint on the nodes where its map is positive
    pti, kwi = S.uniform_3d(n=6000, width=1.2e4, seed=15)
    nz0 = int(1.2e4 / kwi['spacing']['z0']) + 1
    m_z0 = np.ones((nz0, nz0)); m_z0[:nz0 // 2] = 0.25; m_z0[2, 3] = np.nan
    m_d3 = np.ones((ndz, ndz)); m_d3[:, ndz // 2:] = 4.
    m_tt = np.ones((ndz, ndz)); m_tt[:3, :3] = 0.5
    m_dz = np.zeros((ndz, ndz)); m_dz[:2, :] = 0.05; m_dz[5, 5] = np.nan
    kw_i = dict(kwi, constraint_scaling_maps={'d2z0_dx2': m_z0, 'd3z_dx2dt': m_d3, 'd2z_dt2': m_tt, 'dz': m_dz}, max_iterations=3)
    # a gridded bias that multiplies a per-point covariate (setup_grid_bias.py:4-43; BASELINE configs[3]) next to the
    # per-(rgt, cycle) biases: 17-entry rows, columns outside the z0 / dz grids and outside the eliminated block
    ptj, kwj = S.icesat2_tracks(n=8000, width=1.2e4, seed=16, t_span=3.)
    ptj['cov'] = np.sin(ptj['x'] / 2500.) + 0.1 * rng.standard_normal(ptj['x'].size)
    ptj['z'] = ptj['z'] + ptj['cov'] * 0.15 * np.cos(ptj['y'] / 4000.)
    kw_j = dict(kwj, grid_bias_spec=dict(param='cov', bounds=(-6e3, 6e3, -6e3, 6e3), spacing=3000., expected_rms=0.1, expected_rms_grad=1e-3))
    # height offsets on a grid in coordinates that move with the ice (lagrangian_coords: smooth_fit.py:463-465,
    # grid_functions.py:89-93, constraint_functions.py:95-105): two thirds of the points carry such coordinates, the
    # rest NaN (no entry in those rows), a few fall outside the grid
    ptk, kwk = S.uniform_3d(n=6000, width=1.2e4, seed=18)
    has = rng.random(ptk['x'].size) < 0.66
    ptk['x0'] = np.where(has, ptk['x'] - 150. * ptk['time'], np.nan)          # advected back to t = 0
    ptk['y0'] = np.where(has, ptk['y'] + 60. * ptk['time'], np.nan)
    ptk['z'] = ptk['z'] + np.where(has, 0.5 * np.sin(ptk['x0'] / 1500.) * np.cos(ptk['y0'] / 2500.), 0.)
    kw_k = dict(kwk, lagrangian_coords=['y0', 'x0'], spacing=dict(kwk['spacing'], lagrangian_dz=1000.),
                E_RMS=dict(kwk['E_RMS'], lagrangian_dz=50., lagrangian_dzdx=0.5), max_iterations=4)
    return ((pts, kw_a), (ptu, kw_b), (ptc, kw_c), (ptd, kw_d), (pte, kw_e), (ptf, kw_f), (ptg, kw_g), (pth, kw_h), (pti, kw_i),
            (ptj, kw_j), (ptk, kw_k))


def grid_bias_inputs():
    rng = np.random.default_rng(21)
    n = 3000
    x, y, t = rng.uniform(-6e3, 6e3, n), rng.uniform(-6e3, 6e3, n), rng.uniform(-1.5, 1.5, n)
    cov = np.sin(x / 2000.) + 0.1 * rng.standard_normal(n)
    return dict(x=x, y=y, time=t, z=np.zeros(n), sigma=np.ones(n), cov=cov)


def case_grid_bias(write):
    """setup_grid_bias at the operator level (it is not reachable through smooth_fit in the reference either,
    smooth_fit.py:488-490): G_data = z0 + dz interpolation rows + the covariate-scaled bias grid, curvature and magnitude
    constraint rows.  Digests of the reference operators only."""
    print('== grid_bias (reference outputs only)')
    L = H.load()
    pts = grid_bias_inputs()
    D = H.point_data(**pts)
    z0 = L.fd_grid([[-6e3, 6e3], [-6e3, 6e3]], [500., 500.], name='z0')
    dz = L.fd_grid([[-6e3, 6e3], [-6e3, 6e3], [-1.5, 1.5]], [1000., 1000., 0.25], name='dz', col_0=z0.N_nodes)
    grids = {'z0': z0, 'dz': dz}
    G = L.lin_op(z0, name='interp_z').interp_mtx(D.coords()[0:2])
    G.add(L.lin_op(dz, name='interp_dz').interp_mtx(D.coords()))
    ops = []
    bg = L.fd_grid([[-6e3, 6e3], [-6e3, 6e3]], [2000., 2000.], name='cov_bias')
    L.setup_grid_bias(D, G, ops, grids, param='cov', grid=bg, expected_rms=0.1, expected_value=0., expected_rms_grad=1e-3)
    Gc = L.lin_op(None, name='constraints').vstack(ops)
    fx = dict(G_digest=csr_digest(G.toCSR()), G_shape=np.array(G.toCSR().shape), C_digest=csr_digest(Gc.toCSR()),
              C_shape=np.array(Gc.toCSR().shape), expected=np.concatenate([op.expected for op in ops]),
              names=np.array([op.name for op in ops]), col_range=np.array([bg.col_0, bg.col_N]),
              toc_rows=np.array([f'{k}:{len(np.unique(v))}' for k, v in Gc.TOC['rows'].items()]))
    fx.update({'in_' + k: v for k, v in pts.items()})
    print(f'   G {tuple(fx["G_shape"])}  Gc {tuple(fx["C_shape"])}  ops {list(fx["names"])}  cols {fx["col_range"]}')
    if write:
        np.savez_compressed(os.path.join(GOLDEN, 'grid_bias.npz'), **fx)


def compute_E_inputs():
    """A fit small enough for the dense QR / triangular-inverse stand-ins (about 1 300 unknowns): repeat tracks with
    per-(rgt, cycle) biases, dz/dt at two lags and block averages, so every branch of calc_and_parse_errors is taken."""
    pts, kw = S.icesat2_tracks(n=4000, width=6.e3, seed=17, t_span=3.)
    return pts, dict(kw, compute_E=True, dzdt_lags=[1, 4], avg_scales=[2000.], max_iterations=4)


def case_compute_E(write):
    """Error propagation (smooth_fit.py:212-274, 689-712, lin_op.py:540-553) from the reference run alone:
    sparseqr.rz and LSsurf.inv_tr_upper are dense stand-ins (oracle/ref_stubs/sparseqr, oracle/ref_harness), everything
    else -- the rebuilt TCinv with sigma_extra, Ip_r, the row norms of Rinv, grid_error of every averaging operator,
    parse_biases of the errors -- is the unmodified reference."""
    pts, kw = compute_E_inputs()
    print(f'== compute_E: {pts["x"].size} points (reference outputs only)')
    L = H.load()
    D = H.point_data(**pts)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        out = L.smooth_fit(data=D.copy(), VERBOSE=False, **kw)
    E = out['E']
    fx = dict(m0=out['m']['all'], in_TSE=np.packbits(np.asarray(out['data'].three_sigma_edit, dtype=bool)),
              sigma_extra_final=np.asarray(out['data'].sigma_extra, dtype=float), kw=repr(kw))
    for key, g in E.items():
        if key in ('sigma_bias', 'sigma_slope_bias'):
            continue
        fx['E_' + key] = np.asarray(getattr(g, key), dtype=float)
    sb = E.get('sigma_bias')
    if sb:
        fx['E_sigma_bias_val'] = np.asarray(sb['val'], dtype=float)
    print('   E keys:', sorted(E), ' sigma_z0 median %.4f  sigma_dz median %.4f' %
          (np.nanmedian(fx['E_sigma_z0']), np.nanmedian(fx['E_sigma_dz'])))
    fx.update({'in_' + k: v for k, v in pts.items()})
    if write:
        np.savez_compressed(os.path.join(GOLDEN, 'compute_E.npz'), **fx)


def case_notebook_solves(write):
    """BASELINE.json configs[0..1] at the one-line seam sparseqr.solve(A, b): the systems of tests/notebook_cases.py built
    with the unmodified reference classes, solved by the stub sparseqr (SuperLU on the normal equations + refinement).
    The fixture holds the digest of every scaled system and its solution (every 4th node of the 401 x 401 grids)."""
    print('== notebook solves (cfg 1b, cfg 2)')
    L = H.load()
    import sparseqr
    from tests import notebook_cases
    systems = notebook_cases.build(L, lambda d: H.point_data(**d))
    fx = {}
    for name, (A, b, shape) in systems.items():
        x = sparseqr.solve(A, b)
        r = A @ x - b
        g = A.T @ r
        print(f'   {name}: A {A.shape} nnz={A.nnz}  |A^T r| / |A^T b| = {np.linalg.norm(g) / np.linalg.norm(A.T @ b):.1e}  {sparseqr.LAST.get("method")}')
        step = 4 if shape[0] > 200 else 1
        fx[name + '_A'] = csr_digest(A)
        fx[name + '_b'] = digest(np.asarray(b, dtype=float))
        fx[name + '_x'] = x.reshape(shape)[::step, ::step].copy()
        fx[name + '_step'] = step
        fx[name + '_norm'] = float(np.linalg.norm(x))
    if write:
        np.savez_compressed(os.path.join(GOLDEN, 'notebook_solves.npz'), **fx)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--check', action='store_true')
    ap.add_argument('--only-widened', action='store_true')
    ap.add_argument('--case', default=None)
    a = ap.parse_args()
    write = not a.check
    os.makedirs(GOLDEN, exist_ok=True)
    if a.only_widened:
        for name, (p_, k_) in zip(WIDENED, widened_inputs()):
            if a.case in (None, name):
                case_widened(name, p_, k_, write)
        if a.case in (None, 'grid_bias'):
            case_grid_bias(write)
        if a.case in (None, 'notebook_solves'):
            case_notebook_solves(write)
        if a.case in (None, 'compute_E'):
            case_compute_E(write)
        return
    case_lowlevel(write)
    case_fit('strip_nb', *S.notebook_strip(), write)
    case_fit('strip_outliers', *S.notebook_strip(outliers=True, seed=0), write)
    case_fit('uniform3d_small', *S.uniform_3d(n=6000, width=1.2e4, seed=1), write)
    pts, kw = S.icesat2_tracks(n=8000, width=1.2e4, seed=2, t_span=3.)
    case_fit('tracks_bias_small', pts, kw, write)
    for name, (p_, k_) in zip(WIDENED, widened_inputs()):
        case_widened(name, p_, k_, write)
    case_grid_bias(write)
    case_notebook_solves(write)
    case_compute_E(write)
    print('all cases: oracle == reference')


if __name__ == '__main__':
    main()
